"""EnvSpec only: the part of mjrl/utils/gym_env.py (:9-13) the update path needs to size the policy and
baseline.  Environments / MuJoCo stepping stay on the host in mjrl itself (out of scope)."""


class EnvSpec:
    """(observation_dim, action_dim, horizon) -- positional arguments as in the reference's constructor."""

    def __init__(self, obs_dim, act_dim, horizon):
        self.observation_dim, self.action_dim, self.horizon = int(obs_dim), int(act_dim), horizon

    def __repr__(self):
        return "EnvSpec(obs_dim=%d, act_dim=%d, horizon=%r)" % (self.observation_dim, self.action_dim, self.horizon)
