"""EnvSpec only: the part of mjrl/utils/gym_env.py (:9-13) the update path needs to size the policy and
baseline.  Environments / MuJoCo stepping stay on the host in mjrl itself (out of scope)."""


class EnvSpec(object):
    def __init__(self, obs_dim, act_dim, horizon):
        self.observation_dim = obs_dim
        self.action_dim = act_dim
        self.horizon = horizon
