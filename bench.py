#!/usr/bin/env python
"""Benchmark of the post-rollout NPG/TRPO/DAPG update path (BASELINE.json metric: train_step/s and FVP/s on a
1e6-timestep batch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3] [--impl reference]

One JSON line on stdout (rank 0).  A "step" = everything mjrl's train_step does after sampling
(algos/batch_reinforce.py:94-112): returns -> baseline predict -> GAE -> whitening -> VPG -> 10-iteration CG
over Fisher-vector products -> step (+ TRPO line search) -> surrogate/KL re-evaluation -> MLPBaseline.fit
(1 epoch of sequential minibatch Adam).  `value` times it with the trajectories already resident in HBM (CUDA
events on the engine's stream, max over ranks); `e2e` times the public drop-in API (TRPO.update_from_paths(paths))
from host float64 path dicts, host<->device copies included.  Strong scaling: the 1e6-timestep batch is sharded
by trajectory over the ranks; one NCCL all-reduce of the flat gradient and of every FVP result.

`--impl reference` (and the `cpu_baseline` object of the default arm) time the CPU restatement of the reference
(oracle/npg_oracle.py, torch-autograd flavour: two forwards + double backward per FVP, per-path Python loops,
sequential Adam) on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import faulthandler

import numpy as np

faulthandler.enable()
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: obs, act, hidden, n_traj, horizon, algo            (BASELINE.json configs[0..4])
    "cfg1": dict(obs=6, act=2, hidden=(32, 32), n_traj=5, horizon=50, algo="npg", name="point_mass NPG 32x32 5x50"),
    "cfg2": dict(obs=8, act=2, hidden=(64, 64), n_traj=100, horizon=1000, algo="npg", name="Swimmer-v3 NPG 64x64 1e5"),
    "cfg3": dict(obs=17, act=6, hidden=(128, 128), n_traj=1000, horizon=1000, algo="trpo",
                 name="HalfCheetah-v3 TRPO 128x128 1e6"),
    "cfg4": dict(obs=39, act=28, hidden=(256, 256), n_traj=1000, horizon=200, algo="dapg",
                 name="Adroit door-v0 DAPG 256x256 2e5"),
    "cfg5": dict(obs=376, act=17, hidden=(), n_traj=500, horizon=1000, algo="npg", name="Humanoid-v3 linear NPG 5e5"),
}
GAMMA, LAM, CG_ITERS, DAMPING = 0.995, 0.97, 10, 1e-4
NPG_STEP, KL_DIST = 0.05, 0.01
VF = dict(reg_coef=1e-3, batch_size=64, epochs=1, learn_rate=1e-3)


def make_paths(cfg, first, count, seed=0):
    """Deterministic synthetic trajectories (float64 like the sampler delivers); path i depends only on (seed, i)
    so a rank can build its own shard."""
    paths = []
    for i in range(first, first + count):
        rng = np.random.RandomState((seed * 1000003 + i) % (2 ** 31 - 1))
        T = cfg["horizon"]
        paths.append(dict(observations=rng.randn(T, cfg["obs"]), actions=rng.randn(T, cfg["act"]),
                          rewards=rng.randn(T), terminated=False))
    return paths


def flops_per_sample_fvp(cfg):
    sizes = (cfg["obs"],) + tuple(cfg["hidden"]) + (cfg["act"],)
    P = sum(a * b for a, b in zip(sizes[:-1], sizes[1:]))
    if len(cfg["hidden"]) == 0:
        return 4 * P
    return 10 * P - 4 * sizes[0] * sizes[1]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# ======================================================================================= CPU reference arm
def cpu_reference_step_fn(cfg, n_traj_sample):
    """Returns (step_fn, n_samples): one post-rollout step of the CPU restatement on `n_traj_sample` trajectories."""
    import torch
    from oracle import npg_oracle as O          # the CPU baseline leg is the one place bench.py runs the oracle
    paths0 = make_paths(cfg, 0, n_traj_sample)
    spec = O.PolicySpec(cfg["obs"], cfg["act"], cfg["hidden"])
    state = dict(theta=O.init_policy_params(spec, 500), vf=O.VFState(cfg["obs"], (128, 128), seed=1))
    demo = make_paths(cfg, 10 ** 6, max(1, n_traj_sample // 40)) if cfg["algo"] == "dapg" else None

    def step():
        paths = [dict(p) for p in paths0]
        O.compute_returns(paths, GAMMA)
        O.compute_advantages(paths, lambda p: O.vf_predict(state["vf"], p), GAMMA, LAM)
        obs = np.concatenate([p["observations"] for p in paths])
        act = np.concatenate([p["actions"] for p in paths])
        adv = O.whiten(np.concatenate([p["advantages"] for p in paths]))
        gb = None
        if demo is not None:
            gb = O.dapg_batch(obs, act, adv, np.concatenate([p["observations"] for p in demo]),
                              np.concatenate([p["actions"] for p in demo]), 1.0, 0.95, 0.0)
        out = O.policy_update(spec, state["theta"], obs, act, adv, cfg["algo"], step_size=NPG_STEP, kl_dist=KL_DIST,
                              cg_iters=CG_ITERS, damping=DAMPING, dtype=torch.float32, autograd=True, grad_batch=gb)
        state["theta"] = out["new_params"]
        perm = [np.random.permutation(obs.shape[0]) for _ in range(VF["epochs"])]
        O.vf_fit_torch(state["vf"], paths, perm, VF["epochs"], VF["batch_size"], VF["learn_rate"], VF["reg_coef"])
        return out

    return step, n_traj_sample * cfg["horizon"]


def cpu_fvp_time(cfg, n_traj_sample, reps=3):
    import torch
    from oracle import npg_oracle as O
    paths = make_paths(cfg, 0, n_traj_sample)
    spec = O.PolicySpec(cfg["obs"], cfg["act"], cfg["hidden"])
    theta = O.init_policy_params(spec, 500)
    obs = np.concatenate([p["observations"] for p in paths])
    v = np.random.RandomState(1).randn(spec.d).astype(np.float32)
    O.fvp(spec, theta, obs, v, DAMPING, torch.float32, autograd=True)
    t0 = time.time()
    for _ in range(reps):
        O.fvp(spec, theta, obs, v, DAMPING, torch.float32, autograd=True)
    return (time.time() - t0) / reps


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    import torch
    n_s = min(cfg["n_traj"], max(5, 100000 // cfg["horizon"]))      # <= 1e5 timesteps: large enough to be in the linear regime
    step, n = cpu_reference_step_fn(cfg, n_s)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    scale = (cfg["n_traj"] * cfg["horizon"]) / n
    value = 1.0 / (dt * scale)
    fvp_t = cpu_fvp_time(cfg, n_s) * scale
    cores = torch.get_num_threads()
    sample = ("%d of %d trajectories (%d timesteps) per step; time extrapolated linearly x%.0f to the full batch"
              % (n_s, cfg["n_traj"], n, scale))
    line = {"impl": "reference", "metric": "train_step_per_sec", "value": value, "unit": "train_step/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args, world), "fvp_per_sec": 1.0 / fvp_t,
            "cpu_baseline": {"value": value, "unit": "train_step/s", "cores": cores, "os_cpu_count": os.cpu_count(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "train_step/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(cfg, args, world):
    return {"workload": cfg["name"], "obs_dim": cfg["obs"], "act_dim": cfg["act"], "hidden": list(cfg["hidden"]),
            "n_traj": cfg["n_traj"], "horizon": cfg["horizon"], "timesteps": cfg["n_traj"] * cfg["horizon"],
            "algo": cfg["algo"], "cg_iters": CG_ITERS, "damping": DAMPING, "vf_epochs": VF["epochs"],
            "parallelism": "dp%d (trajectory shards, NCCL all-reduce of flat gradient + each FVP)" % world,
            "cache": "batch < L2 on purpose of the workload: obs stays L2-resident across the 10 CG FVPs of a step as in "
                     "production; every step also streams the 1e6-row fit gather + GAE arrays (> L2 in total)"}


# ======================================================================================= GPU arm
def run_gpu(args, cfg, rank, world, local_rank):
    import torch
    from mjrl_b200.algos.dapg import DAPG
    from mjrl_b200.algos.npg_cg import NPG
    from mjrl_b200.algos.trpo import TRPO
    from mjrl_b200.baselines.mlp_baseline import MLPBaseline
    from mjrl_b200.engine import DEMO
    from mjrl_b200.policies.gaussian_linear import LinearPolicy
    from mjrl_b200.policies.gaussian_mlp import MLP
    from mjrl_b200.utils.gym_env import EnvSpec
    from mjrl_b200 import runtime
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mjrl_b200 has no CPU fallback (use --impl reference for the CPU arm)")

    # ---- shard by trajectory (contiguous ranges, strong scaling) ----
    per = [cfg["n_traj"] // world + (1 if r < cfg["n_traj"] % world else 0) for r in range(world)]
    first = sum(per[:rank])
    paths = make_paths(cfg, first, per[rank])
    n_local = per[rank] * cfg["horizon"]
    n_glob = cfg["n_traj"] * cfg["horizon"]
    demo = make_paths(cfg, 10 ** 6 + first, max(1, per[rank] // 40)) if cfg["algo"] == "dapg" else None

    es = EnvSpec(cfg["obs"], cfg["act"], cfg["horizon"])
    pol = LinearPolicy(es, seed=500) if len(cfg["hidden"]) == 0 else MLP(es, hidden_sizes=cfg["hidden"], seed=500)
    torch.manual_seed(1)
    bl = MLPBaseline(es, **VF)
    kw = dict(FIM_invert_args={"iters": CG_ITERS, "damping": DAMPING}, save_logs=False)
    if cfg["algo"] == "trpo":
        agent = TRPO(None, pol, bl, kl_dist=KL_DIST, **kw)
        agent.verbose = False            # the reference prints one line per backtrack; stdout carries the JSON line here
    elif cfg["algo"] == "dapg":
        agent = DAPG(None, pol, bl, demo_paths=demo, kl_dist=KL_DIST, **kw)
    else:
        agent = NPG(None, pol, bl, normalized_step_size=NPG_STEP, **kw)
    eng = agent._eng(n_local + (sum(len(p["actions"]) for p in demo) if demo else 0), len(paths))
    agent._push_policy(eng)
    bl._eng()

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: inputs resident in HBM ----------------
    eng.upload_paths(paths)
    if demo:
        eng.upload_paths(demo, which=DEMO)
    step_args = dict(cg_iters=CG_ITERS, damping=DAMPING)
    if cfg["algo"] == "npg":
        step_args.update(step_size=NPG_STEP)
    else:
        step_args.update(step_size=KL_DIST, demo_lam=1.0)
    stats = []

    def device_step():
        eng.compute_returns(GAMMA)
        # host RNG draw, as optimize_model.py:22 (same order and RNG state as np.random.permutation(n), batched loops)
        perm = runtime.global_permutation(n_glob)
        eng.vf_fit_begin(perm, VF["batch_size"], VF["learn_rate"], VF["reg_coef"])   # side stream: needs only the returns
        eng.vf_predict(prefit=True)           # pre-fit baseline, as in the reference's program order
        eng.compute_advantages(GAMMA, LAM)
        eng.process_paths()
        st = eng.step(cfg["algo"], **step_args)
        eng.vf_fit_end()
        stats.append(st)

    np.random.seed(0)
    for _ in range(args.warmup):
        device_step()
    del stats[:]
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.kernel_launches()
    t_wall = time.time()
    eng.event_record(0)
    for _ in range(args.steps):
        device_step()
    eng.event_record(1)
    ms = eng.event_elapsed_ms(0, 1)
    barrier()
    wall = time.time() - t_wall
    launches = eng.kernel_launches() - launches0
    clk = clocks.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    fvp_ms_kernel = float(np.mean([s.fvp_kernel_ms_sum / max(1, s.fvp_launches) for s in stats]))
    backtracks = [int(s.backtracks) for s in stats]
    phase = {k: float(np.mean([getattr(s, k) for s in stats])) for k in ("time_vpg_ms", "time_npg_ms", "time_eval_ms")}

    # ---------------- FVP/s: device-resident CG (10 x {FVP + all-reduce + fused update}) ----------------
    g = eng.vpg() if cfg["algo"] != "dapg" else eng.vpg(True, 1.0)
    eng.cg(g, iters=CG_ITERS, damping=DAMPING)
    barrier()
    reps = 5
    eng.event_record(2)
    for _ in range(reps):
        eng.lib.mjb_policy_cg(eng.h, None, CG_ITERS, DAMPING, 0.0, None, 0, None)
    eng.event_record(3)
    cg_ms = max_over_ranks(eng.event_elapsed_ms(2, 3)) / reps
    fvp_per_sec = 1e3 / (cg_ms / CG_ITERS)

    # ---------------- e2e: public API from host float64 path dicts ----------------
    def e2e_step():
        fresh = [dict(p) for p in paths]                 # new list object => uploaded again, dicts re-populated
        agent.update_from_paths(fresh, GAMMA, LAM)

    e2e_warm = max(1, min(args.warmup, 3))
    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(e2e_warm):
        e2e_step()
    barrier()
    t0 = time.time()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.time() - t0) / e2e_steps)
    demo_n = sum(len(p["actions"]) for p in demo) if demo else 0
    h2d = n_local * (cfg["obs"] + cfg["act"] + 1) * 8 + demo_n * (cfg["obs"] + cfg["act"]) * 8 + n_glob * 4 * VF["epochs"]
    d2h = n_local * (8 + 4 + 8) + eng.d * 4 + 3 * eng.vf_d * 4

    if rank != 0:
        runtime.shutdown()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (the FVP tile kernel) ----------------
    peaks = load_peaks()
    fl = flops_per_sample_fvp(cfg) * n_local
    by = 4.0 * n_local * cfg["obs"]
    linear = len(cfg["hidden"]) == 0
    t = fvp_ms_kernel * 1e-3
    prof = {}
    pj = os.path.join(ROOT, "profiles", "fvp_ncu_%s.json" % args.config)
    if os.path.exists(pj):
        prof = json.load(open(pj))
    if linear:
        roof = {"bound": "hbm", "achieved": by / t / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": by / t / 1e9 / peaks["hbm_gbs"], "traffic": prof.get("dram_bytes_per_launch")}
    else:
        roof = {"bound": "tensor", "achieved": fl / t / 1e12, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": fl / t / 1e12 / peaks["tflops_sustained"], "traffic": prof.get("dram_bytes_per_launch")}
    sm_clock = (clk or {}).get("sm_mhz") or 1965.0
    fma_peak = 148 * 128 * 2 * sm_clock * 1e6 / 1e12
    tc_active = bool(eng.set_tensor_cores(True))
    if linear:
        kname = ("linear_tc_kernel (tcgen05 kind::f16, M=128 stacked fp16 hi/lo rows, TMEM-resident gradient accumulators)"
                 if tc_active else "linear_kernel<AG,MODE_FVP> (fp32 FMA)")
    else:
        kname = ("fvp_tc_kernel (tcgen05 kind::f16, two-term fp16 split = 3 MMAs per logical product, TMEM accumulators)"
                 if tc_active else "mlp_kernel<H,MT,MODE_FVP> (fp32 FMA)")
    roof.update({"kernel": kname,
                 "launch_ms": fvp_ms_kernel, "algorithmic_flops_per_launch": fl, "algorithmic_bytes_per_launch": by,
                 "peak_source": peaks["source"] + ("; bf16 dense sustained (kernel timed inside a long step)" if not linear else ""),
                 "hbm_frac": by / t / 1e9 / peaks["hbm_gbs"],
                 "fp32_fma_peak_tflops_at_observed_clock": fma_peak, "frac_of_fp32_fma_peak": fl / t / 1e12 / fma_peak,
                 "executed_tensor_flops_per_launch": (3 * fl if tc_active else 0),
                 "note": ("rank-0 shard; HBM-bound shape (SURVEY 8d): obs are streamed exactly once per launch; launch_ms"
                          if linear else "rank-0 shard; compute-bound shape (SURVEY 8d); launch_ms") + " is the mean CUDA-event time of the FVP launches "
                         "inside the timed steps, where the kernel shares the GPU with the concurrent baseline fit "
                         "(tensor-core fit kernel: 1 SM, FVP on 147; cluster fallback: 16 SMs); achieved counts ALGORITHMIC flops (10P-4P1 per timestep), not the 3x split MMAs"})

    # ---------------- CPU baseline (bounded sample, rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_s = min(cfg["n_traj"], max(5, 100000 // cfg["horizon"]))
        step, n = cpu_reference_step_fn(cfg, n_s)
        step()
        t0 = time.time()
        reps_cpu = 2
        for _ in range(reps_cpu):
            step()
        dt = (time.time() - t0) / reps_cpu
        scale = n_glob / n
        cpu = {"value": 1.0 / (dt * scale), "unit": "train_step/s", "cores": torch.get_num_threads(),
               "os_cpu_count": os.cpu_count(), "kind": "port",
               "sample": "%d of %d trajectories (%d timesteps), %d timed steps after 1 warm-up; extrapolated linearly x%.0f"
                         % (n_s, cfg["n_traj"], n, reps_cpu, scale),
               "fvp_per_sec": 1.0 / (cpu_fvp_time(cfg, n_s) * scale)}

    line = {"metric": "train_step_per_sec", "value": 1e3 / ms_per_step, "unit": "train_step/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, args, world), "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": 1.0 / e2e_s, "unit": "train_step/s", "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "mjrl_b200.algos.%s.update_from_paths(paths) on host float64 path dicts (per rank shard)"
                           % {"npg": "npg_cg.NPG", "trpo": "trpo.TRPO", "dapg": "dapg.DAPG"}[cfg["algo"]]},
            "fvp_per_sec": fvp_per_sec, "fvp_ms_in_cg": cg_ms / CG_ITERS, "fvp_kernel_ms": fvp_ms_kernel,
            "wall_ms_per_step": wall / args.steps * 1e3, "phase_ms": phase, "trpo_backtracks": backtracks,
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    runtime.shutdown()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    run_gpu(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
