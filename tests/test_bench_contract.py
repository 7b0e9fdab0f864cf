"""The bench.py JSON contract, checked on the CPU through the reference arm (`--impl reference` times the oracle's
torch flavour on the host cores; the CUDA arm prints the same keys plus roofline / clocks / gpu_launches)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "train_step_per_sec" and d["unit"] == "train_step/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["value"] * d["ms_per_step"] - 1e3) < 1e-3 * 1e3
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_committed_round2_lines_follow_the_contract():
    """The bench lines kept under profiles/ (produced on the GPU box) carry every key of the contract, and their derived
    numbers are consistent: value = 1000 / ms_per_step, roofline.frac = achieved / peak, e2e bytes come from the counters."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_cfg*_1gpu.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_scale_cfg3_*gpu.json")))
    assert len(files) >= 8
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "e2e", "roofline"):
            assert k in d, (f, k)
        assert d["metric"] == "train_step_per_sec" and d["higher_is_better"] is True and d["warmup"] >= 3
        assert abs(d["value"] * d["ms_per_step"] - 1e3) < 1.0
        assert d["gpu_launches"] > 0 and d["data"] == "synthetic" and "workload" in d["config"]
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= d["value"] * 1.02
        assert e["uploads_in_timed_region"] >= e["steps"]
        c = d["clocks"]
        assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if c["samples"]:                       # (cfg1's 6 ms region ended before nvidia-smi's first sample in that run)
            assert c["sm_mhz"] >= 0.9 * c["sm_max_mhz"]
        if d["n_gpus"] == 1:
            assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] > 0
    d = json.loads(open(os.path.join(ROOT, "profiles", "r02_reference_arm_cfg3.json")).read())
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and "FULL batch" in d["cpu_baseline"]["sample"]
