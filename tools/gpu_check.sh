#!/bin/bash
# One gpurun call: full GPU test suite, sanitizer passes on the tcgen05 kernels, fit drift report, default bench line.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tests|san|fit|bench|prof ...]
set -u
mkdir -p gpurun_out
what="${*:-tests san fit bench}"
for w in $what; do
case $w in
lin)
  # new / risky kernel first, alone and under a short timeout: a hang must not take the whole call with it
  timeout 240 python -m pytest tests/test_engine_gpu.py -q -x -k "test_shapes_vs_oracle or (test_vpg_fvp_eval and linear) or (test_policy_steps and linear)" > gpurun_out/pytest_lin.log 2>&1
  rc=$?; echo "lin rc=$rc" >> gpurun_out/pytest_lin.log
  if [ $rc -ne 0 ]; then export MJRL_B200_LIN_TMA=0; echo "TMA linear kernel disabled for the rest of this call" >> gpurun_out/pytest_lin.log; fi ;;
tests)
  timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/pytest_gpu.log; echo "pytest rc=${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log ;;
san)
  SEL='(test_baseline_fit and pm_40x25_ragged) or test_tensor_core_fvp or (test_vpg_fvp_eval and (cheetah or linear))'
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -x -k "$SEL" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
  timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_engine_gpu.py -q -x -k "$SEL" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/sanitizer_synccheck.log ;;
fit)
  python tools/fit_fullsize_report.py > gpurun_out/fit_fullsize.log 2>&1
  python tools/vf_fit_profile.py > gpurun_out/vf_fit_profile.log 2>&1 ;;
bench)
  python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err ;;
esac
done
tail -15 gpurun_out/pytest_lin.log 2>/dev/null
tail -5 gpurun_out/pytest_gpu.log 2>/dev/null
tail -3 gpurun_out/sanitizer_*.log 2>/dev/null
cat gpurun_out/fit_fullsize.log 2>/dev/null
cut -c1-600 gpurun_out/bench_cfg3.json 2>/dev/null
