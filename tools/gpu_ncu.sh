#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + full captures of the three tensor-core kernels.
# (1 GPU; numbers printed by programs running under ncu are never bench values.)
mkdir -p gpurun_out
# the CG loop is a captured CUDA graph in production; ncu lists kernels of eager launches, so the graph is off here
export MJRL_B200_GRAPH=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-hbm-roofline > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fvp_tc_kernel -s 12 -c 1 -f -o gpurun_out/fvp_cfg3 \
    python tools/ncu_targets.py cfg3 > gpurun_out/ncu_fvp.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vf_fit_tc_kernel -s 1 -c 1 -f -o gpurun_out/fit_cfg3 \
    python tools/ncu_targets.py cfg3 > gpurun_out/ncu_fit.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_tc_tma_kernel -s 2 -c 1 -f -o gpurun_out/lin_cfg5 \
    python tools/ncu_targets.py cfg5 > gpurun_out/ncu_lin.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:returns_kernel -s 1 -c 1 -f -o gpurun_out/returns_cfg3 \
    python tools/ncu_targets.py cfg3 > gpurun_out/ncu_ret.log 2>&1
tail -2 gpurun_out/ncu_*.log
ls -la gpurun_out/*.ncu-rep
