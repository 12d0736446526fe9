#!/bin/bash
# in-kernel phase counters per layer (EVK_TC_TIMING) + plans, then a full ncu capture of the 18 convolution launches of one forward
mkdir -p gpurun_out
EVK_TC_VERBOSE=1 EVK_TC_TIMING=1 FRAMES=3 BATCH=24 timeout 200 python tools/tc_experiment.py > gpurun_out/diag_timing.log 2>&1
grep "conv_tc plan" gpurun_out/diag_timing.log | sort -u | cut -c1-220
grep TIMING gpurun_out/diag_timing.log | tail -18 | cut -c1-330
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 36 -c 18 -f -o gpurun_out/r01_prof_conv_v6 python tools/profile_step.py --steps 3 > gpurun_out/ncu_conv_v6.log 2>&1
ls -la gpurun_out/*.ncu-rep
