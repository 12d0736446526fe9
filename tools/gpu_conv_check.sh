#!/bin/bash
# conv parity (bounded), then per-layer timings of E2VID at batch $BATCH (+ per-CTA phase counters with TIMING=1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q 2>&1 | tail -8 | tee gpurun_out/conv_tests.log
for cs in ${CS_LIST:-0}; do
  echo "=== EVK_TC_CS=$cs" | tee -a gpurun_out/tc_exp.log
  EVK_TC_CS=$cs EVK_TC_VERBOSE=1 FRAMES=4 timeout 180 python tools/tc_experiment.py 2>&1 | sort -u | cut -c1-400 | tee -a gpurun_out/tc_exp.log
done
if [ -n "$TIMING" ]; then EVK_TC_TIMING=1 FRAMES=3 timeout 180 python tools/tc_experiment.py 2>&1 | grep TIMING | sort -u -k1,12 | cut -c1-400 | tee -a gpurun_out/tc_exp.log; fi
