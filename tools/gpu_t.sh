#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gpu_networks.py tests/test_gpu_conv.py tests/test_gpu_pipeline.py tests/test_gpu_eval_loop.py tests/test_gpu_lpips.py"
echo "=== default"; timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | cut -c1-150 | tee gpurun_out/t_default.log
echo "=== PDL=0"; EVK_TC_PDL=0 timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | cut -c1-150 | tee gpurun_out/t_nopdl.log
echo "=== FASTLIN=0"; EVK_TC_FASTLIN=0 timeout 600 python -m pytest $T -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | cut -c1-150 | tee gpurun_out/t_nofast.log
echo "=== first failure"; timeout 600 python -m pytest $T -q -x 2>&1 | grep -E "Error|error|assert" | head -20 | cut -c1-300
