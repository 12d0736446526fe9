#!/bin/bash
# round-2 capture after the mixed-operand kernels (run under gpurun, 1 GPU): ncu launch list of the bench command, full capture
# of the 20 convolution launches of one E2VID forward (batch 36), the extra tensor-op counters of the mixed kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --fast --steps 2 --warmup 3 > gpurun_out/r02b_ncu_bench.log 2>&1
ncu --set full --metrics sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum,sm__ops_path_tensor_op_utcqmma_src_fp4_fp6_fp8_dst_fp32_sparsity_off.sum,sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum --clock-control none --import-source on -k regex:conv_tc_kernel -s 80 -c 20 -f -o gpurun_out/r02b_prof_conv python tools/profile_step.py --steps 5 --batch 36 > gpurun_out/r02b_ncu_conv.log 2>&1
ls -la gpurun_out/r02b*
