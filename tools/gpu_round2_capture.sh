#!/bin/bash
# round-2 capture (run under gpurun, 1 GPU): ncu launch list of the bench command, full captures of the convolution family
# (E2VID forward), the HyperE2VID dynamic-decoder kernels, the voxelizer (both paths) and LPIPS; SASS opcode histogram.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --fast --steps 2 --warmup 3 > gpurun_out/r02_ncu_bench.log 2>&1
# E2VID forward = 20 conv_tc launches: the fifth forward is launches 80..99
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 80 -c 20 -f -o gpurun_out/r02_prof_conv python tools/profile_step.py --steps 5 --batch 36 > gpurun_out/r02_ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hyper_apply_u|hyper_context" -s 4 -c 2 -f -o gpurun_out/r02_prof_hyper python tools/profile_step.py --steps 4 --batch 36 --model hyper > gpurun_out/r02_ncu_hyper.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel -s 4 -c 2 -f -o gpurun_out/r02_prof_voxel python tools/profile_step.py --voxel-only --steps 4 > gpurun_out/r02_ncu_voxel.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel -s 4 -c 1 -f -o gpurun_out/r02_prof_voxel_small python tools/profile_step.py --voxel-only --voxel-events 400000 --steps 6 > gpurun_out/r02_ncu_voxel_small.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/r02_launches_lpips_vgg.csv python tools/profile_step.py --lpips lpips-vgg --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 32 -c 16 --csv --log-file gpurun_out/r02_launches_lpips_alex.csv python tools/profile_step.py --lpips lpips --steps 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/r02_launches_firenet.csv python tools/profile_step.py --steps 3 --batch 36 --model firenet > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
