#!/bin/bash
# round capture: GPU tests, both bench arms, model benches, the ncu launch list of the bench command, full captures of the conv / voxelizer kernels, memcheck of the phase-stacked decoder tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | tee gpurun_out/r01_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r01_bench.json 2> gpurun_out/r01_bench.err; tail -2 gpurun_out/r01_bench.err
timeout 600 python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r01_bench_reference.json 2> gpurun_out/r01_bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r01_smoke.log
timeout 600 python tools/bench_models.py --batch 36 > gpurun_out/r01_models.jsonl 2> gpurun_out/r01_models.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
# E2VID forward = 20 conv_tc launches: the fifth forward is launches 80..99
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 80 -c 20 -f -o gpurun_out/r01_prof_conv python tools/profile_step.py --steps 5 --batch 36 > gpurun_out/ncu_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voxel -s 4 -c 2 -f -o gpurun_out/r01_prof_voxel python tools/profile_step.py --voxel-only --steps 4 > gpurun_out/ncu_voxel.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_networks.py -q -k "base32 or two_encoders or three_encoders or firenet_real" > gpurun_out/r01_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r01_memcheck.log; tail -4 gpurun_out/r01_memcheck.log
ls -la gpurun_out/*.ncu-rep
python -c "
import json; d=json.load(open('gpurun_out/r01_bench.json')); print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}); print(d['e2e']); print(d['roofline']['achieved'], d['roofline']['frac']); print(d['cpu_baseline']); print(d['voxelizer']['roofline'])
print(open('gpurun_out/r01_bench_reference.json').read()[:300])"
