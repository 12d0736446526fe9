#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_conv.py tests/test_gpu_pipeline.py tests/test_gpu_eval_loop.py tests/test_gpu_lpips.py -q 2>&1 | tail -15 | tee gpurun_out/head_tests.log
timeout 300 python tools/bench_models.py --models e2vid,firenet,hyper > gpurun_out/head_models.jsonl 2> gpurun_out/head_models.err; tail -3 gpurun_out/head_models.err
EVK_TC_WIDE_ST=0 EVK_HEAD_GROUP=1 timeout 300 python tools/bench_models.py --models e2vid > gpurun_out/head_off_e2vid.jsonl 2>> gpurun_out/head_models.err
EVK_HEAD_GROUP=1 timeout 300 python tools/bench_models.py --models e2vid > gpurun_out/head_wideonly_e2vid.jsonl 2>> gpurun_out/head_models.err
python - <<'P'
import json
for f in ('head_models','head_wideonly_e2vid','head_off_e2vid'):
    for l in open('gpurun_out/%s.jsonl'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['model'], d['frames_per_s'], d['ms_per_step'], d['forward_ms_eager'])
            for r in d['layers']: print('   ', r['ms'], r.get('tflops'), r['op'][:110])
P
