#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_eval_loop.py tests/test_gpu_pipeline.py tests/test_gpu_conv.py -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | cut -c1-250 | head -12 | tee gpurun_out/head_tests.log
run() { tag=$1; shift; env "$@" timeout 300 python tools/bench_models.py --models e2vid --batch 36 > gpurun_out/x_$tag.jsonl 2>> gpurun_out/x.err; }
run base EVK_X=0
run nopair EVK_NO_PIXEL_PAIR=1
python - <<'P'
import json
for f in ('base','nopair'):
    for l in open('gpurun_out/x_%s.jsonl'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['model'], round(d['frames_per_s']), round(d['ms_per_step'],4), round(d['forward_ms_eager'],4), ' '.join('%.3f'%r['ms'] for r in d['layers']))
            print(d['layers'][2]['op'])
P
tail -5 gpurun_out/x.err
