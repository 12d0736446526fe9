#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_metrics.py tests/test_gpu_pipeline.py tests/test_gpu_eval_loop.py -q 2>&1 | grep -E "^FAILED|^ERROR|passed|failed" | cut -c1-200 | tee gpurun_out/head_tests.log
timeout 300 python tools/bench_models.py --models hyper --batch 24 > gpurun_out/x_hyper.jsonl 2>> gpurun_out/x.err
python tools/step_breakdown.py 2>&1 | tail -3
python - <<'P'
import json
for f in ('hyper',):
    for l in open('gpurun_out/x_%s.jsonl'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['model'], round(d['frames_per_s']), round(d['ms_per_step'],4), round(d['forward_ms_eager'],4))
            for r in d['layers'][12:22]: print('   ', r['ms'], r.get('tflops'), r['op'][:100])
P
tail -5 gpurun_out/x.err
