#!/bin/bash
# phase-stacked decoders: parity tests, then the E2VID model bench with both / last only / none
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_networks.py tests/test_gpu_conv.py tests/test_gpu_pipeline.py tests/test_gpu_eval_loop.py -q 2>&1 | tail -15 | tee gpurun_out/poly_tests.log
timeout 300 python tools/bench_models.py --models e2vid > gpurun_out/poly_e2vid.jsonl 2> gpurun_out/poly_e2vid.err; tail -3 gpurun_out/poly_e2vid.err
EVK_POLY_MAX_C=64 timeout 300 python tools/bench_models.py --models e2vid > gpurun_out/poly64_e2vid.jsonl 2>> gpurun_out/poly_e2vid.err
EVK_NO_POLY=1 timeout 300 python tools/bench_models.py --models e2vid > gpurun_out/nopoly_e2vid.jsonl 2>> gpurun_out/poly_e2vid.err
python - <<'P'
import json
for f in ('poly_e2vid','poly64_e2vid','nopoly_e2vid'):
    for l in open('gpurun_out/%s.jsonl'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['frames_per_s'], d['ms_per_step'], d['forward_ms_eager'])
            for r in d['layers'][-10:]: print('   ', r['ms'], r.get('tflops'), r['op'])
P
