#!/bin/bash
mkdir -p gpurun_out
for b in 24 30 36 37 48 49; do
  timeout 300 python tools/bench_models.py --models e2vid --batch $b --steps 30 > gpurun_out/sweep_b$b.jsonl 2>> gpurun_out/sweep.err
done
python - <<'P'
import json
for b in (24,30,36,37,48,49):
    for l in open('gpurun_out/sweep_b%d.jsonl'%b):
        if l.startswith('{'):
            d=json.loads(l); print(b, round(d['frames_per_s']), round(d['ms_per_step'],4), round(d['forward_ms_eager'],4))
P
tail -3 gpurun_out/sweep.err
