"""LPIPS throughput (bench.py's `lpips` key on its own): batched pairs/s and TFLOP/s for both backbones at the BASELINE sizes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
print(json.dumps(bench.lpips_throughput(torch, bench.peaks()), indent=1))
