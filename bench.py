#!/usr/bin/env python
"""Benchmark of the B200-native EVREAL hot path (BASELINE.json: frames/s reconstructed + events/s voxelized).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (N=1 and per rank for N>1, weak scaling): BASELINE.json configs[1] -- E2VID (BN, sigmoid, base 32; seeded
random weights of the shipped checkpoint's shapes) on synthetic ECD-shape streams: 240x180, 1 Mev/s, 24 Hz frames,
5 bins, 'between_frames' windows (~41.7k events), normalize_event_tensor on, pad to 184x240, 'robust' percentile
post-normalisation, clip, MSE + SSIM per frame.  One STEP = frame i of B independent streams run in lock-step
(SequenceBatch): one batched voxelizer launch + one batched network forward + one batched metric launch, the voxelizer of
step i+1 and the metrics of step i-1 on side streams around the network of step i.  `value` = frames/s summed over all
ranks with the raw event arrays resident in HBM; `e2e` = the same loop with the event arrays and reference frames in
pinned HOST memory, every window copied host->device and scores + reconstructed frames copied device->host inside the
timed region.  Both timed regions end after every side stream (copies included) has joined the timing stream.

Other keys of the line (each one its own small measurement, outside the two timed regions above):
  roofline        per-launch CUDA-event timing of the convolution family vs the measured sustained bf16 peak
  voxelizer       BASELINE cfg 5 sweep (640x480, 40k .. 4M events/window) vs the measured HBM bandwidth
  single_stream   the same E2VID loop at B = 1 (cfg 2 is literally one stream; the reference is batch 1)
  gpu_reference   the reference network arithmetic as plain torch modules on this GPU through cuDNN (TF32 off / on,
                  batch 1 and B): the "torch GPU path on the same box" bar of SURVEY 8(d)
  parity          step 1 of one stream through the timed pipeline vs the CPU oracle (outside the timed region)
  cfg3            BASELINE configs[2]: FireNet on 16 HQF-shape sequences SHARDED over the N ranks through the plugin
                  surface (evaluate(): config/*.json + checkpoint, lock-step batches, ONE NCCL all-reduce of the metric
                  sums) -- wall clock of the whole job including model construction, sequence open / upload and the
                  all-reduce, max over ranks; strong scaling
  cfg4            BASELINE configs[3]: HyperE2VID + MSE/SSIM/LPIPS on 8 MVSEC-shape sequences (346x260, 5 Mev/s), same form
  lpips           LPIPS (AlexNet = pyiqa 'lpips', VGG16 = 'lpips-vgg') batched throughput and TFLOP/s (seeded weights)
  cpu_baseline    the reference's per-frame loop on the host cores (oracle port), N = 1 only

--impl reference: the reference's torch-CPU path for the same config (oracle/eval_loop.py: the same ATen/oneDNN
operators eval.py executes per frame, all host threads), one stream, one frame per step.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reconstructed_frames_per_s"
UNIT = "frames/s"
H, W, RATE, DURATION, FPS = 180, 240, 1.0e6, 10.0, 24.0       # ECD-shape (SURVEY 8d cfg 2)
WORKLOAD = "E2VID on synthetic ECD-shape streams (240x180, 1 Mev/s, 24 Hz, 5 bins, normalize+pad 184x240, robust norm, MSE+SSIM)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled every 5 ms DURING a timed region, in-process through NVML (an
    `nvidia-smi -lms` child needs longer to start than a 20-step timed region lasts: round 1's SCALE lines had 0 samples)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4),
               ("hw_power_brake", 0x80))

    def __init__(self, device_index):
        self.samples, self.reasons, self.smax, self.power = [], set(), None, []
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:
            self.error = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.smax, "reasons": [], "samples": 0}
        if self._thread is None:
            out["error"] = getattr(self, 'error', 'NVML unavailable')
            return out
        self._stop.set()
        self._thread.join(timeout=2)
        if self.samples:
            s = sorted(self.samples)
            out.update(sm_mhz=s[len(s) // 2], reasons=sorted(self.reasons), samples=len(s),
                       power_w_max=max(self.power) if self.power else None)
        return out


def make_streams(n_streams, seed0, duration, shape=(H, W, RATE, FPS)):
    from evreal_b200 import synthetic
    from evreal_b200.dataset import MemMapDataset
    from concurrent.futures import ThreadPoolExecutor
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workers = max(1, min(n_streams, (os.cpu_count() or 1) // max(world, 1)))     # numpy's generators and sort release the GIL
    h, w, rate, fps = shape
    with ThreadPoolExecutor(max_workers=workers) as pool:
        arrs = list(pool.map(lambda b: synthetic.make_stream(h, w, rate, duration, fps, seed=seed0 + b), range(n_streams)))
    return [(a, MemMapDataset(a, num_bins=5, voxel_method={'method': 'between_frames'}, resident=False)) for a in arrs]


def e2vid_weights():
    from evreal_b200 import synthetic
    return synthetic.unet_state_dict(seed=0, norm_bn=True)


# ----------------------------------------------------------------------------------------------- reference arm
def cpu_reference_frames_per_s(arrays, warmup, steps, threads):
    """oracle/eval_loop.run_sequence == the reference's per-frame loop on torch-CPU (save_images off)."""
    import torch
    from oracle import eval_loop, networks as on
    torch.set_num_threads(threads)
    w = {k[len('unetrecurrent.'):]: v for k, v in e2vid_weights().items()}
    model = on.UNetRecurrentOracle(w, final_sigmoid=True)
    # item 0 of 'between_frames' is always an empty window: start the timed sample after the warm-up items
    res = eval_loop.run_sequence(arrays, (H, W), model, 3, True, 'robust', max_items=1 + warmup)
    tm = {}
    t0 = time.perf_counter()
    sub = dict(arrays)
    # timed sample: `steps` frames following the warm-up frames (windows are contiguous slices of the same stream)
    lo = 1 + warmup
    sub['images'] = arrays['images'][lo - 1:]
    sub['images_ts'] = arrays['images_ts'][lo - 1:]
    sub['image_event_indices'] = arrays['image_event_indices'][lo - 1:]
    res = eval_loop.run_sequence(sub, (H, W), model, 3, True, 'robust', max_items=1 + steps, timers=tm)
    dt = time.perf_counter() - t0
    frames = res['frames'] - 1            # the first item of the sub-sequence is the empty window again
    return frames / dt, res['events'] / dt, dt, tm


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    steps = args.steps
    arrays, _ = make_streams(1, 0, min(DURATION, (args.warmup + steps + 4) / FPS))[0]
    fps, evps, dt, tm = cpu_reference_frames_per_s(arrays, args.warmup, steps, threads)
    sample = "%d frames of one ECD-shape stream after %d warm-up frames (%.1f s of CPU time)" % (steps, args.warmup, dt)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "events_per_s": evps,
            "config": {"workload": WORKLOAD, "batch_streams_per_gpu": 1, "frames_per_step": 1},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "stage_seconds": tm},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------- our arm
def voxelizer_roofline(pk, torch, _lib):
    """cfg 5: 640x480, 5 bins, 40k .. 4M events/window (1 .. 100 Mev/s at 25 windows/s), f32 SoA resident in HBM.
    Algorithmic bytes = 16*N + 4*bins*H*W.  Four event sets (256 MB > 126 MB L2) are rotated so no
    iteration re-reads L2-resident inputs.  Plus the sizes the pipeline uses: one launch for the B windows of a lock-step
    step at 240x180 (cfg 2 / 3) and 346x260 (cfg 4), raw int16/f64/u8 events (13 B/event)."""
    lib = _lib.load()
    Hv, Wv, n, bins, sets = 480, 640, 4_000_000, 5, 4
    g = torch.Generator(device='cuda').manual_seed(1)
    evs = []
    for s in range(sets):
        x = torch.randint(0, Wv, (n,), device='cuda', generator=g).float()
        y = torch.randint(0, Hv, (n,), device='cuda', generator=g).float()
        t = torch.sort(torch.rand(n, device='cuda', generator=g) * 0.04)[0]
        t = t - t[0]
        p = torch.randint(0, 2, (n,), device='cuda', generator=g).float() * 2 - 1
        evs.append((x, y, t, p))
    grid = torch.empty((bins, Hv, Wv), dtype=torch.float32, device='cuda')
    st = _lib.stream_ptr()

    def timeit(run, iters):
        for i in range(4):
            run(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    def measure(n_ev, iters):
        def run(i):
            x, y, t, p = evs[i % sets]
            _lib.check(lib.evk_voxelize(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), n_ev, bins, Hv, Wv, _lib.ptr(grid), None, st))
        ms = timeit(run, iters)
        bytes_alg = 16.0 * n_ev + 4.0 * bins * Hv * Wv
        return ms, bytes_alg, bytes_alg / (ms * 1e-3) / 1e9

    # cfg 5 sweep: 1..100 Mev/s at 25 windows/s -> 40k..4M events per window (a prefix of the sorted event sets)
    sweep = []
    for rate, n_ev in ((1, 40_000), (2, 80_000), (5, 200_000), (10, 400_000), (20, 800_000), (50, 2_000_000), (100, 4_000_000)):
        ms_i, _, gbs_i = measure(n_ev, 20)
        sweep.append({"Mev_per_s_stream": rate, "events_per_window": n_ev, "us_per_window": ms_i * 1e3,
                      "events_per_s": n_ev / (ms_i * 1e-3), "GB_per_s": gbs_i, "frac": gbs_i / pk["hbm_gbs"]})
    ms, bytes_alg, gbs = measure(n, 20)
    # cfg 5's second distribution: "edge-clustered" -- events on a few moving edges, the best case for merging the
    # contributions of a warp's lanes before they reach L2.  `distinct_cells_per_warp` counts what such a merge could save:
    # the distinct (pixel, bin group) reduction targets among 32 consecutive events.
    uniform_sets = evs
    def distinct_per_warp(x, y, t):
        key = ((y.long() * Wv + x.long()) * 2 + (t / t[-1] * (bins - 1)).floor().clamp_(0, bins - 1).long() // 3)[: (n // 32) * 32].view(-1, 32)
        srt = torch.sort(key, dim=1)[0]
        return float(((srt[:, 1:] != srt[:, :-1]).sum(1) + 1).float().mean())
    evs = []
    for s in range(sets):
        k_edges = 24
        e = torch.randint(0, k_edges, (n,), device='cuda', generator=g)
        t = torch.sort(torch.rand(n, device='cuda', generator=g) * 0.04)[0]
        t = t - t[0]
        ex0 = torch.rand(k_edges, device='cuda', generator=g) * Wv
        ey0 = torch.rand(k_edges, device='cuda', generator=g) * Hv
        ang = torch.rand(k_edges, device='cuda', generator=g) * 3.14159
        length = 60.0 + torch.rand(k_edges, device='cuda', generator=g) * 200.0
        vx = (torch.rand(k_edges, device='cuda', generator=g) - 0.5) * 4000.0            # pixels per second
        vy = (torch.rand(k_edges, device='cuda', generator=g) - 0.5) * 4000.0
        along = (torch.rand(n, device='cuda', generator=g) - 0.5) * length[e]
        jitter = torch.randn(n, device='cuda', generator=g) * 0.7
        x = (ex0[e] + vx[e] * t + along * torch.cos(ang[e]) - jitter * torch.sin(ang[e])).clamp_(0, Wv - 1).floor()
        y = (ey0[e] + vy[e] * t + along * torch.sin(ang[e]) + jitter * torch.cos(ang[e])).clamp_(0, Hv - 1).floor()
        p = torch.randint(0, 2, (n,), device='cuda', generator=g).float() * 2 - 1
        evs.append((x, y, t, p))
    clustered = []
    for rate, n_ev in ((10, 400_000), (100, 4_000_000)):
        ms_i, _, gbs_i = measure(n_ev, 20)
        clustered.append({"Mev_per_s_stream": rate, "events_per_window": n_ev, "us_per_window": ms_i * 1e3,
                          "events_per_s": n_ev / (ms_i * 1e-3), "GB_per_s": gbs_i, "frac": gbs_i / pk["hbm_gbs"]})
    aggregation = {"distinct_cells_per_warp_uniform": distinct_per_warp(*uniform_sets[0][:3]),
                   "distinct_cells_per_warp_edge_clustered": distinct_per_warp(*evs[0][:3]),
                   "what": "distinct (pixel, bin-group) reduction targets among 32 consecutive events of a 4 M-event window: 32 = nothing to merge"}
    del uniform_sets, evs
    # the launches the pipeline issues: B raw windows per call
    import ctypes
    batched = []
    for tag, (hb, wb, n_win, B) in {"cfg2/3 step: 36 windows of 41.7k events at 240x180": (180, 240, 41_700, 36),
                                    "cfg4 step: 24 windows of 111k events at 346x260": (260, 346, 111_000, 24)}.items():
        tot = n_win * B * 3                                   # three rotating sets
        xy = torch.stack([torch.randint(0, wb, (tot,), device='cuda', generator=g), torch.randint(0, hb, (tot,), device='cuda', generator=g)], 1).to(torch.int16)
        tt = torch.sort(torch.rand(tot, device='cuda', generator=g, dtype=torch.float64))[0]
        pp = torch.randint(0, 2, (tot,), device='cuda', generator=g).to(torch.uint8)
        grids = torch.empty((B, bins, hb, wb), dtype=torch.float32, device='cuda')
        oob = torch.zeros(1, dtype=torch.int32, device='cuda')
        wins = [(_lib.EventWindow * B)() for _ in range(3)]
        for s3 in range(3):
            for b in range(B):
                o = (s3 * B + b) * n_win
                wins[s3][b].xy = xy.data_ptr() + o * 4
                wins[s3][b].t = tt.data_ptr() + o * 8
                wins[s3][b].pol = pp.data_ptr() + o
                wins[s3][b].n = n_win
        ms_b = timeit(lambda i: _lib.check(lib.evk_voxelize_raw_batch(wins[i % 3], B, bins, hb, wb, _lib.ptr(grids), _lib.ptr(oob), st)), 30)
        alg = 13.0 * n_win * B + 4.0 * bins * hb * wb * B
        batched.append({"workload": tag, "us_per_launch_group": ms_b * 1e3, "events_per_s": n_win * B / (ms_b * 1e-3),
                        "algorithmic_bytes": alg, "GB_per_s": alg / (ms_b * 1e-3) / 1e9, "frac": alg / (ms_b * 1e-3) / 1e9 / pk["hbm_gbs"]})
        del xy, tt, pp, grids
    return {"workload": "voxelizer only, 640x480, 5 bins, 4M events/window (cfg 5 top point), f32 SoA in HBM, 4 rotating event sets (256 MB > L2)",
            "events_per_s": n / (ms * 1e-3), "ms_per_window": ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes_per_launch": bytes_alg, "peak_source": pk["source"],
                         "note": "one RED.ADD.V4.F32 per event into an L2-resident interleaved grid: bound by the L2 reduction request rate (~83/clk), not by HBM"},
            "sweep_cfg5": sweep, "sweep_cfg5_edge_clustered": clustered, "warp_aggregation": aggregation, "pipeline_sizes": batched}


def network_roofline(model, padded, pk, frames=6):
    """Per-launch CUDA-event timing of the network (evk_model_profile) -> the dominant kernel family's achieved
    TFLOP/s on algorithmic FLOPs (no credit for split-precision passes)."""
    agg = {}
    for f in range(frames):
        rows = model.profile_forward(padded)
        if f < 2:
            continue
        for i, (desc, ms, fl) in enumerate(rows):
            a = agg.setdefault(i, [desc, 0.0, fl, 0])
            a[1] += ms
            a[3] += 1
    layers = [{"op": a[0], "ms": a[1] / a[3], "gflop": a[2] / 1e9,
               "tflops": (a[2] / (a[1] / a[3] * 1e-3) / 1e12) if a[1] > 0 else 0.0} for _, a in sorted(agg.items())]
    total_ms = sum(l["ms"] for l in layers)
    conv = [l for l in layers if l["op"].startswith("conv")]
    conv_ms = sum(l["ms"] for l in conv)
    conv_fl = sum(l["gflop"] for l in conv) * 1e9
    ach = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get("conv_dram_bytes_per_forward")
            traffic_src = "profiles/traffic.json (ncu --set full capture of %s, not measured by this run)" % tj.get("capture", "an earlier run")
        except Exception:
            traffic = None
    return {"bound": "tensor", "achieved": ach, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
            "frac": ach / pk["tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
            "kernel": "implicit-GEMM convolution family (all conv launches of one forward)",
            "launches_per_forward": len(conv), "algorithmic_flops_per_forward": conv_fl,
            "mixed_operand_launches": sum("f16+2xf8" in l["op"] for l in conv),
            "operands": "fp32 semantics from split operands: x.w = x16.w16 + x8.wl8 + xl8.w8 (one fp16 product + two fp8 products at twice "
                        "the rate = 2 tensor-pipe units per algorithmic FLOP) in the layers marked f16+2xf8, three bf16 products (3 units) in "
                        "the rest: the ceiling of `frac` is between 1/3 and 1/2",
            "share_of_forward_time": conv_ms / total_ms if total_ms > 0 else None,
            "forward_ms_eager": total_ms, "peak_source": pk["source"] + ", sustained bf16", "layers": layers}


def torch_gpu_reference(torch, B, iters=8):
    """The reference network arithmetic as plain torch modules ON THIS GPU (cuDNN convolutions): oracle/networks.py with its
    weights moved to the device -- the same functional graph eval.py's model(voxel) executes through ATen -- at batch 1
    (how eval.py runs) and batch B (the batched form a user could write), TF32 off (the parity setting) and on
    (torch's GPU default for convolutions).  CUDA-event timed, network forward only.  A baseline leg like cpu_baseline:
    nothing here is on the product path."""
    from oracle import networks as on
    out = {}
    w = {k[len('unetrecurrent.'):]: v.cuda() for k, v in e2vid_weights().items()}
    g = torch.Generator(device='cuda').manual_seed(3)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for tf32 in (False, True):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for nb in (1, B):
                model = on.UNetRecurrentOracle(w, final_sigmoid=True)
                x = torch.randn((nb, 5, 184, 240), device='cuda', generator=g)
                model.reset_states()
                for _ in range(3):
                    model(x)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(iters):
                    model(x)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                out["tf32_%s_batch_%d" % ("on" if tf32 else "off", nb)] = {"ms_per_forward": ms, "frames_per_s": nb / (ms * 1e-3)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    out["what"] = "E2VID forward only (no voxelizer / metrics), reference graph on torch-CUDA + cuDNN, same seeded weights"
    return out


def parity_check(torch, model_factory, arrays, ds):
    """Step 1 of one stream through the SAME pipeline class the timed loop uses, against the CPU oracle of the reference's
    per-frame loop (outside every timed region).  Returns {"parity_ok": bool, ...}."""
    import numpy as np
    from evreal_b200.pipeline import SequenceBatch
    from oracle import eval_loop, networks as on
    w = {k[len('unetrecurrent.'):]: v for k, v in e2vid_weights().items()}
    a = dict(arrays)
    ts = np.asarray(a['images_ts'], dtype=np.float64).copy()
    ts = ts.reshape(-1)
    start = float(ts[1])
    ts[0] = ts[1] - 15.0                       # item 0 (always empty) is skipped by the 10-second rule: both sides start at item 1
    a['images_ts'] = ts.reshape(-1, 1)
    ref = eval_loop.run_sequence(a, (H, W), on.UNetRecurrentOracle(w, final_sigmoid=True), 3, True, 'robust', start_time_s=start, max_items=3)
    batch = SequenceBatch(model_factory(), [ds], True, 'robust', resident=True, log_scores=True)
    batch.reset()
    worst = 0.0
    ok = True
    for k in (1, 2):
        _, img, _ = batch.step(k)
        got = np.clip(img.cpu().numpy()[0, 0], 0.0, 1.0)
        want = ref['images'][k - 1]
        worst = max(worst, float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-6)))
    batch.finish()
    sc = batch.scores_log.cpu().numpy()[1:3, 0]
    for k in range(2):
        ok &= abs(sc[k, 0] - ref['mse'][k]) <= 1e-4 * ref['mse'][k] and abs(sc[k, 1] - ref['ssim'][k]) <= 1e-4 * abs(ref['ssim'][k]) + 1e-7
    ok &= worst <= 1e-4
    return {"parity_ok": bool(ok), "frames_checked": 2, "max_rel_frame_diff": worst, "tolerance": 1e-4,
            "scores_gpu_mse_ssim": [[float(v) for v in r[:2]] for r in sc], "scores_oracle_mse_ssim": [[ref['mse'][k], ref['ssim'][k]] for k in range(2)]}


def seeded_lpips_weights(net):
    """LPIPS weights do not exist offline (pyiqa downloads them): seeded random weights of the published shapes, in the
    lpips / pyiqa state_dict names.  Throughput is weight-independent; scores are NOT comparable with pyiqa's."""
    from evreal_b200 import lpips as lp
    from oracle import metrics as om
    return lp.state_dict_from_conv_list(om.random_lpips_weights(net, seed=0), 0 if net == 'alex' else 1)


def lpips_throughput(torch, pk):
    """Batched LPIPS on the convolution kernels: pairs/s and TFLOP/s on algorithmic FLOPs, both backbones, both sizes."""
    from evreal_b200.lpips import LpipsNet
    out = []
    g = torch.Generator(device='cuda').manual_seed(5)
    for net, name in (('alex', 'lpips'), ('vgg', 'lpips-vgg')):
        for (hh, ww, B) in ((180, 240, 36), (260, 346, 24)):
            ln = LpipsNet(name, seeded_lpips_weights(net), hh, ww, batch=B)
            a = torch.rand((B, hh, ww), device='cuda', generator=g)
            b = torch.rand((B, hh, ww), device='cuda', generator=g)
            for _ in range(3):
                ln(a, b)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            iters = 5
            for _ in range(iters):
                ln(a, b)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            fl = float(ln.lib.evk_lpips_flops(ln.handle))
            out.append({"backbone": name, "H": hh, "W": ww, "pairs_per_call": B, "ms_per_call": ms, "pairs_per_s": B / (ms * 1e-3),
                        "gflop_per_pair": fl / B / 1e9, "tflops": fl / (ms * 1e-3) / 1e12, "frac_of_sustained_bf16": fl / (ms * 1e-3) / 1e12 / pk["tflops_sustained"],
                        "tensor_core_layers": ln.num_tensor_core_layers})
            del ln
            torch.cuda.empty_cache()
    return out


def write_plugin_tree(root, tag, rank, world, barrier):
    """config/{method,eval,dataset}/*.json + checkpoint + sequences on disk for cfg 3 / cfg 4, written cooperatively by the
    ranks (sequence i by rank i % world).  Returns (method, dataset, metrics, frames per sequence)."""
    import torch
    from evreal_b200 import synthetic
    spec = {
        'cfg3': dict(method='FireNet', dataset='HQF16', n_seq=16, shape='hqf', metrics=['mse', 'ssim'], norm=True, post='none'),
        'cfg4': dict(method='HyperE2VID', dataset='MVSEC8', n_seq=8, shape='mvsec', metrics=['mse', 'ssim', 'lpips'], norm=False, post='none'),
    }[tag]
    hh, ww, rate, dur, fps = synthetic.SHAPES[spec['shape']]
    dur = spec.get('seconds', dur)
    if rank == 0:
        for d in ('config/method', 'config/eval', 'config/dataset', 'pretrained/' + spec['method'], 'data/' + spec['dataset']):
            os.makedirs(os.path.join(root, d), exist_ok=True)
        ck = os.path.join(root, 'pretrained', spec['method'], 'model.pth')
        if tag == 'cfg3':
            torch.save({'config': {'model': dict(synthetic.FIRENET_KWARGS)}, 'state_dict': synthetic.firenet_state_dict(0)}, ck)
        else:
            from evreal_b200 import parse_config
            torch.save({'config': parse_config.ConfigParser({'arch': {'type': 'E2VIDRecurrent', 'args': {'unet_kwargs': dict(synthetic.HYPER_KWARGS)}}}),
                        'state_dict': synthetic.unet_state_dict(0, dynamic_decoder=True)}, ck)
        json.dump({'model_name': spec['method'], 'model_path': ck, 'event_tensor_normalization': spec['norm'], 'post_process_norm': spec['post']},
                  open(os.path.join(root, 'config/method', spec['method'] + '.json'), 'w'))
        json.dump({'save_images': False, 'histeq': 'none', 'eval_infer_all': False, 'ts_tol_ms': 1.0, 'create_video': False,
                   'dataset_kwargs': {'num_bins': 5, 'voxel_method': {'method': 'between_frames'}}},
                  open(os.path.join(root, 'config/eval/std.json'), 'w'))
        seqs = {'seq%02d' % i: {'start_time_s': 0.0, 'end_time_s': dur} for i in range(spec['n_seq'])}
        json.dump({'root_path': os.path.join(root, 'data', spec['dataset']), 'sequences': seqs},
                  open(os.path.join(root, 'config/dataset', spec['dataset'] + '.json'), 'w'))
    barrier()
    from concurrent.futures import ThreadPoolExecutor
    mine = [i for i in range(spec['n_seq']) if i % world == rank]

    def write(i):
        r = synthetic.hqf_rate(i) if tag == 'cfg3' else rate
        synthetic.write_sequence(os.path.join(root, 'data', spec['dataset'], 'seq%02d' % i), hh, ww, r, dur, fps, seed=i)
    with ThreadPoolExecutor(max_workers=max(1, min(len(mine), (os.cpu_count() or 1) // max(world, 1)))) as pool:
        list(pool.map(write, mine))
    barrier()
    return spec, dur


def run_plugin_config(tag, torch, dist, rank, world, args, seconds=None):
    """BASELINE cfg 3 / cfg 4 through evaluate(): returns the dict printed under the key `tag` (rank 0), else None."""
    from evreal_b200 import evaluate as ev

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    root = os.path.join(tempfile.gettempdir(), "evk_bench_%s_%s" % (os.environ.get("MASTER_PORT", str(os.getppid() if world > 1 else os.getpid())), tag))
    if rank == 0 and os.path.exists(root):
        shutil.rmtree(root, ignore_errors=True)
    barrier()
    try:
        from evreal_b200 import synthetic
        if seconds is not None:
            key = 'hqf' if tag == 'cfg3' else 'mvsec'
            hh, ww, rate, _, fps = synthetic.SHAPES[key]
            synthetic.SHAPES[key] = (hh, ww, rate, float(seconds), fps)
        spec, dur = write_plugin_tree(root, tag, rank, world, barrier)
        lpw = seeded_lpips_weights('alex') if 'lpips' in spec['metrics'] else None
        kw = dict(config_root=os.path.join(root, 'config'), write_files=False, rank=rank, world_size=world,
                  lockstep=spec['n_seq'], lpips_weights=lpw)
        # warm-up pass: page cache, kernel plans, cuDNN-free -- and the numbers a second run of the same job sees
        ev.evaluate([spec['method']], ['std'], [spec['dataset']], spec['metrics'], **kw)
        import gc
        gc.collect()                     # the warm-up pass's model / LPIPS handles are freed here, not inside the timed call
        barrier()
        t0 = time.perf_counter()
        res = ev.evaluate([spec['method']], ['std'], [spec['dataset']], spec['metrics'], **kw)      # (ends with the all-reduce)
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        barrier()
        phases = dict(ev.last_timings)
        keys = sorted(phases)
        v = torch.tensor([dt_local] + [float(phases[k]) for k in keys], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        vv = [float(x) for x in v.cpu()]
        tot = torch.tensor([float(phases.get('frames', 0)), float(phases.get('events', 0))], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        frames, events = float(tot[0].item()), float(tot[1].item())
        tr = res['std'][spec['method']][spec['dataset']]
        means = {k: repr(tr.get_average(k)) for k in spec['metrics']}
        out = {"workload": {"cfg3": "FireNet on 16 synthetic HQF-shape sequences (240x180, %.0f s each, U(0.5,2) Mev/s, 25 Hz), MSE+SSIM" % dur,
                            "cfg4": "HyperE2VID on 8 synthetic MVSEC-shape sequences (346x260, %.0f s each, 5 Mev/s, 45 Hz), MSE+SSIM+LPIPS (AlexNet, seeded LPIPS weights)" % dur}[tag],
               "scaling": "strong (a fixed set of sequences sharded over the ranks by longest-processing-time)",
               "n_gpus": world, "sequences": spec['n_seq'], "lockstep_batch_per_rank": (spec['n_seq'] + world - 1) // world,
               "frames_reconstructed": frames, "events_voxelized": events, "frames_scored": tr.get_count('mse'),
               "seconds_wall_max_over_ranks": vv[0], "frames_per_s": frames / vv[0], "events_per_s": events / vv[0],
               "phase_seconds_max_over_ranks": {k: x for k, x in zip(keys, vv[1:]) if k.endswith('_s')},
               "loop_frames_per_s": frames / max(vv[1 + keys.index('loop_s')], 1e-9),
               "failures": int(max(vv[1 + keys.index('failures')], 0)), "means": means,
               "includes": "model construction from the checkpoint, sequence open + upload to HBM, per-frame loop, ONE all-reduce (NCCL) of [sum(score*n), sum(n)] per metric",
               "api": "evreal_b200.evaluate.evaluate(config/*.json + checkpoint, lockstep=B)"}
        return out if rank == 0 else None
    finally:
        barrier()
        if rank == 0:
            shutil.rmtree(root, ignore_errors=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU implementation (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # keep stdout to the ONE JSON line: NCCL prints its version banner to stdout (file descriptor 1, from C) whatever
    # NCCL_DEBUG says, so everything until the final print goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device('cuda', local))
    import evreal_b200 as evk
    from evreal_b200 import _lib, synthetic
    from evreal_b200.pipeline import SequenceBatch
    pk = peaks()
    B, K, Wm = args.batch, args.steps, max(args.warmup, 3)
    n_items = int(DURATION * FPS) - 1
    streams = make_streams(B, rank * B, DURATION)
    datasets = [ds for _, ds in streams]
    make_model = lambda: evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(e2vid_weights()).to(torch.device('cuda', local))
    model = make_model()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch, k_steps, warm, all_ranks=True):
        batch.reset()
        batch.wait_uploaded()                     # `value`: every input byte is resident in HBM before anything is timed
        idx = 1                                   # item 0 of 'between_frames' is always the empty window
        nn = len(batch)
        for _ in range(warm):
            nxt = idx % (nn - 1) + 1
            batch.step(idx, nxt, sync=False)
            idx = nxt
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if all_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        sampler = ClockSampler(local)
        launches = events = h2d = d2h = 0
        e0.record()
        for _ in range(k_steps):
            nxt = idx % (nn - 1) + 1
            _, _, n_ev = batch.step(idx, nxt, sync=False)
            launches += batch.launches
            events += n_ev
            h2d += batch.h2d_bytes
            d2h += batch.d2h_bytes
            idx = nxt
        batch.join()                              # every side stream (pre / post stages, H2D, D2H) joins before the stop event
        e1.record()
        if all_ranks:
            barrier()
        else:
            torch.cuda.synchronize()
        clocks = sampler.stop()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        tot = torch.tensor([float(events), float(launches)], dtype=torch.float64, device='cuda')
        if world > 1 and all_ranks:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        batch.finish()
        batch.check_bounds()
        return float(ms.item()), float(tot[0].item()), int(tot[1].item()), h2d, d2h, clocks

    # ---- resident run: `value`
    resident = SequenceBatch(model, datasets, True, 'robust', resident=True)
    ms, events, launches, _, _, clocks = timed(resident, K, Wm)
    frames = world * B * K
    value = frames / (ms * 1e-3)
    last_scores = resident.scores.cpu().numpy()
    net = voxel = None
    if rank == 0:
        padded = torch.empty((B, 5, resident.Hp, resident.Wp), dtype=torch.float32, device='cuda')
        _lib.check(_lib.load().evk_normalize_pad(_lib.ptr(resident.voxel), _lib.ptr(padded), B, 5, H, W, resident.Hp, resident.Wp, 1, _lib.stream_ptr()))
        net = network_roofline(model, padded, pk)
        del padded
    del resident
    torch.cuda.empty_cache()

    # ---- host-buffer run: `e2e`
    hosted = SequenceBatch(model, datasets, True, 'robust', resident=False)
    ms_e, events_e, _, h2d, d2h, _ = timed(hosted, K, Wm)
    e2e = {"value": frames / (ms_e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
           "events_per_s": events_e / (ms_e * 1e-3), "ms_per_step": ms_e / K,
           "api": "evreal_b200.pipeline.SequenceBatch(resident=False).step -> C ABI (evk_stage_windows_h2d, evk_voxelize_raw_batch, evk_normalize_pad, evk_model_forward, evk_crop, evk_percentile_normalize, evk_mse_ssim)"}
    del hosted
    torch.cuda.empty_cache()

    extras = {}
    cpu = None
    if rank == 0 and not args.fast:
        # ---- B = 1: one stream, the reference's own batch size
        single = SequenceBatch(make_model(), datasets[:1], True, 'robust', resident=True)
        ms1, ev1, _, _, _, _ = timed(single, max(K, 100), Wm, all_ranks=False)
        extras["single_stream"] = {"frames_per_s": max(K, 100) / (ms1 * 1e-3), "ms_per_frame": ms1 / max(K, 100), "batch_streams": 1,
                                   "what": "the same pipeline at B = 1 (latency form; the reference's own batch size)"}
        del single
        torch.cuda.empty_cache()
        extras["parity"] = parity_check(torch, make_model, streams[0][0], datasets[0])
        voxel = voxelizer_roofline(pk, torch, _lib)
        try:
            extras["gpu_reference"] = torch_gpu_reference(torch, B)
        except Exception as e:                     # a baseline leg must never take the bench line down
            extras["gpu_reference"] = {"error": repr(e)}
        torch.cuda.empty_cache()
        try:
            extras["lpips"] = lpips_throughput(torch, pk)
        except Exception as e:
            extras["lpips"] = {"error": repr(e)}
        torch.cuda.empty_cache()
    del model
    torch.cuda.empty_cache()
    if not args.fast:
        for tag, secs in (("cfg3", args.cfg3_seconds), ("cfg4", args.cfg4_seconds)):
            try:
                r = run_plugin_config(tag, torch, dist, rank, world, args, seconds=secs)
            except Exception as e:
                import traceback
                r = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
                if world > 1:
                    raise
            if rank == 0:
                extras[tag] = r
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = args.cpu_frames
            fps, evps, dt, tm = cpu_reference_frames_per_s(streams[0][0], 3, n_cpu, threads)
            cpu = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "events_per_s": evps,
                   "sample": "%d frames of stream 0 after 3 warm-up frames, batch 1 like the reference (%.1f s)" % (n_cpu, dt),
                   "stage_seconds": tm}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (split operands on the tensor cores: fp16 + 2 x fp8 products, or 3 bf16 products where a layer's shape rules that out; fp32 accumulate)", "data": "synthetic",
            "events_per_s": events / (ms * 1e-3),
            "config": {"workload": WORKLOAD, "batch_streams_per_gpu": B, "frames_per_step": B * world,
                       "l2": "inputs larger than L2: every step voxelizes a new window of %d resident streams (%.0f MB of raw events per GPU, each byte read once)" % (B, B * RATE * DURATION * 13 / 1e6),
                       "weights": "seeded random, shapes of pretrained/E2VID (10.7 M parameters)",
                       "overlap": "voxelizer of step i+1 and metrics of step i-1 on side streams; all streams join before the stop event"},
            "e2e": e2e, "gpu_launches": launches, "roofline": net, "voxelizer": voxel, "cpu_baseline": cpu, "clocks": clocks,
            "last_step_scores_mse_ssim": [[float(a), float(b)] for a, b in last_scores[:2]]}
    line.update(extras)
    if "parity" in extras:
        line["parity_ok"] = extras["parity"].get("parity_ok")
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=36, help="independent streams run in lock-step per GPU (36 measured best of 24..49: profiles/r01_batch_sweep_v7.txt)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=100, help="bounded CPU-baseline sample (frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--fast", action="store_true", help="headline value + e2e only (no cfg3 / cfg4 / baselines / sweeps)")
    ap.add_argument("--cfg3-seconds", type=float, default=20.0, help="length of each of the 16 HQF-shape sequences of cfg 3 (HQF's own sequences run 10-60 s)")
    ap.add_argument("--cfg4-seconds", type=float, default=4.0, help="length of each of the 8 MVSEC-shape sequences of cfg 4 (SURVEY 8d: 4 s)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: relaunch under torchrun
        port = 29500 + os.getpid() % 1000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + sys.argv
        return subprocess.call(cmd)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
