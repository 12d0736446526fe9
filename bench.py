#!/usr/bin/env python
"""Benchmark of the B200-native EVREAL hot path (BASELINE.json: frames/s reconstructed + events/s voxelized).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (N=1 and per rank for N>1, weak scaling): BASELINE.json configs[1] -- E2VID (BN, sigmoid, base 32; seeded
random weights of the shipped checkpoint's shapes) on synthetic ECD-shape streams: 240x180, 1 Mev/s, 24 Hz frames,
5 bins, 'between_frames' windows (~41.7k events), normalize_event_tensor on, pad to 184x240, 'robust' percentile
post-normalisation, clip, MSE + SSIM per frame.  One STEP = frame i of B independent streams run in lock-step
(SequenceBatch): one batched voxelizer launch + one batched network forward + one batched metric launch.  `value` = frames/s
summed over all ranks with the raw event arrays resident in HBM; `e2e` = the same loop with the event arrays and
reference frames in pinned HOST memory, every window copied host->device and scores + reconstructed frames copied
device->host inside the timed region (copy streams: the windows of step i+1 are staged while step i computes).  Sequences are independent: ranks share nothing and there is no data-path
collective (the only collective of the product, one all-reduce of metric sums, runs once after the timed region).

--impl reference: the reference's torch-CPU path for the same config (oracle/eval_loop.py: the same ATen/oneDNN
operators eval.py executes per frame, all host threads), one stream, one frame per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
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
    """nvidia-smi clocks / throttle reasons sampled every 50 ms DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        import torch
        self.file = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(device_index).uuid)
            sel = ["-i", uuid]
        except Exception:
            sel = ["-i", str(device_index)]
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"] + sel, stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        rows = [r.strip().split(', ') for r in open(self.file.name).read().splitlines() if r.strip()]
        os.unlink(self.file.name)
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for n, v in zip(names, r[4:8]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm))
        return out


def make_streams(n_streams, seed0, duration):
    from evreal_b200 import synthetic
    from evreal_b200.dataset import MemMapDataset
    from concurrent.futures import ThreadPoolExecutor
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workers = max(1, min(n_streams, (os.cpu_count() or 1) // max(world, 1)))     # numpy's generators and sort release the GIL
    with ThreadPoolExecutor(max_workers=workers) as pool:
        arrs = list(pool.map(lambda b: synthetic.make_stream(H, W, RATE, duration, FPS, seed=seed0 + b), range(n_streams)))
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
    """cfg 5 point: 640x480, 5 bins, 4M events/window (100 Mev/s at 25 windows/s), f32 SoA resident in HBM.
    Algorithmic bytes = 16*N + 4*bins*H*W.  Four event sets (256 MB > 126 MB L2) are rotated so no
    iteration re-reads L2-resident inputs."""
    import numpy as np
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

    def measure(n_ev, iters):
        def run(i):
            x, y, t, p = evs[i % sets]
            _lib.check(lib.evk_voxelize(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), _lib.ptr(p), n_ev, bins, Hv, Wv, _lib.ptr(grid), None, st))
        for i in range(4):
            run(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        bytes_alg = 16.0 * n_ev + 4.0 * bins * Hv * Wv
        return ms, bytes_alg, bytes_alg / (ms * 1e-3) / 1e9

    # cfg 5 sweep: 1..100 Mev/s at 25 windows/s -> 40k..4M events per window (a prefix of the sorted event sets)
    sweep = []
    for rate, n_ev in ((1, 40_000), (2, 80_000), (5, 200_000), (10, 400_000), (20, 800_000), (50, 2_000_000), (100, 4_000_000)):
        ms_i, _, gbs_i = measure(n_ev, 20)
        sweep.append({"Mev_per_s_stream": rate, "events_per_window": n_ev, "us_per_window": ms_i * 1e3,
                      "events_per_s": n_ev / (ms_i * 1e-3), "GB_per_s": gbs_i, "frac": gbs_i / pk["hbm_gbs"]})
    ms, bytes_alg, gbs = measure(n, 20)
    return {"workload": "voxelizer only, 640x480, 5 bins, 4M events/window (cfg 5 top point), f32 SoA in HBM, 4 rotating event sets (256 MB > L2)",
            "events_per_s": n / (ms * 1e-3), "ms_per_window": ms,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "traffic": None, "algorithmic_bytes_per_launch": bytes_alg, "peak_source": pk["source"],
                         "note": "one RED.ADD.V4.F32 per event into an L2-resident interleaved grid: bound by the L2 reduction request rate (~83/clk), not by HBM"},
            "sweep_cfg5": sweep}


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
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("conv_dram_bytes_per_forward")
        except Exception:
            traffic = None
    return {"bound": "tensor", "achieved": ach, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
            "frac": ach / pk["tflops_sustained"], "traffic": traffic,
            "kernel": "implicit-GEMM convolution family (all conv launches of one forward)",
            "launches_per_forward": len(conv), "algorithmic_flops_per_forward": conv_fl,
            "share_of_forward_time": conv_ms / total_ms if total_ms > 0 else None,
            "forward_ms_eager": total_ms, "peak_source": pk["source"] + ", sustained bf16", "layers": layers}


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
    model = evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(e2vid_weights()).to(torch.device('cuda', local))
    method = {'event_tensor_normalization': True, 'post_process_norm': 'robust'}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch, k_steps, warm):
        batch.reset()
        idx = 1                                   # item 0 of 'between_frames' is always the empty window
        for _ in range(warm):
            nxt = idx % (n_items - 1) + 1
            batch.step(idx, nxt)
            idx = nxt
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        launches = events = h2d = d2h = 0
        e0.record()
        for _ in range(k_steps):
            nxt = idx % (n_items - 1) + 1
            _, _, n_ev = batch.step(idx, nxt)
            launches += batch.launches
            events += n_ev
            h2d += batch.h2d_bytes
            d2h += batch.d2h_bytes
            idx = nxt
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        tot = torch.tensor([float(events), float(launches)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
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
        net = network_roofline(model, resident.padded, pk)
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

    cpu = None
    if rank == 0:
        voxel = voxelizer_roofline(pk, torch, _lib)
        if world == 1 and not args.no_cpu_baseline:
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
            "dtype": "f32 (bf16x3 split on tensor cores where enabled, fp32 accumulate)", "data": "synthetic",
            "events_per_s": events / (ms * 1e-3),
            "config": {"workload": WORKLOAD, "batch_streams_per_gpu": B, "frames_per_step": B * world,
                       "l2": "inputs larger than L2: every step voxelizes a new window of %d resident streams (%.0f MB of raw events per GPU, each byte read once)" % (B, B * RATE * DURATION * 13 / 1e6),
                       "weights": "seeded random, shapes of pretrained/E2VID (10.7 M parameters)"},
            "e2e": e2e, "gpu_launches": launches, "roofline": net, "voxelizer": voxel, "cpu_baseline": cpu, "clocks": clocks,
            "last_step_scores_mse_ssim": [[float(a), float(b)] for a, b in last_scores[:2]]}
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
