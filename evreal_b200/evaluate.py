"""Host-side mirror of EVREAL's ``eval.py`` driving the CUDA hot path.

Keeps the reference's plugin surface -- ``config/method/*.json``,
``config/eval/*.json``, ``config/dataset/*.json`` (read relative to a config
root), the checkpoint -> model factory (eval.py:124-158), the per-sequence loop
(eval.py:189-246) and the count-weighted dataset means (eval.py:249-276,
:367-368) -- and adds what the reference lacks: sequences sharded over ranks
(one process per GPU) with a single all-reduce of the ``[sum(score*n), sum(n)]``
vectors at the end.

Everything per frame stays on the GPU: raw events are resident in HBM, the
voxel grid, normalisation+padding, network, crop, percentile normalisation and
MSE/SSIM are kernels of libevreal_b200.so; scores come back to the host once per
sequence.
"""
import glob
import math
import os
import traceback
from collections import OrderedDict

import torch

from . import model as model_arch
from . import parse_config
from .dataset import MemMapDataset
from .eval_metrics import EvalMetricsTracker
from . import eval_utils
from .eval_utils import post_process_normalization, torch2cv2
from .util import CropParameters, read_json, normalize_pad


# ------------------------------------------------------------------ configs
def get_eval_configs(eval_config_names, config_root="config"):
    out = []
    for name in eval_config_names:
        cfg = read_json(os.path.join(config_root, "eval", name + ".json"))
        cfg['name'] = name
        out.append(cfg)
    return out


def get_dataset_configs(dataset_names, config_root="config"):
    out = []
    for name in dataset_names:
        cfg = read_json(os.path.join(config_root, "dataset", name + ".json"))
        cfg['name'] = name
        out.append(cfg)
    return out


def get_method_config(method_name, config_root="config"):
    return read_json(os.path.join(config_root, "method", method_name + ".json"))


def get_sequences(dataset_config, dataset_kwargs):
    """eval.py:38-79, with sequence names sorted (the reference iterates an unsorted glob)."""
    dataset_root = dataset_config['root_path']
    dataset_kwargs = dict(dataset_kwargs)
    dataset_kwargs.update(dataset_config.get('dataset_kwargs', {}))
    if dataset_config.get('get_all_sequences', False):
        pattern = os.path.join(dataset_root, '*', '*') if dataset_config.get('has_subfolders', False) \
            else os.path.join(dataset_root, '*')
        sequences_config = OrderedDict()
        for path in sorted(glob.glob(pattern)):
            name = os.path.basename(path)
            if dataset_config.get('has_subfolders', False):
                name = os.path.basename(os.path.dirname(path)) + "_" + name
            sequences_config[name] = {'sequence_path': path}
    else:
        sequences_config = dataset_config.get('sequences', {})
    sequences = []
    for name, seq in sequences_config.items():
        seq = dict(seq)
        seq['name'] = name
        seq['sequence_path'] = seq.get('sequence_path', os.path.join(dataset_root, name))
        seq['dataset_kwargs'] = dataset_kwargs
        sequences.append(seq)
    return sequences


def open_sequence(sequence, device=None):
    """Instantiate the dataset of one sequence and fill in the default evaluation window (eval.py:71-77)."""
    ds = MemMapDataset(sequence['sequence_path'], device=device, **sequence['dataset_kwargs'])
    min_t, max_t = ds.get_min_max_t()
    sequence.setdefault('start_time_s', min_t)
    sequence.setdefault('end_time_s', max_t)
    sequence['dataset'] = ds
    return ds


# ------------------------------------------------------------------ model factory
def build_model(model_name, checkpoint):
    """Checkpoint dialects by method name (eval.py:124-158) -> (model, state_dict)."""
    if model_name == "SPADE-E2VID":                       # eval.py:130-133: bare state_dict, num_encoders set on the instance
        model = model_arch.SpadeE2vid()
        model.num_encoders = 3
        return model, checkpoint
    if model_name == "SSL-E2VID":
        unet_kwargs = {"base_num_channels": 32, "kernel_size": 5, "num_bins": 5, "num_encoders": 3,
                       "recurrent_block_type": "convlstm", "num_residual_blocks": 2, "skip_type": "sum", "norm": None,
                       "use_upsample_conv": True}
        return model_arch.E2VIDRecurrent(unet_kwargs), checkpoint
    if model_name == "E2VID":
        unet_kwargs = dict(checkpoint['model'])
        unet_kwargs['final_activation'] = 'sigmoid'
        model = model_arch.E2VIDRecurrent(unet_kwargs)
    elif model_name == "FireNet":
        unet_kwargs = dict(checkpoint['config']['model'])
        unet_kwargs['final_activation'] = ''
        model = model_arch.FireNet_legacy(unet_kwargs)
    else:
        model = checkpoint['config'].init_obj('arch', model_arch)
        if model_name == "ET-Net":                        # eval.py:149-152
            model.num_encoders = 3
        elif model_name == "FireNet+":
            model.num_encoders = 0
    return model, checkpoint['state_dict']


def get_model_from_checkpoint_path(model_name, checkpoint_path, device=None):
    parse_config.install()
    checkpoint = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
    model, state_dict = build_model(model_name, checkpoint)
    model.load_state_dict(state_dict)
    model.to(device if device is not None else torch.device('cuda', torch.cuda.current_device()))
    model.eval()
    return model


# ------------------------------------------------------------------ per-sequence loop
def eval_method_on_sequence(dataset_name, eval_config, method_name, model, method_config, sequence, metrics,
                            output_root="outputs", write_files=True, collect_images=None):
    """eval.py:189-246 on the GPU.  Returns (num_evaluated, mean_scores, num_frames_reconstructed, num_events)."""
    dataset = sequence.get('dataset') or open_sequence(sequence)
    has_reference_frames = dataset.has_images
    output_dir = os.path.join(output_root, eval_config['name'], dataset_name, sequence['name'], method_name)
    save_images = eval_config.get('save_images', True) and write_files
    tracker = EvalMetricsTracker(save_images=save_images,
                                 save_processed_images=save_images and eval_config['histeq'] != 'none',
                                 output_dir=output_dir, hist_eq=eval_config['histeq'],
                                 quan_eval_metric_names=metrics,
                                 quan_eval_start_time=sequence['start_time_s'],
                                 quan_eval_end_time=sequence['end_time_s'],
                                 quan_eval_ts_tol_ms=eval_config['ts_tol_ms'],
                                 has_reference_frames=has_reference_frames,
                                 color=eval_config.get('color', False), defer=True, write_files=write_files)
    color = eval_config.get('color', False)
    height, width = int(dataset.sensor_resolution[0]), int(dataset.sensor_resolution[1])
    cropper = CropParameters(width, height, model.num_encoders)
    model.reset_states()
    eval_infer_all = eval_config.get('eval_infer_all', False)
    post_process_norm = method_config.get('post_process_norm', "none")
    event_tensor_normalization = method_config.get('event_tensor_normalization', False)
    idx = 0
    frames = 0
    events = 0
    for idx in range(len(dataset)):
        item = dataset[idx]
        ref_frame = item['frame'] if has_reference_frames else None
        ref_frame_ts = item['frame_timestamp'].item() if has_reference_frames else None
        pred_frame_ts = item['voxel_timestamp'].item()
        # Only start reconstruction when close to eval start (10 seconds)   eval.py:210-216
        if pred_frame_ts < sequence['start_time_s'] - 10 and not eval_infer_all:
            continue
        if pred_frame_ts > sequence['end_time_s'] and not eval_infer_all:
            idx -= 1
            break
        if item['event_count'] <= 1 or item['dt'].item() == 0:
            event_rate = 0
        else:
            event_rate = item['event_count'] / item['dt'].item()
        voxel = item['events'].unsqueeze(0)
        if color:
            # ColorNet pads its Bayer sites itself and returns the merged [3, H, W] frame (eval.py:225-232)
            if event_tensor_normalization:
                voxel = normalize_pad(voxel, height, width, True)
            image = post_process_normalization(torch2cv2(model(voxel)['image']), post_process_norm)
            if collect_images is not None:
                collect_images.append(image.copy())
            tracker.update(idx, image, torch2cv2(ref_frame) if has_reference_frames else None, pred_frame_ts, ref_frame_ts)
            tracker.save_custom_metric(idx, "event_rate", event_rate)
            frames += 1
            events += item['event_count']
            continue
        # normalize_event_tensor + cropper.pad in one kernel (eval.py:222-226)
        voxel = normalize_pad(voxel, cropper.height_crop_size, cropper.width_crop_size, event_tensor_normalization)
        output = model(voxel)
        image = cropper.crop(output['image'])
        image = post_process_normalization(image, post_process_norm)
        image2d = image[0, 0]
        if collect_images is not None:
            collect_images.append(image2d.clone())
        tracker.update(idx, image2d, ref_frame[0] if has_reference_frames else None, pred_frame_ts, ref_frame_ts)
        tracker.save_custom_metric(idx, "event_rate", event_rate)
        frames += 1
        events += item['event_count']
    tracker.finalize(idx)
    dataset.check_bounds()
    return tracker.get_num_quan_evaluations(), tracker.get_mean_scores(), frames, events


LOCKSTEP_METRICS = ('mse', 'ssim', 'lpips', 'lpips-vgg')


def lockstep_supported(eval_config, metrics, datasets):
    """The lock-step form covers what SequenceBatch computes on the device: reference frames present, one sensor resolution
    and bin count, MSE / SSIM / LPIPS, no histogram equalisation, no files, no colour.  All three windowing modes qualify
    (frame pairing and score gates come from MemMapDataset.item_meta)."""
    if eval_config.get('color', False) or not set(metrics) <= set(LOCKSTEP_METRICS) or not datasets:
        return False
    if eval_config.get('histeq', 'none') != 'none':
        return False
    res0 = tuple(int(v) for v in datasets[0].sensor_resolution[:2])
    return all(ds.has_images and ds.num_bins == datasets[0].num_bins and
               tuple(int(v) for v in ds.sensor_resolution[:2]) == res0 for ds in datasets)


def lockstep_item_range(ts, start_time_s, end_time_s, infer_all=False, frame_ts=None, ts_tol_ms=float('inf')):
    """(first item, number of items, score gate per item) of one sequence from its item (voxel) timestamps: reconstruction
    starts at the first item within 10 s of start_time_s and stops before the first item after end_time_s (eval.py:210-216);
    an item's scores count when start_time_s <= t <= end_time_s and its reference frame is within ts_tol_ms of it
    (utils/eval_metrics.py:256-262; with 'between_frames' windows the two timestamps coincide)."""
    first, last = 0, len(ts) - 1
    if not infer_all:
        first = next((i for i, t in enumerate(ts) if not t < start_time_s - 10), len(ts))
        last = next((i - 1 for i, t in enumerate(ts) if i >= first and t > end_time_s), len(ts) - 1)
    count = max(last - first + 1, 0)
    gates = []
    for i in range(first, first + count):
        ok = start_time_s <= ts[i] <= end_time_s
        if ok and frame_ts is not None:
            ok = abs(frame_ts[i] - ts[i]) * 1000 <= ts_tol_ms
        gates.append(ok)
    return first, count, gates


def eval_method_on_sequences_lockstep(eval_config, model, method_config, sequences, metrics, lpips_weights=None):
    """eval_method_on_sequence for several sequences at once: frame k of every sequence in ONE batched voxelizer / network /
    metric launch (pipeline.SequenceBatch).  Per-sequence semantics are the reference's (eval.py:203-246): reconstruction
    starts at the first item within 10 s of start_time_s and stops after end_time_s, scores count inside
    [start_time_s, end_time_s] and within ts_tol_ms of the reference frame, the mean is sum(scores) / n over finite scores.
    Returns [(num_evaluated, mean_scores)]."""
    from .pipeline import SequenceBatch
    datasets = [s.get('dataset') or open_sequence(s) for s in sequences]
    infer_all = eval_config.get('eval_infer_all', False)
    offsets, counts, gates = [], [], []
    for seq, ds in zip(sequences, datasets):
        meta = [ds.item_meta(i) for i in range(len(ds))]
        first, count, gate = lockstep_item_range([m[3] for m in meta], seq['start_time_s'], seq['end_time_s'], infer_all,
                                                 [m[4] for m in meta], eval_config.get('ts_tol_ms', float('inf')))
        offsets.append(first)
        counts.append(count)
        gates.append(gate)
    steps = max(counts) if counts else 0
    lp_name = next((m for m in metrics if m in ('lpips', 'lpips-vgg')), None)
    lpips = None
    if lp_name is not None:
        from .lpips import LpipsMetric
        lpips = (lp_name, lpips_weights if lpips_weights is not None else LpipsMetric(lp_name)._weights())
    sc = None
    if steps > 0:
        batch = SequenceBatch(model, datasets, method_config.get('event_tensor_normalization', False),
                              method_config.get('post_process_norm', 'none'), resident=True, offsets=offsets, counts=counts,
                              lpips=lpips, log_scores=True)
        batch.reset()
        for k in range(steps):
            _, _, n_ev = batch.step(k, sync=False)
            last_timings['events'] = last_timings.get('events', 0) + n_ev
        last_timings['frames'] = last_timings.get('frames', 0) + sum(counts)
        batch.finish()
        batch.check_bounds()
        sc = batch.scores_log.cpu().numpy()
    col = {'mse': 0, 'ssim': 1, 'lpips': 2, 'lpips-vgg': 2}
    out = []
    for b in range(len(datasets)):
        n_eval = sum(gates[b])
        means = {}
        for name in metrics:
            vals = [float(sc[k, b, col[name]]) for k in range(counts[b]) if gates[b][k]]
            vals = [v for v in vals if math.isfinite(v)]
            means[name] = sum(vals) / len(vals) if vals else -1
        out.append((n_eval, means))
    return out


class MetricTracker:
    """Count-weighted running means per metric (eval.py:249-276)."""

    def __init__(self):
        self.data_dict = {}

    def init_key(self, key):
        self.data_dict[key] = {'total': 0.0, 'count': 0, 'average': 0.0}

    def update(self, key, value, count=1):
        if count == 0:
            return
        if key not in self.data_dict:
            self.init_key(key)
        d = self.data_dict[key]
        d['total'] += value * count
        d['count'] += count
        d['average'] = d['total'] / d['count']

    def get_average(self, key):
        if key not in self.data_dict:
            self.init_key(key)
        return self.data_dict[key]['average']

    def get_count(self, key):
        if key not in self.data_dict:
            self.init_key(key)
        return self.data_dict[key]['count']


# ------------------------------------------------------------------ sharding over ranks
def shard_sequences(sequences, weights, world_size):
    """Longest-processing-time assignment of sequences to ranks.  Deterministic: ties broken by name.
    Returns a list (per rank) of indices into ``sequences``."""
    order = sorted(range(len(sequences)), key=lambda i: (-weights[i], sequences[i]['name']))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += weights[i]
    return [sorted(s) for s in shards]


def reduce_metric_sums(local, metric_names, group=None):
    """One all-reduce(SUM) of [sum(score*n), sum(n)] per metric -> MetricTracker semantics over all ranks."""
    import torch.distributed as dist
    vec = torch.zeros(2 * len(metric_names), dtype=torch.float64)
    for j, name in enumerate(metric_names):
        if name in local.data_dict:
            vec[2 * j] = local.data_dict[name]['total']
            vec[2 * j + 1] = local.data_dict[name]['count']
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == 'nccl':
            vec = vec.cuda()
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        vec = vec.cpu()
    out = MetricTracker()
    for j, name in enumerate(metric_names):
        total, count = float(vec[2 * j]), int(round(float(vec[2 * j + 1])))
        if count > 0:
            out.data_dict[name] = {'total': total, 'count': count, 'average': total / count}
    return out


# wall-clock seconds of the last evaluate() call on this rank, by phase (model construction, sequence open + upload,
# per-frame loop, all-reduce), frames reconstructed / events voxelized, failures caught; diagnostic only
last_timings = {}


def _log_exception(what, e):
    print(what)
    print(e)
    print(traceback.format_exc())


def evaluate(method_names, eval_config_names=None, dataset_names=None, metrics=None, config_root="config",
             output_root="outputs", write_files=True, rank=0, world_size=1, lockstep=0, lpips_weights=None,
             async_writer=True):
    """eval.py:413-444.  Returns {eval_config: {method: {dataset: MetricTracker}}} (identical on every rank).
    ``lockstep`` = B > 1: this rank's sequences run B at a time in lock-step (eval_method_on_sequences_lockstep) where that
    form applies (lockstep_supported, write_files False); everything else runs one sequence at a time like the reference.

    Failures are caught where the reference catches them (eval.py:344-352 per method, :357-375 per dataset): a method whose
    model cannot be built is reported and skipped, a dataset that raises keeps the sequences it finished, and the run goes
    on.  With several ranks every rank still takes part in every per-dataset all-reduce (a failed rank contributes what it
    has, possibly nothing), so a failure on one rank never leaves the others waiting in a collective.
    ``async_writer``: per-frame text / PNG output goes through one writer thread (eval_utils.AsyncWriter), flushed per
    sequence -- same files, off the critical path."""
    if eval_config_names is None:
        eval_config_names = ['std']
    if metrics is None:
        metrics = ['mse', 'ssim']
    import time
    results = OrderedDict()
    tm = {'model_s': 0.0, 'open_s': 0.0, 'loop_s': 0.0, 'reduce_s': 0.0, 'failures': 0, 'frames': 0, 'events': 0}
    last_timings.clear()
    last_timings.update(tm)
    writer = eval_utils.AsyncWriter() if (write_files and async_writer) else None
    if lpips_weights is not None:
        from . import lpips as lpips_mod
        lpips_mod.set_default_weights(lpips_weights)
    old_writer = eval_utils.set_writer(writer)
    try:
        for eval_config in get_eval_configs(eval_config_names, config_root):
            per_method = OrderedDict()
            dataset_configs = get_dataset_configs(dataset_names, config_root)
            for method_name in method_names:
                model = method_config = None
                try:
                    method_config = get_method_config(method_name, config_root)
                    t0 = time.perf_counter()
                    model = get_model_from_checkpoint_path(method_config['model_name'], method_config['model_path'])
                    if eval_config.get('color', False):
                        model = model_arch.ColorNet(model)
                    last_timings['model_s'] += time.perf_counter() - t0
                except Exception as e:
                    last_timings['failures'] += 1
                    _log_exception(f"Exception while getting method {method_name}", e)
                    model = None
                per_dataset = OrderedDict()
                for dataset_config in dataset_configs:
                    local = MetricTracker()
                    if model is not None:
                        try:
                            _eval_dataset(eval_config, method_name, model, method_config, dataset_config, metrics, output_root,
                                          write_files, rank, world_size, lockstep, lpips_weights, local, writer)
                        except Exception as e:
                            last_timings['failures'] += 1
                            _log_exception(f"Exception while evaluating method {method_name} on {dataset_config['name']} dataset:", e)
                    # every rank, always: the collective count stays matched whatever happened above
                    t0 = time.perf_counter()
                    per_dataset[dataset_config['name']] = reduce_metric_sums(local, metrics)
                    last_timings['reduce_s'] += time.perf_counter() - t0
                per_method[method_name] = per_dataset
            results[eval_config['name']] = per_method
    finally:
        eval_utils.set_writer(old_writer)
        if writer is not None:
            try:
                writer.flush()
            finally:
                writer.close()
    return results


def _eval_dataset(eval_config, method_name, model, method_config, dataset_config, metrics, output_root, write_files, rank,
                  world_size, lockstep, lpips_weights, local, writer):
    """This rank's sequences of one dataset (the body of eval.py:357-368); updates ``local`` sequence by sequence, so what
    was finished before an exception is kept (eval.py:373-375)."""
    import time
    sequences = get_sequences(dataset_config, eval_config.get('dataset_kwargs', {}))
    weights = [os.path.getsize(os.path.join(s['sequence_path'], 'events_ts.npy')) for s in sequences]
    mine = shard_sequences(sequences, weights, world_size)[rank]
    if lockstep > 1 and not write_files and mine:
        t0 = time.perf_counter()
        for i in mine:
            open_sequence(sequences[i])
        last_timings['open_s'] += time.perf_counter() - t0
        if lockstep_supported(eval_config, metrics, [sequences[i]['dataset'] for i in mine]):
            t0 = time.perf_counter()
            for c0 in range(0, len(mine), lockstep):
                chunk = [sequences[i] for i in mine[c0:c0 + lockstep]]
                for n_eval, mean_scores in eval_method_on_sequences_lockstep(eval_config, model, method_config, chunk, metrics,
                                                                             lpips_weights):
                    for metric_name, score in mean_scores.items():
                        local.update(metric_name, score, n_eval)
            last_timings['loop_s'] += time.perf_counter() - t0
            for i in mine:
                sequences[i].pop('dataset', None)
            mine = []
    for i in mine:
        seq = sequences[i]
        t0 = time.perf_counter()
        open_sequence(seq)
        t1 = time.perf_counter()
        try:
            n_eval, mean_scores, n_frames, n_events = eval_method_on_sequence(
                dataset_config['name'], eval_config, method_name, model, method_config, seq, metrics,
                output_root, write_files)
            last_timings['frames'] += n_frames
            last_timings['events'] += n_events
        finally:
            if writer is not None:
                writer.flush()                # the sequence's files are complete before the next one starts
        last_timings['open_s'] += t1 - t0
        last_timings['loop_s'] += time.perf_counter() - t1
        for metric_name, score in mean_scores.items():
            local.update(metric_name, score, n_eval)
        seq.pop('dataset', None)


def main(argv=None):
    """The reference's command line (eval.py:447-455: -c / -m / -d / -qm) over evaluate().  Under torchrun (RANK /
    WORLD_SIZE / LOCAL_RANK in the environment) the process joins an NCCL group and takes its shard of the sequences."""
    import argparse
    ap = argparse.ArgumentParser(description='event2im evaluation on the B200 hot path (flags of EVREAL eval.py)')
    ap.add_argument('-c', '--config', nargs='+', type=str, help='evaluation configs')
    ap.add_argument('-m', '--method', nargs='+', type=str, required=True, help='methods')
    ap.add_argument('-d', '--dataset', nargs='+', type=str, required=True, help='datasets')
    ap.add_argument('-qm', '--metrics', nargs='+', type=str, help='quantitative evaluation metrics')
    ap.add_argument('--config-root', default='config')
    ap.add_argument('--output-root', default='outputs')
    ap.add_argument('--lockstep', type=int, default=0, help='sequences run in lock-step per GPU (0: one at a time, like the reference)')
    ap.add_argument('--lpips-weights', default=None, help='torch file with the LPIPS tensors (pyiqa downloads them; none ship here)')
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
    weights = torch.load(args.lpips_weights, map_location='cpu') if args.lpips_weights else None
    results = evaluate(args.method, args.config, args.dataset, args.metrics, config_root=args.config_root,
                       output_root=args.output_root, write_files=args.lockstep <= 1, rank=rank, world_size=world,
                       lockstep=args.lockstep, lpips_weights=weights)
    if rank == 0:
        for cfg_name, per_method in results.items():
            for method, per_dataset in per_method.items():
                for dataset, tracker in per_dataset.items():
                    scores = {k: v['average'] for k, v in tracker.data_dict.items()}
                    print(f"{cfg_name} / {method} / {dataset}: {scores}")
    if world > 1:
        torch.distributed.destroy_process_group()
    return results


if __name__ == '__main__':
    main()
