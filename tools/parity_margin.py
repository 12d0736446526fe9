"""Parity margin of the two operand decompositions on the shipped checkpoints: max|frame - golden| / max|golden| per checkpoint
(frames of the REAL reference classes, tests/golden/full_checkpoints.npz / spade.npz / etnet.npz) with mixed operands
(fp16 + 2 x fp8, the default) and with bf16x3 (EVK_MIXED=0).  The bar is 1e-4.  Prints one JSON line per checkpoint."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import GOLDEN, gen_events, golden
from evreal_b200 import evaluate as ev
from evreal_b200.util import CropParameters, normalize_pad
from oracle import event_voxel as ov


def run(name, path, H, W, norm_ev, ref):
    model = ev.get_model_from_checkpoint_path(name, path)
    crop = CropParameters(W, H, model.num_encoders)
    model.reset_states()
    errs = []
    for f in range(ref.shape[0]):
        e = gen_events(40 + f, 15000 + 7000 * f, H, W)
        v = ov.events_to_voxel_oracle(*[torch.from_numpy(a) for a in e], 5, (H, W))
        x = normalize_pad(v[None].cuda(), crop.height_crop_size, crop.width_crop_size, bool(norm_ev))
        got = crop.crop(model(x)['image'])[0, 0].cpu().numpy()
        errs.append(float(np.max(np.abs(got - ref[f])) / np.max(np.abs(ref[f]))))
    desc = model.op_descriptions()
    return max(errs), sum('f16+2xf8' in d for d in desc), sum('tcgen05' in d for d in desc)


g = golden('full_checkpoints')
cases = [(n, g[n + '.frames'], [int(v) for v in g[n + '.meta']]) for n in ('E2VID', 'E2VID+', 'HyperE2VID', 'SSL-E2VID', 'FireNet', 'FireNet+')]
for extra, key in (('SPADE-E2VID', 'spade'), ('ET-Net', 'etnet')):
    cases.append((extra, golden(key)['ckpt.frames'], [180, 240, 0, 3]))
for name, ref, meta in cases:
    path = os.path.join(GOLDEN, '_ckpt', name + '.pth')
    if not os.path.exists(path):
        continue
    row = {'checkpoint': name, 'frames': int(ref.shape[0]), 'size': meta[:2]}
    for mode, env in (('mixed', '1'), ('bf16x3', '0')):
        os.environ['EVK_MIXED'] = env
        err, n_mixed, n_tc = run(name, path, meta[0], meta[1], meta[2], ref)
        row[mode] = {'max_rel_err': err, 'mixed_layers': n_mixed, 'tensor_core_layers': n_tc}
    print(json.dumps(row), flush=True)
