"""Where a lock-step SequenceBatch step spends its time: graph forward alone vs every pre/post launch (CUDA events, eager)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evreal_b200 as evk
from evreal_b200 import synthetic, _lib
from evreal_b200.dataset import MemMapDataset
from evreal_b200.pipeline import SequenceBatch

B = int(os.environ.get('BATCH', '24'))
H, W, rate, fps = 180, 240, 1e6, 24.0
steps = 30
dss = [MemMapDataset(synthetic.make_stream(H, W, rate, (steps + 12) / fps, fps, seed=b), num_bins=5,
                     voxel_method={'method': 'between_frames'}, resident=False) for b in range(B)]
model = evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, norm_bn=True)).to('cuda')
batch = SequenceBatch(model, dss, True, 'robust', resident=True)
batch.reset()
for i in range(1, 6):
    batch.step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(6, 6 + steps):
    batch.step(i)
e1.record(); torch.cuda.synchronize()
print('step ms', e0.elapsed_time(e1) / steps)
e0.record()
for i in range(steps):
    model.forward_into(batch.padded, batch.recon_p)
e1.record(); torch.cuda.synchronize()
print('graph forward ms', e0.elapsed_time(e1) / steps)
# monkeypatched per-call timing of the C ABI entry points used by step()
lib = batch.lib
names = ['evk_voxelize_raw_batch', 'evk_u8_to_f32_batch', 'evk_normalize_pad', 'evk_crop', 'evk_percentile_normalize', 'evk_mse_ssim']
acc = {n: 0.0 for n in names}
class Wrap:
    def __init__(self, lib): self._lib = lib
    def __getattr__(self, n):
        f = getattr(self._lib, n)
        if n not in acc: return f
        def g(*a):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); r = f(*a); e.record(); e.synchronize(); acc[n] += s.elapsed_time(e); return r
        return g
batch.lib = Wrap(lib)
for i in range(6, 6 + steps):
    batch.step(i)
torch.cuda.synchronize()
print(json.dumps({n: round(v / steps * 1000, 1) for n, v in acc.items()}), 'us per step; sum', round(sum(acc.values()) / steps * 1000, 1))
