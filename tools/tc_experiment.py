"""Diagnostics for the tcgen05 convolution: per-layer CUDA-event times of E2VID at batch 8 (and, with
EVK_TC_TIMING=1, per-CTA clock64 phase breakdowns printed by the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evreal_b200 as evk
from evreal_b200 import synthetic
batch = int(os.environ.get('BATCH', '8'))
which = os.environ.get('MODEL', 'e2vid')
if which == 'firenet':
    model = evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(synthetic.firenet_state_dict(0)).to('cuda')
    x = torch.randn(batch, 5, 192, 240, device='cuda')
elif which == 'hyper':
    model = evk.E2VIDRecurrent(dict(synthetic.HYPER_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, dynamic_decoder=True)).to('cuda')
    x = torch.randn(batch, 5, 264, 352, device='cuda')
else:
    model = evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, norm_bn=True)).to('cuda')
    x = torch.randn(batch, 5, 184, 240, device='cuda')
agg = {}
frames = int(os.environ.get('FRAMES', '6'))
for f in range(frames):
    rows = model.profile_forward(x)
    if f < 2: continue
    for i, (d, ms, fl) in enumerate(rows):
        a = agg.setdefault(i, [d, 0.0, fl]); a[1] += ms / (frames - 2)
print('DBG', os.environ.get('EVK_TC_DBG'), 'total conv ms', sum(a[1] for a in agg.values() if a[0].startswith('conv')), 'all', sum(a[1] for a in agg.values()))
for i, a in sorted(agg.items()):
    print('  %-66s %.3f ms %6.1f TF/s' % (a[0], a[1], a[2] / a[1] / 1e9))
