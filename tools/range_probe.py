"""Activation-range probe of the mixed operand decomposition (fp16 saturates at 65504): E2VID topology without BatchNorm, seeded
weights, inputs scaled up -- error against the CPU oracle with mixed operands and with bf16x3 (EVK_MIXED=0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from evreal_b200 import E2VIDRecurrent, synthetic
from oracle import networks as on
sd = synthetic.unet_state_dict(3, norm_bn=False)
w = {k[len('unetrecurrent.'):]: v for k, v in sd.items()}
kw = dict(synthetic.E2VID_KWARGS); kw['norm'] = None
g = torch.Generator().manual_seed(1)
for scale in (1.0, 30.0, 1000.0, 30000.0):
    xs = [torch.randn(1, 5, 48, 64, generator=g) * scale for _ in range(3)]
    o = on.UNetRecurrentOracle(w, 3, 2, final_sigmoid=True)
    res = {}
    for mode in ('1', '0'):
        os.environ['EVK_MIXED'] = mode
        m = E2VIDRecurrent(kw).load_state_dict(sd).to('cuda'); m.reset_states()
        res[mode] = [m(x.cuda())['image'].cpu().numpy() for x in xs]
    errs = {}
    o.reset_states()
    refs = [o(x).numpy() for x in xs]
    for mode in res:
        errs[mode] = max(float(np.max(np.abs(a - r)) / max(np.max(np.abs(r)), 1e-6)) for a, r in zip(res[mode], refs))
    # largest activation entering a mixed layer ~ first encoder output
    print('input scale %g: mixed %.2e  bf16x3 %.2e' % (scale, errs['1'], errs['0']), flush=True)
