import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
import evreal_b200 as evk
from evreal_b200 import synthetic
torch.zeros(1).cuda()
for name in ('hyper', 'e2vid'):
    for mx in ('1', '0'):
        os.environ['EVK_MIXED'] = mx
        if name == 'hyper':
            m = evk.E2VIDRecurrent(dict(synthetic.HYPER_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, dynamic_decoder=True)).to('cuda'); x = torch.randn(8, 5, 264, 352, device='cuda')
        else:
            m = evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, norm_bn=True)).to('cuda'); x = torch.randn(8, 5, 184, 240, device='cuda')
        torch.cuda.synchronize(); t0 = time.time(); m(x); torch.cuda.synchronize(); t1 = time.time(); m(x); m(x); torch.cuda.synchronize(); t2 = time.time()
        print(name, 'mixed' if mx == '1' else 'bf16x3', 'first forward (build) %.3f s, next two %.4f s' % (t1 - t0, t2 - t1), flush=True)
