"""Where the time of building a device program goes (model handle of the cfg-3 / cfg-4 shapes): the C-ABI calls timed from
Python.  python tools/build_phases.py"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from evreal_b200 import _lib, synthetic
import evreal_b200 as evk
torch.zeros(1).cuda()
lib = _lib.load()
cases = {'firenet B=2 180x240': (evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(synthetic.firenet_state_dict(0)), 2, 192, 240),
         'firenet B=16 180x240': (evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(synthetic.firenet_state_dict(0)), 16, 192, 240),
         'hyper B=8 260x346': (evk.E2VIDRecurrent(dict(synthetic.HYPER_KWARGS)).load_state_dict(synthetic.unet_state_dict(0, dynamic_decoder=True)), 8, 264, 352)}
for rep in range(2):
    for name, (model, B, H, W) in cases.items():
        dev = torch.device('cuda', 0)
        cfg = model._config(B, H, W)
        h = ctypes.c_void_p()
        t = [time.perf_counter()]
        _lib.check(lib.evk_model_create(ctypes.byref(cfg), ctypes.byref(h))); t.append(time.perf_counter())
        for n, ten in model._sd.items():
            shape = (ctypes.c_int64 * max(ten.dim(), 1))(*ten.shape)
            _lib.check(lib.evk_model_load_tensor(h, n.encode(), ctypes.c_void_p(ten.data_ptr()), shape, ten.dim()))
        t.append(time.perf_counter())
        _lib.check(lib.evk_model_finalize(h, _lib.stream_ptr(dev))); t.append(time.perf_counter())
        _lib.check(lib.evk_model_reset_states(h, _lib.stream_ptr(dev))); torch.cuda.synchronize(); t.append(time.perf_counter())
        x = torch.zeros(B, 5, H, W, device='cuda'); y = torch.empty(B, 1, H, W, device='cuda')
        for k in range(3):
            _lib.check(lib.evk_model_forward(h, _lib.ptr(x), _lib.ptr(y), _lib.stream_ptr(dev))); torch.cuda.synchronize(); t.append(time.perf_counter())
        lib.evk_model_destroy(h); torch.cuda.synchronize(); t.append(time.perf_counter())
        d = [1e3 * (b - a) for a, b in zip(t, t[1:])]
        print('%-22s create %.1f  load %.1f  finalize %.1f  reset %.1f  forward#1 %.1f  #2 %.1f  #3 %.1f  destroy %.1f ms' % ((name,) + tuple(d)), flush=True)
