"""Host-side mirror of EVREAL's model plugin classes (``model/model.py``, ``model/legacy.py``).

Same class names, constructor arguments, attributes (``num_encoders``,
``num_bins``, ``states``, ``prev_recs``) and methods (``reset_states``,
``forward -> {'image': ...}``, ``load_state_dict``, ``to``, ``eval``,
``parameters``) as the reference, so ``eval.py:124-158``'s factory and
``eval.py:109-115``'s ``load_model`` work on them unchanged.  The arithmetic
runs in the CUDA library through the C ABI (``evk_model_*``); these classes
hold no torch parameters and there is no CPU path.
"""
import ctypes

import torch

from . import _lib

ARCH_UNET, ARCH_FIRENET_LEGACY, ARCH_FIRENET, ARCH_SPADE, ARCH_ETNET = 0, 1, 2, 3, 4


class _NativeModel:
    _prefix = ''
    _arch = ARCH_UNET

    def __init__(self):
        self._sd = None
        self._handle = None
        self._key = None
        self._cache = {}          # key -> handle: one device program per (batch, H, W, device, precision) seen so far
        self._device = None
        self.precision = 0        # 0: tensor-core split-bf16 where applicable, 1: fp32 SIMT everywhere

    # ---- nn.Module-like surface used by eval.py:109-115
    def load_state_dict(self, state_dict, strict=True):
        sd = {}
        for k, v in state_dict.items():
            if k.endswith('num_batches_tracked'):
                continue
            if self._prefix:
                if not k.startswith(self._prefix):
                    if strict:
                        raise RuntimeError("Unexpected key(s) in state_dict: \"%s\"" % k)
                    continue
                k = k[len(self._prefix):]
            sd[k] = v.detach().to(device='cpu', dtype=torch.float32).contiguous()
        self._sd = sd
        self._release()
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.EvkError("evreal_b200 models run on CUDA only (got device %s)" % device)
        self._device = device if device.index is not None else torch.device('cuda', torch.cuda.current_device())
        return self

    def cuda(self, device=None):
        return self.to(torch.device('cuda', torch.cuda.current_device() if device is None else device))

    def eval(self):
        return self

    def parameters(self):
        return iter(())

    def state_dict(self):
        return {self._prefix + k: v for k, v in (self._sd or {}).items()}

    # ---- handle management
    def _config(self, batch, height, width):
        raise NotImplementedError

    def clone(self):
        """A second model of the same weights with its own device programs and recurrent state (ColorNet)."""
        import copy
        c = copy.copy(self)
        c._cache, c._handle, c._key = {}, None, None
        return c

    MAX_HANDLES = 4             # device programs kept per model (ColorNet alternates between two shapes every frame)

    def _release(self):
        for h in self._cache.values():
            _lib.load().evk_model_destroy(h)
        self._cache = {}
        self._handle = None
        self._key = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure(self, batch, height, width, device):
        key = (batch, height, width, device, self.precision)
        if self._handle is not None and self._key == key:
            return
        if key in self._cache:          # a shape seen before: its program and its recurrent state are still there
            self._handle, self._key = self._cache[key], key
            return
        if self._sd is None:
            raise _lib.EvkError("load_state_dict() must be called before the first forward")
        lib = _lib.load()
        while len(self._cache) >= self.MAX_HANDLES:
            old = next(iter(self._cache))
            lib.evk_model_destroy(self._cache.pop(old))
        cfg = self._config(batch, height, width)
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.evk_model_create(ctypes.byref(cfg), ctypes.byref(handle)))
            try:
                for name, t in self._sd.items():
                    shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                    _lib.check(lib.evk_model_load_tensor(handle, name.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim()))
                _lib.check(lib.evk_model_finalize(handle, _lib.stream_ptr(device)))
                _lib.check(lib.evk_model_reset_states(handle, _lib.stream_ptr(device)))
            except Exception:
                lib.evk_model_destroy(handle)
                raise
        self._cache[key] = handle
        self._handle, self._key = handle, key

    # ---- reference API
    def reset_states(self):
        for key, handle in self._cache.items():
            dev = key[3]
            with torch.cuda.device(dev):
                _lib.check(_lib.load().evk_model_reset_states(handle, _lib.stream_ptr(dev)))

    def forward(self, event_tensor):
        _lib.require_cuda()
        x = event_tensor
        if not x.is_cuda:
            x = x.to(self._device or torch.device('cuda', torch.cuda.current_device()), non_blocking=True)
        x = x.float().contiguous()
        assert x.dim() == 4, "expected N x num_bins x H x W"
        N, C, H, W = x.shape
        assert C == self.num_bins, "expected %d bins, got %d" % (self.num_bins, C)
        self._ensure(N, H, W, x.device)
        out = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().evk_model_forward(self._handle, _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr(x.device)))
        return {'image': out}

    __call__ = forward

    def forward_into(self, event_tensor, out):
        """forward() into a caller-owned [N,1,H,W] CUDA buffer (no allocation; the streaming pipeline's path)."""
        x = event_tensor
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4
        N, C, H, W = x.shape
        assert C == self.num_bins and tuple(out.shape) == (N, 1, H, W) and out.is_contiguous()
        self._ensure(N, H, W, x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().evk_model_forward(self._handle, _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr(x.device)))
        if hasattr(self, 'prev_recs'):
            self.prev_recs = out
        return out

    def io_buffers(self):
        """(input, output) device addresses of the handle-owned buffers the NEXT forward uses (double-buffered by parity)."""
        i, o = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.load().evk_model_io_buffers(self._handle, ctypes.byref(i), ctypes.byref(o)))
        return i.value, o.value

    def forward_raw(self, in_ptr, out_ptr):
        """evk_model_forward on raw device addresses (the streaming pipeline: handle-owned buffers, no copies), on the
        current stream of the handle's device."""
        dev = self._key[3]
        _lib.check(_lib.load().evk_model_forward(self._handle, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr),
                                                 _lib.stream_ptr(dev)))

    # ---- introspection used by bench / tests
    def flops_per_forward(self):
        return float(_lib.load().evk_model_flops(self._handle)) if self._handle else 0.0

    def profile_forward(self, event_tensor):
        """One eager forward with CUDA events around every launch -> [(description, ms, flops), ...]."""
        x = event_tensor.float().contiguous()
        N, C, H, W = x.shape
        self._ensure(N, H, W, x.device)
        lib = _lib.load()
        cap = 256
        ms = (ctypes.c_float * cap)()
        fl = (ctypes.c_double * cap)()
        n = ctypes.c_int(0)
        out = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(lib.evk_model_profile(self._handle, _lib.ptr(x), _lib.ptr(out), _lib.stream_ptr(x.device), cap,
                                             ms, fl, ctypes.byref(n)))
        rows = []
        buf = ctypes.create_string_buffer(160)
        for i in range(min(n.value, cap)):
            _lib.check(lib.evk_model_op_desc(self._handle, i, buf, 160))
            rows.append((buf.value.decode(), float(ms[i]), float(fl[i])))
        return rows

    def op_descriptions(self):
        """Descriptions of the launches of one forward (after the first forward built the device program)."""
        if not self._handle:
            return []
        lib = _lib.load()
        buf = ctypes.create_string_buffer(160)
        rows = []
        for i in range(256):
            if lib.evk_model_op_desc(self._handle, i, buf, 160) != 0:
                break
            rows.append(buf.value.decode())
        return rows

    def last_launch_count(self):
        return int(_lib.load().evk_model_last_launch_count(self._handle)) if self._handle else 0

    def _get_states(self):
        if self._handle is None:
            return None
        lib = _lib.load()
        dev = self._key[3]
        out = []
        with torch.cuda.device(dev):
            for i in range(lib.evk_model_num_states(self._handle)):
                shp = (ctypes.c_int64 * 4)()
                _lib.check(lib.evk_model_state_shape(self._handle, i, shp))
                t = torch.empty(tuple(shp), dtype=torch.float32, device=dev)
                _lib.check(lib.evk_model_get_state(self._handle, i, _lib.ptr(t), _lib.stream_ptr(dev)))
                out.append(t)
        return out

    def _set_states(self, tensors):
        lib = _lib.load()
        dev = self._key[3]
        with torch.cuda.device(dev):
            for i, t in enumerate(tensors):
                t = t.to(device=dev, dtype=torch.float32).contiguous()
                _lib.check(lib.evk_model_set_state(self._handle, i, _lib.ptr(t), _lib.stream_ptr(dev)))


class _UNetFamily(_NativeModel):
    _arch = ARCH_UNET

    def __init__(self, unet_kwargs):
        super().__init__()
        kw = dict(unet_kwargs)
        self.num_bins = kw['num_bins']            # legacy attribute names of the reference
        self.num_encoders = kw['num_encoders']
        self._base = kw.get('base_num_channels', 32)
        self._num_res = kw.get('num_residual_blocks', 2)
        self._k = kw.get('kernel_size', 5)
        self._num_out = kw.get('num_output_channels', 1)
        fa = kw.get('final_activation', None)
        if fa not in (None, '', 'none', 'sigmoid') and hasattr(torch, str(fa)):
            # the reference applies getattr(torch, name, None) (model/unet.py:95-96,137-138: unknown names mean "no
            # activation"); only the activations of shipped checkpoints are built, and a real torch function must not
            # silently run as the identity
            raise _lib.EvkError("final_activation=%r is not built (supported: none, sigmoid)" % (fa,))
        self._sigmoid = fa == 'sigmoid'
        if kw.get('norm', None) not in (None, 'none', 'BN'):
            raise _lib.EvkError("norm=%r is not built (eval-mode BatchNorm is folded into the convolutions; InstanceNorm "
                                "has no running statistics to fold)" % (kw.get('norm'),))
        self._dynamic = bool(kw.get('use_dynamic_decoder', False))
        if kw.get('skip_type', 'sum') != 'sum':
            raise NameError("name 'skip_%s' is not defined" % kw.get('skip_type'))     # model/unet.py:31
        if kw.get('recurrent_block_type', 'convlstm') != 'convlstm':
            raise _lib.EvkError("UNetRecurrent with recurrent_block_type=%r is not built (all shipped checkpoints use "
                                "convlstm)" % kw.get('recurrent_block_type'))
        if kw.get('channel_multiplier', 2) != 2:
            raise _lib.EvkError("channel_multiplier != 2 is not built")

    def _config(self, batch, height, width):
        return _lib.ModelConfig(arch=ARCH_UNET, num_bins=self.num_bins, base_channels=self._base,
                                num_encoders=self.num_encoders, num_residual_blocks=self._num_res,
                                kernel_size=self._k, num_output_channels=self._num_out,
                                final_sigmoid=int(self._sigmoid), dynamic_decoder=int(self._dynamic), batch=batch,
                                height=height, width=width, precision=self.precision)

    @property
    def states(self):
        """[(hidden, cell), ...] per encoder, NCHW copies (model/model.py:116-118)."""
        s = self._get_states()
        if s is None:
            return [None] * self.num_encoders
        return [(s[2 * i], s[2 * i + 1]) for i in range(self.num_encoders)]

    @states.setter
    def states(self, states):
        if self._handle is None:
            if all(s is None for s in states):
                return
            raise _lib.EvkError("states can only be set after the first forward fixed the tensor shapes")
        if all(s is None for s in states):
            _NativeModel.reset_states(self)
            return
        flat = []
        for h, c in states:
            flat += [h, c]
        self._set_states(flat)


class E2VIDRecurrent(_UNetFamily):
    """model/model.py:108-144 (E2VID, SSL-E2VID, HyperE2VID checkpoints)."""
    _prefix = 'unetrecurrent.'

    def __init__(self, unet_kwargs):
        super().__init__(unet_kwargs)
        self.prev_recs = None

    def reset_states(self):
        super().reset_states()
        self.prev_recs = None

    def forward(self, event_tensor):
        out = super().forward(event_tensor)
        self.prev_recs = out['image']
        return out

    __call__ = forward


class FlowNet(_UNetFamily):
    """model/model.py:14-43 (E2VID+ checkpoint).  Only the image channel is computed:
    eval.py never reads 'flow'."""
    _prefix = 'unetflow.'


class _FireNetBase(_NativeModel):
    def _fire_config(self, arch, batch, height, width):
        return _lib.ModelConfig(arch=arch, num_bins=self.num_bins, base_channels=self._base, num_encoders=0,
                                num_residual_blocks=2, kernel_size=self._k, num_output_channels=1, final_sigmoid=0,
                                dynamic_decoder=0, batch=batch, height=height, width=width, precision=self.precision)

    @property
    def states(self):
        s = self._get_states()
        return [None, None] if s is None else s

    @states.setter
    def states(self, states):
        if self._handle is None or all(s is None for s in states):
            _NativeModel.reset_states(self)
            return
        self._set_states(list(states))


class FireNet_legacy(_FireNetBase):
    """model/legacy.py:155-187 (pretrained/FireNet)."""
    _prefix = 'net.'
    _arch = ARCH_FIRENET_LEGACY

    def __init__(self, config={}, unet_kwargs={}):
        super().__init__()
        if unet_kwargs:
            config = unet_kwargs
        assert 'num_bins' in config
        self.num_bins = int(config['num_bins'])
        self.num_encoders = int(config.get('num_encoders', 4))            # legacy.py:127-130
        self._base = int(config.get('base_num_channels', 32))
        self._k = int(config.get('kernel_size', 5))
        if str(config.get('recurrent_block_type', 'convgru')) != 'convgru':
            raise _lib.EvkError("FireNet_legacy with convlstm is not built")
        if config.get('norm', None) not in (None, 'none', 'None', 'BN'):
            raise _lib.EvkError("FireNet_legacy with norm=%r is not built" % (config.get('norm'),))
        if int(config.get('num_residual_blocks', 2)) != 2 or config.get('recurrent_blocks', {'resblock': [0]}) != {'resblock': [0]}:
            raise _lib.EvkError("only the shipped FireNet topology (2 residual blocks, recurrent resblock 0) is built")
        self.num_recurrent_units = 2

    def _config(self, batch, height, width):
        return self._fire_config(ARCH_FIRENET_LEGACY, batch, height, width)


class FireNet(_FireNetBase):
    """model/model.py:147-190 (pretrained/FireNet+)."""
    _prefix = ''
    _arch = ARCH_FIRENET

    def __init__(self, num_bins=5, base_num_channels=16, kernel_size=3):
        super().__init__()
        self.num_bins = num_bins
        self._base = base_num_channels
        self._k = kernel_size
        self.num_recurrent_units = 2

    def _config(self, batch, height, width):
        return self._fire_config(ARCH_FIRENET, batch, height, width)


class SpadeE2vid(_NativeModel):
    """model/spade_e2v.py:113-179 (Unet6, pretrained/SPADE-E2VID; `model.SpadeE2vid` in the reference's model/__init__.py).
    The class has no constructor arguments; eval.py:130-133 sets ``num_encoders = 3`` on the instance.  ``states`` are the four
    ConvLSTM (hidden, cell) pairs; the previous 3-channel reconstruction that conditions the SPADE layers lives in the device
    program and is cleared by ``reset_states``."""
    _prefix = ''
    _arch = ARCH_SPADE

    def __init__(self):
        super().__init__()
        self.num_bins = 5
        self.num_encoders = 3
        self.prev_recs = None

    def _config(self, batch, height, width):
        return _lib.ModelConfig(arch=ARCH_SPADE, num_bins=5, base_channels=32, num_encoders=3, num_residual_blocks=2, kernel_size=5,
                                num_output_channels=3, final_sigmoid=1, dynamic_decoder=0, batch=batch, height=height, width=width,
                                precision=self.precision)

    def reset_states(self):
        super().reset_states()
        self.prev_recs = None

    @property
    def states(self):
        s = self._get_states()
        if s is None:
            return None
        return [(s[2 * i], s[2 * i + 1]) for i in range(4)]

    @states.setter
    def states(self, states):
        if states is None or self._handle is None:
            _NativeModel.reset_states(self)
            return
        flat = []
        for h, c in states:
            flat += [h, c]
        self._set_states(flat)


class EITR(_NativeModel):
    """model/eitr/eitr.py:4-16 over model/eitr/u_trans.py:13-123 (mls_tpa, pretrained/ET-Net): ``EITR(eitr_kwargs)`` with
    ``num_bins`` and ``norm``; eval.py:149-150 sets ``num_encoders = 3`` on the instance.  ``states`` are the three ConvLSTM
    (hidden, cell) pairs of the recurrent encoders.  Only norm = None (the shipped checkpoint) is built."""
    _prefix = ''
    _arch = ARCH_ETNET

    def __init__(self, eitr_kwargs):
        super().__init__()
        self.num_bins = int(eitr_kwargs['num_bins'])
        if eitr_kwargs.get('norm', None) not in (None, 'none', 'None'):
            raise _lib.EvkError("ET-Net with norm=%r is not built (the shipped checkpoint has none)" % (eitr_kwargs.get('norm'),))
        self.num_encoders = 3

    def _config(self, batch, height, width):
        return _lib.ModelConfig(arch=ARCH_ETNET, num_bins=self.num_bins, base_channels=32, num_encoders=3, num_residual_blocks=0,
                                kernel_size=5, num_output_channels=1, final_sigmoid=1, dynamic_decoder=0, batch=batch, height=height,
                                width=width, precision=self.precision)

    @property
    def states(self):
        s = self._get_states()
        if s is None:
            return [None] * 3
        return [(s[2 * i], s[2 * i + 1]) for i in range(3)]

    @states.setter
    def states(self, states):
        if self._handle is None or all(s is None for s in states):
            _NativeModel.reset_states(self)
            return
        flat = []
        for h, c in states:
            flat += [h, c]
        self._set_states(flat)


# ---------------------------------------------------------------------------------------------- colour (CED)
_BAYER = (('R', 0, 0), ('G', 0, 1), ('B', 1, 1), ('W', 1, 0))       # channel -> (row offset, column offset) of its 2x2 site


def _shift_replicate(img, dx, dy):
    """Move an image by (dx, dy) >= 0 pixels, filling the vacated border from the first valid line (utils/color_utils.py:5-17)."""
    import numpy as np
    out = np.roll(np.roll(img, dy, axis=0), dx, axis=1)
    if dy > 0:
        out[:dy, :] = out[dy, :][None, :]
    if dx > 0:
        out[:, :dx] = out[:, dx][:, None]
    return out


def merge_channels_into_color_image(channels):
    """Half-resolution R/G/B/W reconstructions + the full-resolution grey one -> full-resolution BGR uint8 image: every
    colour plane is doubled bilinearly and aligned to the R site, green is the mean of G and W, and the LAB lightness of
    the result is replaced by the grey reconstruction (utils/color_utils.py:52-88).  Output tooling on the host (cv2), not
    part of the measured path."""
    import cv2
    import numpy as np
    up = {k: cv2.resize(channels[k], dsize=None, fx=2, fy=2, interpolation=cv2.INTER_LINEAR) for k in ('R', 'G', 'B', 'W')}
    blue = _shift_replicate(up['B'], 1, 1)
    green = cv2.addWeighted(src1=_shift_replicate(up['G'], 1, 0), alpha=0.5, src2=_shift_replicate(up['W'], 0, 1), beta=0.5,
                            gamma=0.0, dtype=cv2.CV_8U)
    lab = cv2.cvtColor(src=np.dstack([blue, green, up['R']]), code=cv2.COLOR_BGR2LAB)
    lab[:, :, 0] = channels['grayscale']
    return cv2.cvtColor(src=lab, code=cv2.COLOR_LAB2BGR)


class ColorNet:
    """model/model.py:46-105: the events of a Bayer-pattern sensor (CED) split into R/G/B/W sites, each reconstructed by
    the wrapped recurrent model with its own state, plus a full-resolution grey reconstruction; the five images are merged
    into one colour frame.

    The reference runs five batch-1 forwards per frame and swaps ``model.states`` / ``model.prev_recs`` in and out around
    each.  Recurrent state here is per SAMPLE of a device program, so the four half-resolution sites are ONE batch-4
    forward of a second program of the same weights and the grey image a batch-1 forward of the wrapped model: no state
    swapping, two launches of the network per frame instead of five."""

    def __init__(self, model):
        self.model = model
        self.half = model.clone()
        self.reset_states()

    def reset_states(self):
        self.model.reset_states()
        self.half.reset_states()

    @property
    def num_encoders(self):
        return self.model.num_encoders

    def eval(self):
        return self

    def to(self, device):
        self.model.to(device)
        self.half.to(device)
        return self

    def forward(self, event_tensor):
        from .util import CropParameters
        x = event_tensor
        assert x.dim() == 4 and x.shape[0] == 1, "ColorNet expects one 1 x num_bins x H x W event tensor"
        if not x.is_cuda:
            x = x.to(self.model._device or torch.device('cuda', torch.cuda.current_device()), non_blocking=True)
        height, width = int(x.shape[-2]), int(x.shape[-1])
        assert height % 2 == 0 and width % 2 == 0, "Bayer split needs even sensor dimensions"
        crop_half = CropParameters(width // 2, height // 2, self.model.num_encoders)
        crop_full = CropParameters(width, height, self.model.num_encoders)
        sites = torch.cat([x[:, :, r::2, c::2] for _, r, c in _BAYER], dim=0)          # [4, bins, H/2, W/2]
        half = crop_half.crop(self.half(crop_half.pad(sites))['image'])               # [4, 1, H/2, W/2]
        full = crop_full.crop(self.model(crop_full.pad(x))['image'])                  # [1, 1, H, W]
        q = lambda t: (t * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()           # np.clip(img * 255, 0, 255).astype(uint8)
        half8, full8 = q(half), q(full)
        channels = {name: half8[i, 0] for i, (name, _, _) in enumerate(_BAYER)}
        channels['grayscale'] = full8[0, 0]
        bgr = merge_channels_into_color_image(channels)                               # H x W x 3 uint8
        return {'image': torch.from_numpy(bgr).permute(2, 0, 1).float().div(255)}     # transforms.functional.to_tensor

    __call__ = forward
