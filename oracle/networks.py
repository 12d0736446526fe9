"""Oracle: the recurrent reconstruction networks (stage 2), torch fp32 on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A *functional* restatement: weights are a flat ``{name: tensor}`` dict that uses
the reference's ``state_dict`` names with the wrapper prefix (``unetrecurrent.``,
``unetflow.``, ``net.``) stripped, so any shipped checkpoint -- or a seeded
random one of the same shapes -- can be fed to both this oracle and the CUDA
runner.  Convolutions go through ``torch.nn.functional`` (the same ATen/oneDNN
kernels the reference's ``nn.Module``s dispatch to), which is what makes this the
fair CPU baseline as well.

Follows, in /root/reference:
  model/submodules.py:8-35     ConvLayer (conv [+ eval BN] [+ activation])
  model/submodules.py:69-97    UpsampleConvLayer (bilinear x2, align_corners=False, then conv)
  model/submodules.py:100-127  DynamicUpsampleLayer
  model/submodules.py:152-184  ResidualBlock
  model/submodules.py:187-245  ConvLSTM (gate order: in, remember, out, cell)
  model/submodules.py:248-287  ConvGRU
  model/unet.py:85-143         UNetRecurrent.forward
  model/model.py:108-144       E2VIDRecurrent (prev_recs handling), :14-43 FlowNet
  model/model.py:147-190       FireNet (FireNet+ checkpoint)
  model/legacy.py:32-111,155+  UNetFire / FireNet_legacy (FireNet checkpoint)
  model/hyper/hyper_dynamic.py:7-92  context fusion, atom generation, dynamic conv
  model/eitr/*.py              ET-Net (u_trans.py:13-123 mls_tpa, transformer_encoder.py, transformer_decoder.py, position_encoding.py)
  model/spade_e2v.py:7-179     SPADE-E2VID (Unet6: stride-1 recurrent encoder at full resolution, pixel-shuffle decoders with
                               SPADE normalisation conditioned on the previous reconstruction, recurrent last decoder)
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _bn(w, pfx, x):
    """Eval-mode BatchNorm2d with running statistics."""
    return F.batch_norm(x, w[pfx + '.running_mean'], w[pfx + '.running_var'],
                        w[pfx + '.weight'], w[pfx + '.bias'], False, 0.0, BN_EPS)


def _has_bn(w, pfx):
    return (pfx + '.running_mean') in w


def conv_layer(w, pfx, x, stride=1, padding=0, relu=True):
    """submodules.py:8-35.  BN presence is read off the weight dict."""
    y = F.conv2d(x, w[pfx + '.conv2d.weight'], w.get(pfx + '.conv2d.bias'), stride, padding)
    if _has_bn(w, pfx + '.norm_layer'):
        y = _bn(w, pfx + '.norm_layer', y)
    return torch.relu(y) if relu else y


def conv_lstm(w, pfx, x, state):
    """submodules.py:187-245."""
    if state is None:
        z = torch.zeros_like(x)
        state = (z, z)
    h_prev, c_prev = state
    g = F.conv2d(torch.cat((x, h_prev), 1), w[pfx + '.Gates.weight'], w[pfx + '.Gates.bias'], 1, 1)
    i, f, o, c = g.chunk(4, 1)
    i, f, o, c = torch.sigmoid(i), torch.sigmoid(f), torch.sigmoid(o), torch.tanh(c)
    cell = f * c_prev + i * c
    hidden = o * torch.tanh(cell)
    return hidden, cell


def conv_gru(w, pfx, x, state):
    """submodules.py:248-287."""
    if state is None:
        state = torch.zeros_like(x)
    xs = torch.cat([x, state], 1)
    update = torch.sigmoid(F.conv2d(xs, w[pfx + '.update_gate.weight'], w[pfx + '.update_gate.bias'], 1, 1))
    reset = torch.sigmoid(F.conv2d(xs, w[pfx + '.reset_gate.weight'], w[pfx + '.reset_gate.bias'], 1, 1))
    cand = torch.tanh(F.conv2d(torch.cat([x, state * reset], 1),
                               w[pfx + '.out_gate.weight'], w[pfx + '.out_gate.bias'], 1, 1))
    return state * (1 - update) + cand * update


def residual_block(w, pfx, x):
    """submodules.py:152-184."""
    y = F.conv2d(x, w[pfx + '.conv1.weight'], w.get(pfx + '.conv1.bias'), 1, 1)
    if _has_bn(w, pfx + '.bn1'):
        y = _bn(w, pfx + '.bn1', y)
    y = torch.relu(y)
    y = F.conv2d(y, w[pfx + '.conv2.weight'], w.get(pfx + '.conv2.bias'), 1, 1)
    if _has_bn(w, pfx + '.bn2'):
        y = _bn(w, pfx + '.bn2', y)
    return torch.relu(y + x)


def upsample_conv(w, pfx, x, k):
    """submodules.py:69-97."""
    xu = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    return conv_layer(w, pfx, xu, 1, k // 2, relu=True)


def transposed_conv(w, pfx, x, k):
    """submodules.py:38-66 (ConvTranspose2d stride 2, output_padding 1)."""
    y = F.conv_transpose2d(x, w[pfx + '.transposed_conv2d.weight'], w.get(pfx + '.transposed_conv2d.bias'),
                           stride=2, padding=k // 2, output_padding=1)
    if _has_bn(w, pfx + '.norm_layer'):
        y = _bn(w, pfx + '.norm_layer', y)
    return torch.relu(y)


def dynamic_upsample(w, pfx, x, ev, prev, k=5, num_atoms=6):
    """submodules.py:100-127 + hyper_dynamic.py:7-92."""
    xu = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    ctx = torch.cat((ev, prev), 1)
    ctx = F.interpolate(ctx, scale_factor=0.25, mode='bilinear', align_corners=False)
    ctx = F.conv2d(ctx, w[pfx + '.context_fusion.conv.weight'], w[pfx + '.context_fusion.conv.bias'], 1, 1)
    g = pfx + '.dynamic_atom_generation'
    c = F.conv2d(ctx, w[g + '.bases_net.0.weight'], w[g + '.bases_net.0.bias'], 1, 1)
    c = torch.tanh(_bn(w, g + '.bases_net.1', c))
    c = F.conv2d(c, w[g + '.bases_net.3.weight'], w[g + '.bases_net.3.bias'], 1, 1)
    c = torch.tanh(_bn(w, g + '.bases_net.4', c))
    N, _, H, W = c.shape
    bases = w[g + '.bases']                                  # [K, k*k]
    c = c.view(N, num_atoms, bases.shape[0], H, W)
    atoms = torch.einsum('bmkhw,kl->bmlhw', c, bases)        # [N, atoms, k*k, H, W]
    C = xu.shape[1]
    cols = F.unfold(xu, kernel_size=k, stride=1, padding=k // 2).view(N, C, k * k, H, W)
    inter = torch.einsum('bmlhw,bclhw->bcmhw', atoms, cols).reshape(N, C * num_atoms, H, W)
    y = F.conv2d(inter, w[pfx + '.dynamic_conv.compositional_coefficients'], w[pfx + '.dynamic_conv.bias'])
    return torch.relu(y)


class UNetRecurrentOracle:
    """E2VID / E2VID+ / SSL-E2VID / HyperE2VID topology (unet.py:107-143)."""

    def __init__(self, weights, num_encoders=3, num_residual_blocks=2, kernel_size=5,
                 final_sigmoid=False, dynamic_decoder=False, upsample_conv_decoder=True):
        self.w = weights
        self.num_encoders = num_encoders
        self.num_residual_blocks = num_residual_blocks
        self.k = kernel_size
        self.final_sigmoid = final_sigmoid
        self.dynamic_decoder = dynamic_decoder
        self.upsample_conv_decoder = upsample_conv_decoder
        self.reset_states()

    def reset_states(self):
        self.states = [None] * self.num_encoders
        self.prev_recs = None

    @torch.no_grad()
    def __call__(self, x):
        w, k = self.w, self.k
        if self.prev_recs is None:                       # model.py:139-141
            self.prev_recs = torch.zeros(x.shape[0], 1, x.shape[2], x.shape[3])
        ev = x
        x = conv_layer(w, 'head', x, 1, k // 2)
        head = x
        blocks = []
        for i in range(self.num_encoders):
            x = conv_layer(w, 'encoders.%d.conv' % i, x, 2, k // 2)
            self.states[i] = conv_lstm(w, 'encoders.%d.recurrent_block' % i, x, self.states[i])
            x = self.states[i][0]
            blocks.append(x)
        for j in range(self.num_residual_blocks):
            x = residual_block(w, 'resblocks.%d' % j, x)
        for i in range(self.num_encoders):
            x = x + blocks[self.num_encoders - i - 1]
            if i == 0 and self.dynamic_decoder:
                x = dynamic_upsample(w, 'decoders.0', x, ev, self.prev_recs, k)
            elif self.upsample_conv_decoder:
                x = upsample_conv(w, 'decoders.%d' % i, x, k)
            else:
                x = transposed_conv(w, 'decoders.%d' % i, x, k)
        img = conv_layer(w, 'pred', x + head, 1, 0, relu=False)
        if self.final_sigmoid:
            img = torch.sigmoid(img)
        out = img[:, 0:1]
        self.prev_recs = out
        return out


class FireNetLegacyOracle:
    """pretrained/FireNet: legacy.py:79-111 -- head(conv+GRU), resblock0(+GRU), resblock1, pred."""
    num_encoders = 4          # legacy.py:127-130 default when the config has no num_encoders

    def __init__(self, weights):
        self.w = weights
        self.reset_states()

    def reset_states(self):
        self.states = [None, None]

    @torch.no_grad()
    def __call__(self, x):
        w = self.w
        x = conv_layer(w, 'head.conv', x, 1, 1)
        x = self.states[0] = conv_gru(w, 'head.recurrent_block', x, self.states[0])
        x = residual_block(w, 'resblocks.0.conv', x)
        x = self.states[1] = conv_gru(w, 'resblocks.0.recurrent_block', x, self.states[1])
        x = residual_block(w, 'resblocks.1', x)
        return conv_layer(w, 'pred', x, 1, 0, relu=False)


class FireNetOracle:
    """pretrained/FireNet+: model.py:178-190 -- head, G1, R1, G2, R2, pred."""
    num_encoders = 0          # eval.py:154-155

    def __init__(self, weights):
        self.w = weights
        self.reset_states()

    def reset_states(self):
        self.states = [None, None]

    @torch.no_grad()
    def __call__(self, x):
        w = self.w
        x = conv_layer(w, 'head', x, 1, 1)
        x = self.states[0] = conv_gru(w, 'G1', x, self.states[0])
        x = residual_block(w, 'R1', x)
        x = self.states[1] = conv_gru(w, 'G2', x, self.states[1])
        x = residual_block(w, 'R2', x)
        return conv_layer(w, 'pred', x, 1, 0, relu=False)


class SpadeE2vidOracle:
    """pretrained/SPADE-E2VID: model/spade_e2v.py:113-179 (Unet6).  Per SAMPLE where the reference is per tensor (the
    first-frame min / max of x[:, :3], :140-145): the reference only ever runs batch 1, and batching independent
    sequences must not couple them.  Note the reference's in-place quirk: x_org is a VIEW of the input, so the head
    convolution of the first frame sees the normalised first three bins."""
    num_encoders = 3          # eval.py:131-132

    def __init__(self, weights):
        self.w = weights
        self.reset_states()

    def reset_states(self):
        self.states = None
        self.prev_recs = None

    def _rec(self, pfx, x, state, stride):
        w = self.w
        y = F.conv2d(x, w[pfx + '.conv0.weight'], None, stride, 2)
        y = torch.relu(_bn(w, pfx + '.bn', y))
        st = conv_lstm(w, pfx + '.recurrent_block', y, state)
        return st[0], st

    def _res(self, pfx, x):
        w = self.w
        y = torch.relu(_bn(w, pfx + '.bn1', F.conv2d(x, w[pfx + '.conv1.weight'], None, 1, 1)))
        y = _bn(w, pfx + '.bn2', F.conv2d(y, w[pfx + '.conv2.weight'], None, 1, 1))
        return torch.relu(y + x)

    def _up(self, pfx, x, x_org):
        w = self.w
        y = F.pixel_shuffle(F.conv2d(x, w[pfx + '.conv0.weight'], None, 1, 1), 2)
        n = pfx + '.norm'
        normalized = F.batch_norm(y, w[n + '.param_free_norm.running_mean'], w[n + '.param_free_norm.running_var'], None, None,
                                  False, 0.0, BN_EPS)
        seg = F.interpolate(x_org, size=y.shape[-2:], mode='nearest')
        actv = torch.relu(F.conv2d(seg, w[n + '.mlp_shared.0.weight'], w[n + '.mlp_shared.0.bias'], 1, 1))
        gamma = F.conv2d(actv, w[n + '.mlp_gamma.weight'], w[n + '.mlp_gamma.bias'], 1, 1)
        beta = F.conv2d(actv, w[n + '.mlp_beta.weight'], w[n + '.mlp_beta.bias'], 1, 1)
        return torch.relu(normalized * (1 + gamma) + beta)

    @torch.no_grad()
    def __call__(self, x):
        w = self.w
        x = x.clone()
        prev = self.states if self.states is not None else [None] * 4
        if self.prev_recs is None:
            x_org = x[:, :3]                                   # a view: the head below sees the change
            for n in range(x.shape[0]):
                x_org[n] -= x_org[n].min()
                if x_org[n].max() > 0:
                    x_org[n] /= x_org[n].max()
        else:
            x_org = self.prev_recs
        head = torch.relu(F.conv2d(x, w['fc.weight'], w['fc.bias'], 1, 2))
        x0, s0 = self._rec('rec0', head, prev[0], 1)
        x1, s1 = self._rec('rec1', x0, prev[1], 2)
        x2, s2 = self._rec('rec2', x1, prev[2], 2)
        y = self._res('res1', self._res('res0', x2))
        y = self._up('up0', y + x2, x_org)
        y = self._up('up1', y + x1, x_org)
        y, s3 = self._rec('up2', y + x0, prev[3], 1)
        img = F.conv2d(torch.relu(y + head), w['conv_img.weight'], w['conv_img.bias'])
        img = torch.sigmoid(_bn(w, 'bn_img', img))
        self.states = [s0, s1, s2, s3]
        self.prev_recs = img
        return img.mean(1, keepdim=True)


def sine_position_table(n, d=256):
    """model/eitr/position_encoding.py:15-23 (float64 table, stored as float32)."""
    import numpy as np
    pos = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(d)
    ang = pos / np.power(10000, 2 * (j // 2) / d)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang.astype(np.float32))


def _mha(w, pfx, q_in, kv_in, nhead=8):
    """nn.MultiheadAttention forward (eval, no masks): tokens [L, N, E]; in_proj split q | k | v; q scaled by head_dim^-0.5."""
    L, N, E = q_in.shape
    S = kv_in.shape[0]
    Wi, bi = w[pfx + '.in_proj_weight'], w[pfx + '.in_proj_bias']
    q = F.linear(q_in, Wi[:E], bi[:E])
    k = F.linear(kv_in, Wi[E:2 * E], bi[E:2 * E])
    v = F.linear(kv_in, Wi[2 * E:], bi[2 * E:])
    hd = E // nhead
    q = q.reshape(L, N * nhead, hd).transpose(0, 1) * (hd ** -0.5)
    k = k.reshape(S, N * nhead, hd).transpose(0, 1)
    v = v.reshape(S, N * nhead, hd).transpose(0, 1)
    a = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    o = torch.bmm(a, v).transpose(0, 1).reshape(L, N, E)
    return F.linear(o, w[pfx + '.out_proj.weight'], w[pfx + '.out_proj.bias'])


def _ln(w, pfx, x):
    return F.layer_norm(x, (x.shape[-1],), w[pfx + '.weight'], w[pfx + '.bias'], 1e-5)


def _ffn(w, pfx, x):
    return F.linear(torch.relu(F.linear(x, w[pfx + '.linear1.weight'], w[pfx + '.linear1.bias'])), w[pfx + '.linear2.weight'],
                    w[pfx + '.linear2.bias'])


class ETNetOracle:
    """pretrained/ET-Net (model/eitr/u_trans.py:13-123, transformer_encoder.py, transformer_decoder.py): E2VID's head, three
    recurrent stride-2 encoders and upsample-conv decoders around a multi-scale token path -- the 1/8-resolution feature map
    and 2x2 / 4x4 patch embeddings of the 1/4 and 1/2 maps, each through three pre-norm encoder layers (sine positions added
    once), then two-layer decoders with cross attention to the coarser scale's tokens; the six token sets are averaged."""
    num_encoders = 3          # eval.py:152-153

    def __init__(self, weights):
        self.w = weights
        self.reset_states()

    def reset_states(self):
        self.states = [None] * 3

    def _encoder(self, pfx, src, pos, layers=3):
        w = self.w
        x = src + pos
        for i in range(layers):
            p = '%s.encoder.layers.%d' % (pfx, i)
            n = _ln(w, p + '.norm1', x)
            x = x + _mha(w, p + '.self_attn', n, n)
            x = x + _ffn(w, p, _ln(w, p + '.norm2', x))
        return x

    def _decoder(self, pfx, tgt, memory, layers=2):
        w = self.w
        x = tgt
        for i in range(layers):
            p = '%s.decoder.layers.%d' % (pfx, i)
            n = _ln(w, p + '.norm1', x)
            x = x + _mha(w, p + '.self_attn', n, n)
            x = x + _mha(w, p + '.cross_attn', _ln(w, p + '.norm21', x), _ln(w, p + '.norm22', memory))
            x = x + _ffn(w, p, _ln(w, p + '.norm3', x))
        return x

    @torch.no_grad()
    def __call__(self, x):
        w = self.w
        x = conv_layer(w, 'head', x, 1, 2)
        head = x
        blocks = []
        for i in range(3):
            x = conv_layer(w, 'DownsampleConv.%d.conv' % i, x, 2, 2)
            self.states[i] = conv_lstm(w, 'DownsampleConv.%d.recurrent_block' % i, x, self.states[i])
            x = self.states[i][0]
            blocks.append(x)
        N, _, H, W = head.shape
        tok = lambda t: t.flatten(2).permute(2, 0, 1)                     # [N, 256, h, w] -> [L, N, 256]
        words = [tok(blocks[2]), tok(F.conv2d(blocks[1], w['split1.weight'], w['split1.bias'], 2)),
                 tok(F.conv2d(blocks[0], w['split2.weight'], w['split2.bias'], 4))]
        pos = sine_position_table(words[0].shape[0])[:, None, :]
        hs = [self._encoder('trans_encoder%d' % i, words[i], pos) for i in range(3)]
        hc = [self._decoder('trans_decoder0', hs[0], hs[0]), self._decoder('trans_decoder1', hs[1], hs[0]),
              self._decoder('trans_decoder2', hs[2], hs[1])]
        t = (hs[0] + hs[1] + hs[2] + hc[0] + hc[1] + hc[2]) / 6
        y = t.permute(1, 2, 0).reshape(N, 256, H // 8, W // 8)
        for i in range(3):
            y = upsample_conv(w, 'UpsampleConv.%d' % i, y + blocks[2 - i], 5)
        img = conv_layer(w, 'pred', y + head, 1, 0, relu=False)
        return torch.sigmoid(img)


# ---------------------------------------------------------------------------
# seeded random weights of the shipped architectures (SURVEY A.1 shapes)
# ---------------------------------------------------------------------------
def _conv(g, out_c, in_c, k, bias=True, gain=1.0):
    fan_in = in_c * k * k
    d = {'weight': torch.randn(out_c, in_c, k, k, generator=g) * (gain / fan_in ** 0.5)}
    if bias:
        d['bias'] = torch.randn(out_c, generator=g) * 0.1
    return d


def _bn_params(g, c):
    return {'weight': 1.0 + 0.2 * torch.randn(c, generator=g), 'bias': 0.1 * torch.randn(c, generator=g),
            'running_mean': 0.1 * torch.randn(c, generator=g),
            'running_var': 0.5 + torch.rand(c, generator=g)}


def _put(w, pfx, d):
    for k, v in d.items():
        w[pfx + '.' + k] = v


def random_unet_weights(seed=0, base=32, num_encoders=3, num_res=2, k=5, bins=5, norm_bn=False,
                        num_out=1, dynamic_decoder=False):
    """Weight dict with the reference's names and shapes for the UNetRecurrent family."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    _put(w, 'head.conv2d', _conv(g, base, bins, k))
    cin = base
    for i in range(num_encoders):
        cout = cin * 2
        _put(w, 'encoders.%d.conv.conv2d' % i, _conv(g, cout, cin, k, bias=not norm_bn, gain=1.4))
        if norm_bn:
            _put(w, 'encoders.%d.conv.norm_layer' % i, _bn_params(g, cout))
        _put(w, 'encoders.%d.recurrent_block.Gates' % i, _conv(g, 4 * cout, 2 * cout, 3))
        cin = cout
    for j in range(num_res):
        for n, b in (('conv1', 'bn1'), ('conv2', 'bn2')):
            _put(w, 'resblocks.%d.%s' % (j, n), _conv(g, cin, cin, 3, bias=not norm_bn, gain=1.0))
            if norm_bn:
                _put(w, 'resblocks.%d.%s' % (j, b), _bn_params(g, cin))
    for i in range(num_encoders):
        cout = cin // 2
        if i == 0 and dynamic_decoder:
            p = 'decoders.0'
            _put(w, p + '.context_fusion.conv', _conv(g, 32, bins + 1, 3))
            _put(w, p + '.dynamic_atom_generation.bases_net.0', _conv(g, 64, 32, 3))
            _put(w, p + '.dynamic_atom_generation.bases_net.1', _bn_params(g, 64))
            _put(w, p + '.dynamic_atom_generation.bases_net.3', _conv(g, 72, 64, 3))
            _put(w, p + '.dynamic_atom_generation.bases_net.4', _bn_params(g, 72))
            w[p + '.dynamic_atom_generation.bases'] = torch.randn(12, k * k, generator=g) * 0.3
            w[p + '.dynamic_conv.compositional_coefficients'] = \
                torch.randn(cout, cin * 6, 1, 1, generator=g) * (1.4 / (cin * 6) ** 0.5)
            w[p + '.dynamic_conv.bias'] = torch.randn(cout, generator=g) * 0.1
        else:
            _put(w, 'decoders.%d.conv2d' % i, _conv(g, cout, cin, k, bias=not norm_bn, gain=1.4))
            if norm_bn:
                _put(w, 'decoders.%d.norm_layer' % i, _bn_params(g, cout))
        cin = cout
    _put(w, 'pred.conv2d', _conv(g, num_out, base, 1, bias=not norm_bn))
    if norm_bn:
        _put(w, 'pred.norm_layer', _bn_params(g, num_out))
    return w


def random_firenet_weights(seed=0, base=16, bins=5, legacy=True):
    g = torch.Generator().manual_seed(seed)
    w = {}
    names = (dict(head='head.conv.conv2d', g1='head.recurrent_block', r1='resblocks.0.conv',
                  g2='resblocks.0.recurrent_block', r2='resblocks.1', pred='pred.conv2d') if legacy else
             dict(head='head.conv2d', g1='G1', r1='R1', g2='G2', r2='R2', pred='pred.conv2d'))
    _put(w, names['head'], _conv(g, base, bins, 3))
    for gk in ('g1', 'g2'):
        for gate in ('reset_gate', 'update_gate', 'out_gate'):
            _put(w, names[gk] + '.' + gate, _conv(g, base, 2 * base, 3, gain=1.4))
    for rk in ('r1', 'r2'):
        for c in ('conv1', 'conv2'):
            _put(w, names[rk] + '.' + c, _conv(g, base, base, 3, gain=1.2))
    _put(w, names['pred'], _conv(g, 1, base, 1))
    return w
