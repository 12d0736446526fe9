"""CPU probe (build container only: needs /root/reference/pretrained): error of the split-bf16 x3 arithmetic on the REAL E2VID
decoder weights, phase-stacked form (composite 5x5 weights on the replicate-padded low-resolution map, poly.cu) against the
plain form (bilinear x2, then 5x5), both emulated with three bf16 products accumulated in fp32 and compared with float64.
    python tools/poly_numerics_probe.py"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("EVREAL_REFERENCE", "/root/reference")


def split(t):
    hi = t.to(torch.bfloat16).to(torch.float32)
    lo = (t - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def conv3(x, w, **kw):
    """hi*hi + lo*hi + hi*lo in fp32 (what the tensor-core kernel accumulates)."""
    xh, xl = split(x)
    wh, wl = split(w)
    return F.conv2d(xh, wh, **kw) + F.conv2d(xl, wh, **kw) + F.conv2d(xh, wl, **kw)


def phase_coef(a):
    c = np.zeros((5, 5))
    for d in range(-2, 3):
        s = a + d
        m, r = (s + 4) // 2 - 2, (s + 4) & 1
        if r == 0:
            c[d + 2][m + 1] += 0.25; c[d + 2][m + 2] += 0.75
        else:
            c[d + 2][m + 2] += 0.75; c[d + 2][m + 3] += 0.25
    return torch.tensor(c)


def main():
    ck = torch.load(os.path.join(REF, 'pretrained/E2VID/model.pth'), map_location='cpu', weights_only=False)
    sd = ck['state_dict']
    torch.manual_seed(0)
    for i, (H, W) in ((1, (46, 60)), (2, (92, 120))):
        w = sd['unetrecurrent.decoders.%d.conv2d.weight' % i].double()
        g, b = sd['unetrecurrent.decoders.%d.norm_layer.weight' % i].double(), sd['unetrecurrent.decoders.%d.norm_layer.running_var' % i].double()
        w = w * (g / torch.sqrt(b + 1e-5)).view(-1, 1, 1, 1)                      # eval-mode BatchNorm folded like pack_conv
        x = torch.relu(torch.randn(1, w.shape[1], H, W, dtype=torch.float64)) * 0.7
        ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False), w, padding=2)
        scale = float(ref.abs().max())
        plain = conv3(F.interpolate(x.float(), scale_factor=2, mode='bilinear', align_corners=False), w.float(), padding=2)
        xp = F.pad(x, (2, 2, 2, 2), mode='replicate').float()
        out = torch.zeros_like(plain)
        for a in range(2):
            for bb in range(2):
                wc = torch.einsum('oiyx,yt,xs->oits', w, phase_coef(a), phase_coef(bb)).float()
                out[:, :, a::2, bb::2] = conv3(xp, wc)
        inner = (slice(None), slice(None), slice(2, -2), slice(2, -2))          # the border ring gets the (equally accurate) correction
        print('decoder %d (%d -> %d ch @%dx%d): max|ref| %.3f   plain x3 err %.2e   phase-stacked x3 err %.2e   (relative to max|ref|)' %
              (i, w.shape[1], w.shape[0], 2 * H, 2 * W, scale, float((plain.double() - ref)[inner].abs().max()) / scale,
               float((out.double() - ref)[inner].abs().max()) / scale))


if __name__ == '__main__':
    main()
