"""GPU: the mixed operand decomposition of the tensor-core convolutions (conv.cuh, ConvParams::mixed: one fp16 product + two
fp8 products instead of the three bf16 products) -- which layers take it, that EVK_MIXED=0 restores bf16x3, that both meet the
1e-4 bar against the CPU oracle at the shipped E2VID width, and that recurrent state set from the host lands in the
mixed-format companions."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(env):
    from evreal_b200 import E2VIDRecurrent, synthetic
    old = os.environ.get('EVK_MIXED')
    if env is None:
        os.environ.pop('EVK_MIXED', None)
    else:
        os.environ['EVK_MIXED'] = env
    try:
        m = E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(synthetic.unet_state_dict(3, norm_bn=True)).to('cuda')
        m.reset_states()
        m(torch.zeros(1, 5, 48, 64, device='cuda'))          # builds the device program under this environment
        m.reset_states()
    finally:
        if old is None:
            os.environ.pop('EVK_MIXED', None)
        else:
            os.environ['EVK_MIXED'] = old
    return m


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-6))


def test_mixed_layers_and_bf16x3_fallback_agree_with_the_oracle():
    from evreal_b200 import synthetic
    from oracle import networks as on
    g = torch.Generator().manual_seed(11)
    xs = [torch.randn(1, 5, 48, 64, generator=g) for _ in range(4)]
    mixed, plain = _model(None), _model('0')
    dm, dp = mixed.op_descriptions(), plain.op_descriptions()
    n_mixed = sum('f16+2xf8' in d for d in dm)
    assert n_mixed >= 10, dm                                  # ConvLSTMs, stride-2 encoders 1-2, residual blocks, the three decoders
    assert any('lstm' in d and 'f16+2xf8' in d for d in dm) and any('stacked phases' in d and 'f16+2xf8' in d for d in dm)
    assert not any('f16+2xf8' in d for d in dp) and sum('bf16x3' in d for d in dp) == sum('tcgen05' in d for d in dm)
    sd = {k: v for k, v in synthetic.unet_state_dict(3, norm_bn=True).items()}
    o = on.UNetRecurrentOracle({k[len('unetrecurrent.'):]: v for k, v in sd.items()}, 3, 2, final_sigmoid=True)
    for x in xs:
        want = o(x).numpy()
        a, b = mixed(x.cuda())['image'].cpu().numpy(), plain(x.cuda())['image'].cpu().numpy()
        assert _rel(a, want) <= 1e-4 and _rel(b, want) <= 1e-4, (_rel(a, want), _rel(b, want))
        assert not np.array_equal(a, b)                       # (two different arithmetic paths did run)


def test_states_set_from_the_host_reach_the_mixed_companions():
    g = torch.Generator().manual_seed(12)
    xs = [torch.randn(1, 5, 48, 64, generator=g).cuda() for _ in range(3)]
    m = _model(None)
    for x in xs[:2]:
        m(x)
    st = m.states
    want = m(xs[2])['image'].cpu().numpy()
    m2 = _model(None)
    m2(xs[0])
    m2.states = st
    got = m2(xs[2])['image'].cpu().numpy()
    assert _rel(got, want) <= 2e-6, _rel(got, want)
