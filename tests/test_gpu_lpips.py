"""GPU parity: LPIPS (evk_lpips_* -- backbone on the tensor-core / fp32 convolution kernels, max pooling, normalised
squared feature distance) against oracle/metrics.py::lpips_oracle, a torch-CPU restatement of the public LPIPS v0.1
algorithm (utils/eval_metrics.py:100-156 delegates to pyiqa, which is absent offline together with its weights:
PARITY UNPINNED against pyiqa proper -- seeded weights here).  Tolerance 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _frames(n, H, W, seed):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    ref = np.stack([0.5 + 0.4 * np.sin(xx / (7.0 + i)) * np.cos(yy / (5.0 + i)) for i in range(n)]).astype(np.float32)
    img = np.clip(ref + g.normal(0, 0.08, ref.shape), 0, 1).astype(np.float32)
    return img, ref


@pytest.mark.parametrize('net,H,W', [('alex', 180, 240), ('alex', 97, 131), ('vgg', 64, 80)])
@pytest.mark.parametrize('precision', [0, 1])
def test_lpips_matches_oracle(net, H, W, precision):
    from evreal_b200.lpips import LpipsNet, state_dict_from_conv_list
    from oracle import metrics as om
    w = om.random_lpips_weights(net, seed=3)
    img, ref = _frames(4, H, W, seed=11)
    want = om.lpips_oracle(torch.from_numpy(img)[:, None].repeat(1, 3, 1, 1), torch.from_numpy(ref)[:, None].repeat(1, 3, 1, 1),
                           w, net).double().numpy()
    backbone = 0 if net == 'alex' else 1
    m = LpipsNet(backbone, state_dict_from_conv_list(w, backbone), H, W, batch=4, precision=precision)
    if precision == 0:
        assert m.num_tensor_core_layers >= (4 if net == 'alex' else 12)      # everything but the Cin=3 stem
    got = m(torch.from_numpy(img).cuda(), torch.from_numpy(ref).cuda()).cpu().numpy()
    assert got.shape == (4,)
    assert np.all(want > 1e-3)
    assert np.allclose(got, want, rtol=1e-4, atol=0), (got, want)
    # partial batch (finish_queue of a 3-frame tail) and run-to-run determinism
    got3 = m(torch.from_numpy(img[:3]).cuda(), torch.from_numpy(ref[:3]).cuda()).cpu().numpy()
    assert np.array_equal(got3, got[:3])
    # identical images -> 0
    same = m(torch.from_numpy(ref).cuda(), torch.from_numpy(ref).cuda()).cpu().numpy()
    assert np.all(np.abs(same) < 1e-12)


def test_lpips_metric_queue_contract():
    """Queue of 4 + finish_queue flush, like the class PyIqaMetricFactory builds (utils/eval_metrics.py:118-156)."""
    from evreal_b200.eval_metrics import create_metric
    from evreal_b200.lpips import LpipsMetric, state_dict_from_conv_list
    from evreal_b200 import _lib
    from oracle import metrics as om
    w = om.random_lpips_weights('alex', seed=5)
    img, ref = _frames(6, 90, 120, seed=2)
    m = LpipsMetric('lpips', state_dict=state_dict_from_conv_list(w, 0))
    for i in range(6):
        m.update(img[i], ref[i])
        assert m.get_num_updated() == (4 if i == 3 else 0)
    assert m.get_num_scores() == 4
    m.finish_queue()
    assert m.get_num_updated() == 2 and m.get_num_scores() == 6
    want = om.lpips_oracle(torch.from_numpy(img)[:, None].repeat(1, 3, 1, 1), torch.from_numpy(ref)[:, None].repeat(1, 3, 1, 1),
                           w, 'alex').double().numpy()
    assert np.allclose(np.array(m.get_all_scores()), want, rtol=1e-4, atol=0)
    # no weights available -> a clear error instead of invented numbers
    bare = create_metric('lpips')
    bare.batch_size = 1
    with pytest.raises(_lib.EvkError):
        bare.update(img[0], ref[0])
