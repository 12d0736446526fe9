"""GPU parity of the configuration bench.py actually times (round-1 verdict, weak #2): SequenceBatch at B = 36, full size,
CUDA-graph replay, overlapped side streams -- rows {0, 17, 35} of every step against the CPU oracle of the per-frame loop
and against the same stream run alone (B = 1, no overlap).  Same for FireNet at B = 36 and HyperE2VID at B = 8."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STEPS = 3


def _streams(B, H, W, rate, fps, rows):
    from evreal_b200 import synthetic
    from evreal_b200.dataset import MemMapDataset
    dur = (STEPS + 2.2) / fps
    arrs = [synthetic.make_stream(H, W, rate, dur, fps, seed=100 + b) if b in rows else None for b in range(B)]
    filler = synthetic.make_stream(H, W, rate, dur, fps, seed=99)
    arrs = [dict(a) if a is not None else dict(filler) for a in arrs]
    for a in arrs:
        # frame 0 moved 15 s into the past: the reference's "skip items more than 10 s before the evaluation window" rule
        # (eval.py:212-213) then skips exactly item 0 (the always-empty window), and the oracle starts at item 1 from zero
        # state like the batch below
        ts = np.asarray(a['images_ts'], dtype=np.float64).reshape(-1).copy()
        ts[0] = ts[1] - 15.0
        a['images_ts'] = ts.reshape(-1, 1)
    dss = [MemMapDataset(a, num_bins=5, voxel_method={'method': 'between_frames'}, resident=False) for a in arrs]
    return arrs, dss


def _run_batch(model, dss, norm, post, overlap, resident=True):
    from evreal_b200.pipeline import SequenceBatch
    batch = SequenceBatch(model, dss, norm, post, resident=resident, overlap=overlap, log_scores=True)
    batch.reset()
    images = []
    for k in range(1, STEPS + 1):
        _, img, _ = batch.step(k, sync=False)
        if not resident:
            batch.result_event.synchronize()
        else:
            batch.wait_results()
        images.append(img.clone().cpu().numpy() if resident else img.numpy().copy())
    batch.finish()
    batch.check_bounds()
    return np.stack(images), batch.scores_log.cpu().numpy()[1:STEPS + 1]


def _check_rows(name, make_model, make_oracle, H, W, rate, fps, B, norm, post, num_encoders, rows):
    from oracle import eval_loop
    arrs, dss = _streams(B, H, W, rate, fps, rows)
    images, scores = _run_batch(make_model(), dss, norm, post, overlap=True)
    assert images.shape[:2] == (STEPS, B)
    for b in rows:
        start = float(np.asarray(arrs[b]['images_ts']).reshape(-1)[1])
        ref1 = eval_loop.run_sequence(arrs[b], (H, W), make_oracle(), num_encoders, norm, post, start_time_s=start, max_items=STEPS + 1)
        assert ref1['indices'] == list(range(1, STEPS + 1)) and ref1['frames'] == STEPS
        for k in range(STEPS):
            want = ref1['images'][k]
            got = np.clip(images[k, b, 0], 0.0, 1.0)
            err = np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-6)
            assert err <= 1e-4, (name, 'row', b, 'step', k, err)
            assert abs(scores[k, b, 0] - ref1['mse'][k]) <= 1e-4 * ref1['mse'][k], (name, b, k)
            assert abs(scores[k, b, 1] - ref1['ssim'][k]) <= 1e-4 * abs(ref1['ssim'][k]) + 1e-7, (name, b, k)
        # the same stream alone, eager order of stages (no side streams): the batch row must be the same frames
        single_img, single_sc = _run_batch(make_model(), [dss[b]], norm, post, overlap=False)
        # (voxelizer float atomics commute but do not associate: a few ulp from run to run)
        assert np.max(np.abs(single_img[:, 0] - images[:, b])) <= 2e-5 * max(np.max(np.abs(images[:, b])), 1e-6), (name, b)
        assert np.allclose(single_sc[:, 0, :2], scores[:, b, :2], rtol=2e-5, atol=1e-9), (name, b)


def test_e2vid_batch36_graph_overlap_rows_vs_oracle():
    import evreal_b200 as evk
    from evreal_b200 import synthetic
    from oracle import networks as on
    sd = synthetic.unet_state_dict(seed=0, norm_bn=True)
    w = {k[len('unetrecurrent.'):]: v for k, v in sd.items()}
    _check_rows('e2vid', lambda: evk.E2VIDRecurrent(dict(synthetic.E2VID_KWARGS)).load_state_dict(sd).to('cuda'),
                lambda: on.UNetRecurrentOracle(w, final_sigmoid=True), 180, 240, 1.0e6, 24.0, 36, True, 'robust', 3, (0, 17, 35))


def test_firenet_batch36_graph_overlap_rows_vs_oracle():
    import evreal_b200 as evk
    from evreal_b200 import synthetic
    from oracle import networks as on
    sd = synthetic.firenet_state_dict(seed=0)
    w = {k[len('net.'):]: v for k, v in sd.items()}
    _check_rows('firenet', lambda: evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(sd).to('cuda'),
                lambda: on.FireNetLegacyOracle(w), 180, 240, 1.0e6, 25.0, 36, True, 'none', 4, (0, 17, 35))


def test_hyper_e2vid_batch8_graph_overlap_rows_vs_oracle():
    import evreal_b200 as evk
    from evreal_b200 import synthetic
    from oracle import networks as on
    sd = synthetic.unet_state_dict(seed=0, dynamic_decoder=True)
    w = {k[len('unetrecurrent.'):]: v for k, v in sd.items()}
    _check_rows('hyper', lambda: evk.E2VIDRecurrent(dict(synthetic.HYPER_KWARGS)).load_state_dict(sd).to('cuda'),
                lambda: on.UNetRecurrentOracle(w, dynamic_decoder=True), 260, 346, 5.0e6, 45.0, 8, False, 'none', 3, (0, 7))


def test_host_mode_overlap_matches_resident():
    """events in pinned host memory, copies on their own streams, all stages overlapped: same frames as the resident run"""
    import evreal_b200 as evk
    from evreal_b200 import synthetic
    sd = synthetic.firenet_state_dict(seed=0)
    _, dss = _streams(6, 180, 240, 1.0e6, 25.0, range(6))
    make = lambda: evk.FireNet_legacy(dict(synthetic.FIRENET_KWARGS)).load_state_dict(sd).to('cuda')
    img_r, sc_r = _run_batch(make(), dss, True, 'none', overlap=True, resident=True)
    img_h, sc_h = _run_batch(make(), dss, True, 'none', overlap=True, resident=False)
    assert np.max(np.abs(img_r - img_h)) <= 2e-5 * np.max(np.abs(img_r))
    assert np.allclose(sc_r[:, :, :2], sc_h[:, :, :2], rtol=2e-5, atol=1e-9)
