"""CPU: the host-only helpers of evreal_b200.eval_utils (tensor <-> image conversions, result / timestamp files) behave like
utils/eval_utils.py:38-77."""
import numpy as np
import torch

from evreal_b200 import eval_utils as eu


def test_cv2torch_and_torch2cv2_shapes_and_values():
    img = np.arange(12, dtype=np.float32).reshape(3, 4)
    t1 = eu.cv2torch(img)
    assert tuple(t1.shape) == (1, 1, 3, 4) and torch.equal(t1[0, 0], torch.from_numpy(img))
    t3 = eu.cv2torch(img, num_ch=3)                      # LPIPS input: the grey plane repeated
    assert tuple(t3.shape) == (1, 3, 3, 4) and all(torch.equal(t3[0, c], torch.from_numpy(img)) for c in range(3))
    chw = torch.arange(24, dtype=torch.float32).reshape(2, 3, 4)
    assert tuple(eu.cv2torch(chw).shape) == (1, 2, 3, 4)
    assert tuple(eu.cv2torch(chw[None]).shape) == (1, 2, 3, 4)
    back = eu.torch2cv2(t1)
    assert back.shape == (3, 4) and np.array_equal(back, img)
    hwc = eu.torch2cv2(chw[None])
    assert hwc.shape == (3, 4, 2) and np.array_equal(hwc[:, :, 1], chw[1].numpy())


def test_result_and_timestamp_files(tmp_path):
    p = str(tmp_path / 'mse.txt')
    eu.append_result(p, 3, 0.123456789)
    eu.append_result(p, [4, 5], [1.0, 2.5])
    eu.append_result(p, 6, 7, is_int=True)
    assert open(p).read() == '3 0.12346\n4 1.00000\n5 2.50000\n6 7\n'
    q = str(tmp_path / 'timestamps.txt')
    eu.append_timestamp(q, 0, 1.5)
    eu.append_timestamp(q, 1, 0.1)
    assert open(q).read() == '0 1.500000000000000\n1 0.100000000000000\n'
