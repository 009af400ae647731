"""CPU checks of oracle/multiscale.py (section 8(f) rows; parity unpinned, see its header): the resize restatement
against torch's own bilinear kernel, the luma formula on known values, the merge order, the level geometry."""
import numpy as np
import torch

from oracle import multiscale as oms


def test_resize_level_is_torch_bilinear():
    img = np.random.default_rng(0).integers(0, 256, (97, 131, 3)).astype(np.uint8)
    for hs, ws in ((68, 92), (48, 64), (97, 131), (32, 32)):
        mine = oms.resize_level(img, hs, ws)
        t = torch.from_numpy(img.astype(np.float32)).permute(2, 0, 1)[None]
        ref = torch.nn.functional.interpolate(t, size=(hs, ws), mode="bilinear", align_corners=False)[0].permute(1, 2, 0).numpy() / 255
        assert mine.shape == (hs, ws, 3) and mine.dtype == np.float32
        np.testing.assert_allclose(mine, ref, atol=4e-6)
    same = oms.resize_level(img, 97, 131)                       # identity scale: exactly the /255 image
    np.testing.assert_array_equal(same, img.astype(np.float32) / np.float32(255))


def test_luma_known_values():
    px = np.array([[[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128]]], dtype=np.uint8)
    np.testing.assert_array_equal(oms.rgb_to_gray(px)[0], [255, 0, 76, 150, 29, 128])      # Pillow's documented results


def test_level_sizes():
    assert [oms.level_size(1024, 0.7, l) for l in range(3)] == [1024, 717, 502]
    assert oms.level_size(40, 0.5, 3) == 32                                                # floor of 32 pixels


def test_merge_order_and_coordinates():
    a = (np.array([[10, 20], [30, 40]]), np.array([0.9, 0.5], dtype=np.float32))
    b = (np.array([[1, 2], [3, 4], [5, 6]]), np.array([0.9, 0.7, 0.5], dtype=np.float32))
    xy, sc, lv = oms.merge_levels([a, b], [(1.0, 1.0), (2.0, 2.0)], 4)
    np.testing.assert_array_equal(sc, np.array([0.9, 0.9, 0.7, 0.5], dtype=np.float32))
    np.testing.assert_array_equal(lv, [0, 1, 1, 0])                                        # ties: the finer level first
    np.testing.assert_array_equal(xy[1], [(1 + 0.5) * 2 - 0.5, (2 + 0.5) * 2 - 0.5])
    np.testing.assert_array_equal(xy[0], [10, 20])
