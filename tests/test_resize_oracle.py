"""Pins oracle/resize_oracle.py (restatement of Pillow's 8-bit bilinear resample + ToTensor, the loader
transforms of coco_dataset.py:27,40-41) against committed Pillow outputs and, when importable, Pillow itself.
CPU only."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from oracle import resize_oracle as ro
import make_resize_golden as mg


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "resize_cases.npz"))


def test_oracle_matches_pillow_golden(golden):
    for i, (h, w) in enumerate(golden["shapes"]):
        a = mg.make_crop(1000 + i, int(h), int(w))
        assert np.array_equal(ro.resize_u8(a), golden["resized_%d" % i]), "crop %dx%d" % (h, w)


def test_to_tensor_matches_torchvision_golden(golden):
    h, w = golden["shapes"][0]
    a = mg.make_crop(1000, int(h), int(w))
    t = ro.resize_to_tensor(a)
    assert t.dtype == np.float32 and t.shape == (1, 32, 100)
    assert np.array_equal(t, golden["to_tensor_0"])


def test_identity_and_single_axis_cases():
    a = mg.make_crop(5, 32, 100)
    assert np.array_equal(ro.resize_u8(a), a)                      # both passes skipped
    b = mg.make_crop(6, 32, 37)
    assert ro.resize_u8(b).shape == (32, 100)
    const = np.full((11, 29), 173, np.uint8)
    assert (ro.resize_u8(const) == 173).all()                      # weights sum to 1 after rounding


@pytest.mark.skipif(importlib.util.find_spec("PIL") is None, reason="Pillow not installed")
def test_oracle_matches_installed_pillow():
    from PIL import Image
    rs = np.random.RandomState(3)
    for _ in range(25):
        h, w = int(rs.randint(1, 200)), int(rs.randint(1, 600))
        a = (rs.rand(h, w) * 256).astype(np.uint8)
        ref = np.asarray(Image.fromarray(a, "L").resize((100, 32), Image.BILINEAR))
        assert np.array_equal(ro.resize_u8(a), ref), "crop %dx%d" % (h, w)


def test_pack_crops_layout():
    from multimodal_scene_text_recognition_b200 import pipeline
    crops = [mg.make_crop(i, h, w) for i, (h, w) in enumerate([(3, 5), (1, 1), (7, 2)])]
    pixels, offsets, sizes, max_hw = pipeline.pack_crops(crops)
    assert offsets.tolist() == [0, 15, 16] and sizes.tolist() == [[3, 5], [1, 1], [7, 2]] and max_hw == (7, 5)
    assert pixels.numel() == 30 and np.array_equal(pixels.numpy()[16:], crops[2].reshape(-1))
    with pytest.raises(ValueError):
        pipeline.pack_crops([np.zeros((0, 4), np.uint8)])
    with pytest.raises(ValueError):
        pipeline.pack_crops([np.zeros((4, 4), np.float32)])
