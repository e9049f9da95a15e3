"""The resize oracle (oracle/pil_resize_oracle.py) pinned against Pillow itself: bit for bit (runs anywhere, no GPU)."""
import numpy as np
import pytest

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

import pil_resize_oracle as R  # noqa: E402

SHAPES = [((48, 64), (36, 54)),       # shrink both axes (antialiased: support > 1)
          ((30, 40), (48, 72)),       # enlarge both
          ((45, 80), (45, 64)),       # width only
          ((37, 53), (20, 53)),       # height only
          ((72, 128), (48, 72)),      # 720p -> Collective's 480 x 720 at 1/10 scale
          ((48, 64), (48, 64)),       # same size: a copy
          ((7, 5), (3, 11))]


@pytest.mark.parametrize("src,dst", SHAPES, ids=str)
def test_oracle_equals_pillow(src, dst):
    rng = np.random.default_rng(src[0] * 100 + dst[1])
    img = rng.integers(0, 256, size=src + (3,), dtype=np.uint8)
    img[: src[0] // 3] = 255                      # saturated and flat areas: the clip and the rounding constant
    img[-(src[0] // 4):, : src[1] // 2] = 0
    want = np.array(Image.fromarray(img).resize((dst[1], dst[0]), Image.BILINEAR))
    got = R.resize_bilinear_u8(img, dst)
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want), int(np.abs(got.astype(int) - want.astype(int)).max())


def test_torchvision_resize_is_that_pillow_call():
    """transforms.functional.resize on a PIL image (volleyball.py:239) == Image.resize(..., BILINEAR)."""
    tv = pytest.importorskip("torchvision.transforms.functional")
    rng = np.random.default_rng(3)
    img = Image.fromarray(rng.integers(0, 256, size=(40, 60, 3), dtype=np.uint8))
    a = np.array(tv.resize(img, (24, 36)))
    b = np.array(img.resize((36, 24), Image.BILINEAR))
    assert np.array_equal(a, b)
