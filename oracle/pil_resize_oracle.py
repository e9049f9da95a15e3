"""pil_resize_oracle.py — TEST INFRASTRUCTURE (only tests/ may import this; never the product path).

A numpy restatement of Pillow's bilinear resize for 8-bit images, the algorithm behind the reference loader's
`transforms.functional.resize(img, self.image_size)` (volleyball.py:239, collective.py:183; torchvision's PIL path calls
`img.resize(size[::-1], Image.BILINEAR)`).  Pillow is a third-party dependency of the reference (requirements: Pillow,
unpinned; this container has 12.2.0); its algorithm, src/libImaging/Resample.c:

  precompute_coeffs():  scale = in / out; filterscale = max(scale, 1); support = 1.0 * filterscale (bilinear);
      per output xx: center = (xx + 0.5) * scale; xmin = int(center - support + 0.5) clipped to >= 0;
      xmax = int(center + support + 0.5) clipped to <= in; k[x] = triangle((x + xmin - center + 0.5) / filterscale),
      normalised to sum 1 (double precision).
  normalize_coeffs_8bpc():  kk = int(0.5 + k * 2^22)   (PRECISION_BITS = 32 - 8 - 2).
  ImagingResampleHorizontal_8bpc / Vertical_8bpc:  out = clip8((2^21 + sum pixel * kk) >> 22), horizontal pass first
      into a uint8 image, then the vertical pass.

Pinned: tests/test_ingest_cpu.py checks this restatement against Pillow itself, bit for bit, on up- and down-scaling
shapes (incl. Collective's 480 x 720 targets); the CUDA kernel (csrc/ingest.cu) is then held to the same bytes.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def axis_table(in_size, out_size):
    """-> (bounds [out, 2] int, coeff [out, ksize] int32) as precompute_coeffs + normalize_coeffs_8bpc build them."""
    scale = float(np.float32(in_size) - np.float32(0.0)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    coeff = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            k.append(1.0 - a if a < 1.0 else 0.0)
        ww = sum(k)
        for x in range(xmax):
            v = (k[x] / ww if ww != 0.0 else k[x]) * (1 << PRECISION_BITS)
            coeff[xx, x] = int(-0.5 + v) if v < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, coeff


def _pass(img, bounds, coeff, axis):
    """img uint8 [h, w, c]; resample along `axis` (0 = vertical, 1 = horizontal)."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for o in range(bounds.shape[0]):
        first, cnt = bounds[o]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for i in range(cnt):
            acc += src[first + i] * coeff[o, i]
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img, size):
    """img: uint8 [h, w, 3]; size = (oh, ow) -> uint8 [oh, ow, 3] == np.array(Image.fromarray(img).resize((ow, oh), BILINEAR))."""
    h, w = img.shape[:2]
    oh, ow = size
    if (h, w) == (oh, ow):
        return img.copy()
    cur = img
    if w != ow:
        cur = _pass(cur, *axis_table(w, ow), axis=1)
    if h != oh:
        cur = _pass(cur, *axis_table(h, oh), axis=0)
    return cur
