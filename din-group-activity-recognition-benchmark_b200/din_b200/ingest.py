"""ingest.py — the loader's per-frame work on the device.

The reference's datasets decode and resize every frame on a DataLoader worker (volleyball.py:237-244,
collective.py:181-186): `Image.open(path)` -> `transforms.functional.resize(img, image_size)` -> `np.array(img)` ->
`transpose(2, 0, 1)` -> `.float()`.  Here the JPEG bytes go to the GPU: nvJPEG decodes them (library code, like cuBLAS for
a plain GEMM), `din_resize_bilinear_u8` resizes with Pillow's exact fixed-point arithmetic (bit-identical to
`Image.resize(..., BILINEAR)` on the same decoded pixels), and the result is the uint8 `[n, H, W, 3]` tensor the stem
kernels ingest directly (`model((frames_u8.view(B, T, H, W, 3), boxes))`): no float image ever exists.

Decoder note: nvJPEG and libjpeg-turbo (Pillow) are different JPEG decoders; on 4:4:4 streams they agree to +-1 LSB on a
fraction of a percent of the samples, on chroma-subsampled streams their upsampling filters differ at sharp colour
edges (tests/test_ingest_gpu.py states both tolerances).  The resize itself is exact.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def resize_u8(frames, size):
    """frames: uint8 [n, h, w, 3] (CUDA) -> uint8 [n, H, W, 3], == PIL's Image.resize((W, H), Image.BILINEAR) per frame."""
    if not (isinstance(frames, torch.Tensor) and frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 4
            and frames.shape[3] == 3 and frames.is_contiguous()):
        raise _lib.DinError("resize_u8: expected a contiguous CUDA uint8 tensor [n, h, w, 3]")
    n, h, w, _ = frames.shape
    oh, ow = size
    out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=frames.device)
    tmp = torch.empty((n, h, ow, 3), dtype=torch.uint8, device=frames.device) if (h != oh and w != ow) else None
    with torch.cuda.device(frames.device):
        check(_lib.load().din_resize_bilinear_u8(frames.data_ptr(), n, h, w, out.data_ptr(), oh, ow,
                                                 None if tmp is None else tmp.data_ptr(), _stream()),
              "din_resize_bilinear_u8")
    return out


def jpeg_size(data):
    """(height, width) of a JPEG byte string."""
    h, w = C.c_int(0), C.c_int(0)
    check(_lib.load().din_jpeg_image_info(data, len(data), C.byref(h), C.byref(w)), "din_jpeg_image_info")
    return h.value, w.value


def decode_resize(jpegs, size, device="cuda", cpu_threads=8, out=None):
    """jpegs: a sequence of JPEG byte strings (the files' contents) -> uint8 [n, H, W, 3] on `device`, decoded by nvJPEG and
    resized to size = (H, W) as the reference's loader does (volleyball.py:237-240)."""
    jpegs = [bytes(j) for j in jpegs]
    n = len(jpegs)
    if n == 0:
        raise _lib.DinError("decode_resize: no frames")
    oh, ow = size
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.DinError("decode_resize: the ingest path runs on a CUDA device (there is no CPU fallback)")
    with torch.cuda.device(device):
        need = 0
        for j in jpegs:
            h, w = jpeg_size(j)
            if (h, w) != (oh, ow):
                need += ((h * w * 3 + 255) & ~255) + ((h * ow * 3 + 255) & ~255)
        if out is None:
            out = torch.empty((n, oh, ow, 3), dtype=torch.uint8, device=device)
        elif not (out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (n, oh, ow, 3) and out.is_contiguous()):
            raise _lib.DinError("decode_resize: out must be a contiguous CUDA uint8 tensor [n, H, W, 3]")
        ws = torch.empty((need,), dtype=torch.uint8, device=device) if need else None
        ptrs = (C.c_char_p * n)(*jpegs)
        lens = (C.c_size_t * n)(*[len(j) for j in jpegs])
        check(_lib.load().din_jpeg_decode_resize_u8(ptrs, lens, n, out.data_ptr(), oh, ow,
                                                    None if ws is None else ws.data_ptr(), need, int(cpu_threads), _stream()),
              "din_jpeg_decode_resize_u8")
        # nvJPEG reads the host bitstreams asynchronously: they (and the workspace) must outlive the decode
        torch.cuda.current_stream().synchronize()
    return out
