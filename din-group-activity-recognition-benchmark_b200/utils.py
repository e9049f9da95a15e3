"""Host-side helpers the reference's `utils.py` exposes to the stage-2 path (utils.py:8-19, 101-105).

`prep_images` is kept for API compatibility only: on the hot path it is fused into the stem convolution
kernel (csrc/stem_pool.cu) and this function is never called by the models in this package.
"""
import torch


def prep_images(images):
    """(x / 255 - 0.5) * 2  (reference utils.py:8-19)."""
    return images.div(255.0).sub(0.5).mul(2.0)


def print_log(file_path, *args):
    print(*args)
    if file_path is not None:
        with open(file_path, "a") as f:
            print(*args, file=f)
