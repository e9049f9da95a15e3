"""din_b200 — host-side plumbing for libdin_sm100.so (the sm_100a hot-path library).

PyTorch is used for device memory, streams and torch.distributed only; every arithmetic step of the
DIN stage-2 forward path runs in the hand-written CUDA kernels behind the C ABI declared in
``include/din_sm100.h``.  There is deliberately no CPU / eager fallback: if the shared library is
missing or a call fails, an exception is raised.
"""
from . import _lib  # noqa: F401
