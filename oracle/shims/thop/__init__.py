"""Import shim: the reference imports thop at module scope but only calls it in __main__ demos."""


def profile(*a, **k):
    raise NotImplementedError("thop is not installed (import shim)")


def clever_format(*a, **k):
    raise NotImplementedError("thop is not installed (import shim)")
