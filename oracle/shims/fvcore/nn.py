"""Import shim: the reference imports fvcore.nn at module scope but only calls it in __main__ demos."""


def _missing(*a, **k):
    raise NotImplementedError("fvcore is not installed (import shim)")


activation_count = flop_count = parameter_count = parameter_count_table = _missing
