"""Stub: the golden-vector generator never plots."""


def subplots(*a, **k):  # pragma: no cover
    raise RuntimeError("matplotlib stub: plotting is not available")


def close(*a, **k):  # pragma: no cover
    pass
