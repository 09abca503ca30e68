"""Stub: plotting is out of scope; nothing here is ever called."""


def plot(*args, **kwargs):
    raise RuntimeError("matplotlib stub: plotting is not available")
