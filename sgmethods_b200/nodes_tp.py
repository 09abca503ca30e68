"""Tensor-product knots from a 1-D family (reference: sgmethods/nodes_tp.py:6-27)."""


def tp_knots(scalar_knots, num_knots_dir):
    """Tuple whose i-th entry is ``scalar_knots(num_knots_dir[i])``."""
    return tuple(scalar_knots(m) for m in num_knots_dir)
