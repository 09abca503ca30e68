"""Compatibility shim for code that uses the reference's per-term wrapper directly
(sgmethods/tp_interpolant_wrapper.py, TPInterpolatorWrapper :7-64).

``SGInterpolant.interpolate`` never goes through this class: selecting the active columns
and the constant (no active direction) case are folded into the fused kernels.  The shim
keeps the constructor / call protocol and evaluates through the same CUDA operators.
"""
import numpy as np


class TPInterpolatorWrapper:
    """One tensor-product term: an operator on the directions that have more than one node.

    Called with points of the full parameter space it selects the columns ``active_dims``
    and applies the operator; without active directions the term is the constant given by
    its single nodal value, repeated for every point."""

    def __init__(self, active_nodes_tuple, active_dims, f_on_nodes, tp_interpolant):
        self.active_dims = active_dims
        self.f_on_nodes = f_on_nodes
        self.l = tp_interpolant(active_nodes_tuple, f_on_nodes)

    def _constant(self, n_points):
        assert self.f_on_nodes.shape[0] == 1
        self.f_on_nodes = np.reshape(self.f_on_nodes, (1, -1))
        return np.tile(self.f_on_nodes, (n_points, 1))

    def __call__(self, x_new):
        pts = np.asarray(x_new)
        if pts.ndim == 1:
            pts = pts[:, None]
        cols = pts[:, self.active_dims]
        return self._constant(cols.shape[0]) if cols.shape[1] == 0 else self.l(cols)
