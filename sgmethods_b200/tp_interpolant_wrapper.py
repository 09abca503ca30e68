"""Wrapper around one tensor-product operator (reference: sgmethods/tp_interpolant_wrapper.py,
TPInterpolatorWrapper :7-64).  SGInterpolant.interpolate does not go through this class (its
column purge and constant path are folded into the fused kernels); it is kept so that code
using the wrapper directly keeps working, evaluating through the same CUDA kernels."""
import numpy as np


class TPInterpolatorWrapper:
    """Evaluates ``tp_interpolant(active_nodes_tuple, f_on_nodes)`` on the active columns of
    the input; a term without active directions is the constant ``f_on_nodes``."""

    def __init__(self, active_nodes_tuple, active_dims, f_on_nodes, tp_interpolant):
        self.f_on_nodes = f_on_nodes
        self.active_dims = active_dims
        self.l = tp_interpolant(active_nodes_tuple, self.f_on_nodes)

    def __call__(self, x_new):
        if len(x_new.shape) == 1:
            x_new = np.reshape(x_new, (-1, 1))
        x_new = x_new[:, self.active_dims]
        if x_new.shape[1] == 0:
            assert self.f_on_nodes.shape[0] == 1
            self.f_on_nodes = np.reshape(self.f_on_nodes, (1, -1))
            return np.repeat(self.f_on_nodes, x_new.shape[0], axis=0)
        return self.l(x_new)
