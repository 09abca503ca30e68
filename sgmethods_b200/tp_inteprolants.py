"""Tensor-product interpolation operators = the plugin boundary of the sparse-grid interpolant.

Same class names and protocol as the reference's sgmethods/tp_inteprolants.py
(``cls(nodes_tuple, f_on_nodes)`` then ``obj(x_new) -> (P, dim_f)``, :1-11):

  TPPwLinearInterpolator     (:21-53,   scipy RegularGridInterpolator, linear extrapolation)
  TPPwQuadraticInterpolator  (:55-160)
  TPPwCubicInterpolator      (:162-277)
  TPLagrangeInterpolator     (:279-360, first barycentric form)

Here a class is (a) the tag that selects the fused CUDA kernel when passed to
``SGInterpolant(..., tp_interpolant=cls)`` and (b) a working operator on its own: calling an
instance evaluates a one-term plan on the GPU with the same kernels (no CPU fallback).
User-defined operator classes are not supported by the fused path: SGInterpolant raises
NotImplementedError for them.
"""
import numpy as np

from . import _lib


class _TPOperator:
    KIND = None

    def __init__(self, nodes_tuple, f_on_nodes):
        self.nodes_tuple = nodes_tuple
        self.f_on_nodes = f_on_nodes
        # kept for attribute compatibility with the reference classes
        self.f = f_on_nodes
        self.n_dims = len(nodes_tuple)
        self._plan = None

    def _tables(self):
        axes = [np.atleast_1d(np.asarray(g, dtype=np.float64)) for g in self.nodes_tuple]
        shape = tuple(a.shape[0] for a in axes)
        f = np.asarray(self.f_on_nodes, dtype=np.float64)
        if f.ndim == len(shape):                     # scalar field
            f = f.reshape(shape + (1,))
        assert shape == f.shape[:-1]
        return axes, shape, np.ascontiguousarray(f.reshape(-1, f.shape[-1]))

    def __call__(self, x_new):
        from ._device import DevicePlan, torch_cuda
        torch = torch_cuda()
        axes, shape, f = self._tables()
        x_new = np.asarray(x_new, dtype=np.float64)
        assert x_new.shape[1] == len(shape)
        if self.KIND == _lib.KIND_CUBIC and any(s > 1 and s < 4 for s in shape):
            raise UnboundLocalError("pw-cubic needs at least 4 knots per direction")
        if self._plan is None:
            # one knot family per direction with more than one node
            term_fam, fam_off, knots = [], [0], []
            for a in axes:
                if a.shape[0] > 1:
                    term_fam.append(len(knots))
                    knots.append(a)
                    fam_off.append(fam_off[-1] + a.shape[0])
                else:
                    term_fam.append(-1)
            S = int(np.prod(shape))
            self._plan = DevicePlan(self.KIND, len(shape), S, [1.0], [term_fam], [0, S],
                                    np.arange(S, dtype=np.int32), fam_off,
                                    np.concatenate(knots) if knots else np.zeros(0))
        dev = torch.device("cuda", self._plan.device)
        out = self._plan.eval(torch.from_numpy(np.ascontiguousarray(x_new)).to(dev),
                              torch.from_numpy(f).to(dev))
        if self.KIND == _lib.KIND_QUADRATIC and self._plan.take_flags() & _lib.FLAG_QUADRATIC_RIGHT:
            raise IndexError("pw-quadratic: point right of the last knot (or NaN)")
        return out.cpu().numpy()


class TPPwLinearInterpolator(_TPOperator):
    """Tensor-product piecewise-linear interpolation (linear extrapolation outside)."""
    KIND = _lib.KIND_LINEAR


class TPPwQuadraticInterpolator(_TPOperator):
    """Tensor-product piecewise-quadratic interpolation (odd number of knots per direction)."""
    KIND = _lib.KIND_QUADRATIC


class TPPwCubicInterpolator(_TPOperator):
    """Tensor-product piecewise-cubic interpolation (at least 4 knots per direction)."""
    KIND = _lib.KIND_CUBIC


class TPLagrangeInterpolator(_TPOperator):
    """Tensor-product global Lagrange interpolation, first barycentric form."""
    KIND = _lib.KIND_LAGRANGE


def resolve_kind(tp_interpolant):
    """Kernel kind of an operator class or factory.  Factories (the reference's tests pass
    ``lambda nodes, f: TPPwLinearInterpolator(nodes, f)``) are probed once with a 7-knot
    dummy grid, which every shipped operator accepts."""
    if isinstance(tp_interpolant, type) and issubclass(tp_interpolant, _TPOperator):
        return tp_interpolant.KIND
    if callable(tp_interpolant):
        try:
            probe = tp_interpolant((np.linspace(-1., 1., 7),), np.zeros((7, 1)))
        except Exception as exc:
            raise NotImplementedError("tp_interpolant could not be probed; only the operator "
                                      "classes of sgmethods_b200.tp_inteprolants (or factories "
                                      "returning them) run on the fused CUDA path") from exc
        if isinstance(probe, _TPOperator):
            return probe.KIND
    raise NotImplementedError("user-defined tensor-product operators are not supported by the "
                              "fused CUDA path (no CPU fallback)")
