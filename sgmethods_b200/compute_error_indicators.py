"""A-posteriori refinement indicators for adaptive sparse-grid interpolation
(Gerstner-Griebel indicators on the reduced margin).

Same function name and argument list as the reference's
sgmethods/compute_error_indicators.py (compute_GG_indicators_RM :16-121).  The reference is
broken as shipped (it reads ``mid_set.reducedMargin`` / ``mid_set.midSet``; MidSet defines
``reduced_margin`` / ``mid_set``); this module implements the
documented behaviour.  For every multi-index nu of the reduced margin that has no recycled
value, the indicator is  L2_err_param( I_{Lambda + nu}[u](yy_rnd), u_interp ).

B200 design: the extended interpolant of every candidate is built with the O(sum S_t) native
table builder (the reference rebuilds an O(M^2) interpolant per candidate), samples are
recycled through the hashed row matcher, the random points are uploaded once and every
candidate is evaluated by the fused kernel.  ``f`` and ``L2_err_param`` stay host callbacks.
"""
import numpy as np

from .sparse_grid_interpolant import SGInterpolant
from .tp_inteprolants import TPPwLinearInterpolator
from .utils import find_mid, lexic_sort


def compute_GG_indicators_RM(old_RM, old_estimators, mid_set, knots, lev2knots,
                             f, SG, u_on_SG, yy_rnd, L2_err_param, u_interp,
                             TP_inteporlant=TPPwLinearInterpolator, n_parallel=1):
    """Args as in the reference: old reduced margin and its indicators (recycled where the
    index is still in the reduced margin), the current MidSet, knots / lev2knots, the
    function ``f``, the current sparse grid ``SG`` with the values ``u_on_SG``, random
    parameters ``yy_rnd``, the norm callback ``L2_err_param(u_ext, u)`` and the current
    interpolant's values ``u_interp`` on ``yy_rnd``.  Returns one indicator per row of
    ``mid_set.reduced_margin``."""
    assert old_RM.shape[0] == old_estimators.size
    reduced_margin = mid_set.reduced_margin
    card_rm = reduced_margin.shape[0]
    indicators = np.zeros(card_rm)

    if old_RM.size == 0:
        to_compute = list(range(card_rm))
    else:
        dim_diff = reduced_margin.shape[1] - old_RM.shape[1]
        assert dim_diff >= 0
        if dim_diff > 0:
            old_RM = np.hstack((old_RM, np.zeros((old_RM.shape[0], dim_diff), dtype=int)))
        to_compute = []
        for i in range(card_rm):
            found, pos = find_mid(reduced_margin[i], old_RM)
            if found:
                indicators[i] = old_estimators[pos]
            else:
                to_compute.append(i)

    yy_dev = None
    if to_compute:                             # keep the random points resident on the GPU
        from ._device import torch_cuda
        from .sparse_grid_interpolant import _to_device
        torch = torch_cuda()
        yy_dev = _to_device(torch, yy_rnd, torch.device("cuda", torch.cuda.current_device()))
    for i in to_compute:
        nu = reduced_margin[i, :]
        extended = lexic_sort(np.vstack((mid_set.mid_set, nu)))
        interp_ext = SGInterpolant(extended, knots, lev2knots, tp_interpolant=TP_inteporlant,
                                   n_parallel=n_parallel, verbose=False)
        u_on_SG_ext = interp_ext.sample_on_SG(f, old_xx=SG, old_samples=u_on_SG)
        u_interp_ext = interp_ext.interpolate(yy_dev, u_on_SG_ext)
        indicators[i] = L2_err_param(u_interp_ext.cpu().numpy(), u_interp)
    return indicators
