"""numpy restatement of the reference's composite drivers (TEST INFRASTRUCTURE ONLY):

  * ml_interpolate / ml_terms  -- sgmethods/multilevel_interpolant.py:89-150, with the
    non-existent method names of the shipped file (sampleOnSG, oldXx, oldSamples) read as
    sample_on_SG, old_xx, old_samples;
  * gg_indicators              -- sgmethods/compute_error_indicators.py:78-121, with
    mid_set.reducedMargin / mid_set.midSet read as reduced_margin / mid_set (B.2).

Parity status: these two reference drivers cannot run as shipped (AttributeError on first
use), so no reference output exists for them; the restatement is pinned only through its
building blocks (sg_oracle, which IS pinned to reference outputs).  "parity unpinned" for
the driver level itself.
"""
import numpy as np

from . import sg_oracle


def _restrict(plan_coarse, plan_fine, samples):
    vals, _ = sg_oracle.sample_on_grid(plan_coarse.SG, lambda y: 1 / 0, plan_fine.SG, samples)
    return vals


def ml_interpolate(plans, kind, yy, ml_samples):
    """I_0[s_K] then, for k = 1..K, += (I_k[s_{K-k}] - I_{k-1}[restricted s_{K-k}])."""
    n = len(plans)
    out = sg_oracle.interpolate(plans[0], yy, ml_samples[n - 1], kind, squeeze=False)
    for k in range(1, n):
        s = ml_samples[n - 1 - k]
        s_red = _restrict(plans[k - 1], plans[k], s)
        out = out + (sg_oracle.interpolate(plans[k], yy, s, kind, squeeze=False)
                     - sg_oracle.interpolate(plans[k - 1], yy, s_red, kind, squeeze=False))
    return out


def ml_terms(plans, kind, yy, ml_samples):
    """[(I_{K-k} - I_{K-k-1})[u_k]]_k, the last entry without subtraction."""
    n = len(plans)
    terms = []
    for k in range(n):
        lev = n - 1 - k
        vals = sg_oracle.interpolate(plans[lev], yy, ml_samples[k], kind, squeeze=False)
        if k < n - 1:
            s_red = _restrict(plans[lev - 1], plans[lev], ml_samples[k])
            vals = vals - sg_oracle.interpolate(plans[lev - 1], yy, s_red, kind, squeeze=False)
        terms.append(vals)
    return terms


def gg_indicators(mid_set_rows, reduced_margin, knots, lev2knots, f, SG, u_on_SG, yy_rnd,
                  norm, u_interp, kind="linear"):
    """One indicator per reduced-margin index: norm(I_{set + nu}[u](yy), u_interp)."""
    out = np.zeros(reduced_margin.shape[0])
    for i, nu in enumerate(reduced_margin):
        ext = np.vstack((mid_set_rows, nu))
        ext = ext[np.lexsort(np.rot90(ext))]
        plan = sg_oracle.make_plan(ext, knots, lev2knots)
        vals, _ = sg_oracle.sample_on_grid(plan.SG, f, SG, u_on_SG)
        out[i] = norm(sg_oracle.interpolate(plan, yy_rnd, vals, kind, squeeze=False), u_interp)
    return out
