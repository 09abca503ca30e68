"""numpy restatement of the reference's sparse-grid interpolant (TEST INFRASTRUCTURE ONLY).

Follows, function by function (paths relative to /root/reference):
  * combination_coefficients  -- sgmethods/sparse_grid_interpolant.py:77-92
  * term_geometry             -- sgmethods/sparse_grid_interpolant.py:94-107, nodes_tp.py:6-27
  * build_sparse_grid         -- sgmethods/sparse_grid_interpolant.py:109-139
  * sample_on_grid            -- sgmethods/sparse_grid_interpolant.py:141-250
  * tp_linear                 -- sgmethods/tp_inteprolants.py:38-53 -> scipy
                                 RegularGridInterpolator(method='linear', bounds_error=False,
                                 fill_value=None): scipy/interpolate/_rgi.py:375-482 (__call__),
                                 :520-549 (_evaluate_linear), Cython find_indices
  * tp_linear_scipy           -- the literal third-party call of tp_inteprolants.py:49-53
  * tp_quadratic              -- sgmethods/tp_inteprolants.py:83-160
  * tp_cubic                  -- sgmethods/tp_inteprolants.py:189-277
  * tp_lagrange               -- sgmethods/tp_inteprolants.py:308-360
  * interpolate               -- sgmethods/sparse_grid_interpolant.py:252-279 +
                                 sgmethods/tp_interpolant_wrapper.py:41-64

The arithmetic is written so that every floating-point operation happens in the same order
as in the reference (weights are left-to-right products starting from 1, corner/node
contributions are added one after the other in C order, the term is scaled by its
combination coefficient and added to the running output), which makes the restatement
bit-identical to the reference on the golden fixtures (tests/test_oracle_golden.py).

Parity status: PINNED (see oracle/__init__.py).
"""
from itertools import product

import numpy as np

NODE_TOL = 1.e-10          # sparse_grid_interpolant.py:129 and :220


# --------------------------------------------------------------------------- tables
def combination_coefficients(mid_set):
    """c_nu = 1 + sum over later rows r of the lexsorted set with (r - nu) in {0,1}^N of
    (-1)^{|r-nu|}, searched only up to the first row whose leading entry is nu_0 + 2
    (sparse_grid_interpolant.py:77-92).  Returns one int64 per row (zeros included)."""
    mid_set = np.asarray(mid_set)
    card = mid_set.shape[0]
    first_rows = np.unique(mid_set[:, 0], return_index=True)[1]
    stop_of_lead = np.concatenate((first_rows[2:], [card, card]))
    coeffs = np.zeros(card, dtype=np.int64)
    for row in range(card):
        nu = mid_set[row]
        window = mid_set[row + 1:stop_of_lead[nu[0]]] - nu
        binary = window[np.all((window >= 0) & (window <= 1), axis=1)]
        coeffs[row] = 1 + np.sum(np.power(-1, np.sum(binary, axis=1)))
    return coeffs


class OraclePlan:
    """The list attributes the reference's SGInterpolant holds after its constructor
    (sparse_grid_interpolant.py:40-63)."""

    def __init__(self):
        self.N = 0
        self.coeffs = []        # combination_coeffs
        self.mids = []          # active_mids
        self.dims = []          # active_tp_dims
        self.nodes = []         # active_tp_nodes_list (tuples of 1-D knot arrays)
        self.shapes = []        # shapes of map_tp_to_SG
        self.maps = []          # map_tp_to_SG
        self.SG = None
        self.num_nodes = 0


def term_geometry(mid_set, knots, lev2knots):
    """Keep rows with c != 0 in set order; m = lev2knots(nu).astype(int); active dims are the
    ascending positions with m > 1; knots per active dim = knots(m_i); a term without
    active dims gets the pseudo shape (1,) and the tuple (knots(1),)
    (sparse_grid_interpolant.py:94-107)."""
    mid_set = np.asarray(mid_set)
    plan = OraclePlan()
    plan.N = mid_set.shape[1]
    coeffs = combination_coefficients(mid_set)
    for row in np.flatnonzero(coeffs):
        nu = mid_set[row]
        m = lev2knots(nu).astype(int)
        act = np.where(m > 1)[0]
        m_act = m[act]
        if m_act.shape[0] == 0:
            m_act = np.array([1])
        plan.coeffs.append(coeffs[row])
        plan.mids.append(nu)
        plan.dims.append(act)
        plan.nodes.append(tuple(knots(mi) for mi in m_act))
        plan.shapes.append(tuple(int(v) for v in m_act))
    return plan


def build_sparse_grid(plan):
    """Walk the terms in order and the nodes of each term in C order, embed each node in R^N
    with zeros in the inactive directions, look it up among ALL rows found so far with
    sum|diff| < 1e-10, number it by first appearance (sparse_grid_interpolant.py:109-139).
    Quadratic in the number of nodes, like the reference."""
    N = plan.N
    rows = np.zeros((0, N))
    plan.maps = []
    for dims, nodes, shape in zip(plan.dims, plan.nodes, plan.shapes):
        tmap = np.zeros(shape, dtype=np.int64)
        axes = [np.atleast_1d(np.asarray(g, dtype=float)) for g in nodes]
        for multi in product(*[range(s) for s in shape]):
            node = np.zeros(N)
            if dims.size:
                node[dims] = [axes[j][multi[j]] for j in range(len(multi))]
            hits = np.where(np.sum(np.abs(rows - node), axis=1) < NODE_TOL)[0]
            assert hits.shape[0] <= 1
            if hits.shape[0]:
                tmap[multi] = hits[0]
            else:
                rows = np.vstack((rows, node))
                tmap[multi] = rows.shape[0] - 1
        plan.maps.append(tmap)
    plan.SG = rows
    plan.num_nodes = rows.shape[0]
    return plan


def make_plan(mid_set, knots, lev2knots):
    return build_sparse_grid(term_geometry(mid_set, knots, lev2knots))


def plan_from_tables(N, coeffs, dims, nodes, maps, SG=None):
    """Assemble an OraclePlan from tables computed elsewhere (used by the CPU-baseline leg
    of bench.py at sizes where the reference constructor, O(M^2), is infeasible, after the
    tables were proven identical on small sets)."""
    plan = OraclePlan()
    plan.N = N
    plan.coeffs = list(coeffs)
    plan.dims = [np.asarray(d, dtype=np.int64) for d in dims]
    plan.nodes = list(nodes)
    plan.maps = [np.asarray(m) for m in maps]
    plan.shapes = [m.shape for m in plan.maps]
    plan.SG = SG
    plan.num_nodes = 0 if SG is None else SG.shape[0]
    return plan


def sample_on_grid(SG, f, old_xx=None, old_samples=None, N=None):
    """Values of f on the rows of SG, reusing old samples whose (zero-padded or truncated)
    coordinates match a row within 1e-10 in the 1-norm (sparse_grid_interpolant.py:141-250,
    serial branch).  Returns (values, n_recycled)."""
    N = SG.shape[1] if N is None else N
    if old_samples is not None:
        old_samples = np.atleast_1d(old_samples)
        dim_f = old_samples.shape[1]
    else:
        dim_f = np.atleast_1d(f(SG[0])).size
    vals = np.zeros((SG.shape[0], dim_f))
    if old_xx is None or old_samples is None:
        old_xx = np.zeros((0, N))
        old_samples = np.zeros((0, dim_f))
    assert old_xx.shape[0] == old_samples.shape[0]
    assert old_samples.shape[1] == dim_f
    if old_xx.shape[1] < N:
        old_xx = np.hstack((old_xx, np.zeros((old_xx.shape[0], N - old_xx.shape[1]))))
    elif old_xx.shape[1] > N:
        # NB the reference truncates the tail norm to int before comparing with 0 (:208-209)
        tail = np.linalg.norm(old_xx[:, N:], ord=1, axis=1).astype(int)
        keep = np.where(tail == 0)[0]
        old_xx = old_xx[keep, :N]
        old_samples = old_samples[keep]
    recycled = 0
    for n in range(SG.shape[0]):
        hits = np.where(np.linalg.norm(old_xx - SG[n], 1, axis=1) < NODE_TOL)[0]
        assert hits.size <= 1
        if hits.size:
            vals[n] = old_samples[hits[0]]
            recycled += 1
        else:
            vals[n] = f(SG[n])
    return vals, recycled


# --------------------------------------------------------------------------- operators
def _as_axes(nodes):
    return [np.atleast_1d(np.asarray(g, dtype=float)) for g in nodes]


def _brackets_linear(g, x):
    """scipy find_indices(extrapolate): i with g[i] <= x < g[i+1], 0 below the grid, n-2 at
    or above the last knot; y = (x - g[i]) / (g[i+1] - g[i]); NaN -> y = NaN."""
    n = g.shape[0]
    i = np.clip(np.searchsorted(g, x, side="right") - 1, 0, n - 2)
    i = np.where(np.isnan(x), 0, i)
    y = (x - g[i]) / (g[i + 1] - g[i])
    return i, y


def tp_linear(nodes, values, x):
    """Multilinear interpolation/extrapolation on a rectilinear grid, arithmetic order of
    scipy's _evaluate_linear: for each hypercube vertex in itertools.product order (first
    dimension slowest, lower knot first) weight = ((1*w_0)*w_1)..., value = value +
    values[vertex] * weight, starting from 0; rows with a NaN coordinate give NaN."""
    axes = _as_axes(nodes)
    k = len(axes)
    P = x.shape[0]
    lo, w_hi = [], []
    for j in range(k):
        i, y = _brackets_linear(axes[j], x[:, j])
        lo.append(i)
        w_hi.append(y)
    w_lo = [1 - y for y in w_hi]
    vals = values.reshape(tuple(a.shape[0] for a in axes) + (-1,))
    acc = np.zeros((P, vals.shape[-1]))
    for corner in product((0, 1), repeat=k):
        w = np.ones(P)
        for j, bit in enumerate(corner):
            w = w * (w_hi[j] if bit else w_lo[j])
        idx = tuple(lo[j] + corner[j] for j in range(k))
        acc = acc + vals[idx] * w[:, None]
    acc[np.any(np.isnan(x[:, :k]), axis=1)] = np.nan
    return acc


def tp_linear_scipy(nodes, values, x):
    """The literal call the reference makes (tp_inteprolants.py:49-53)."""
    from scipy.interpolate import RegularGridInterpolator
    return RegularGridInterpolator(tuple(nodes), values, method="linear",
                                   bounds_error=False, fill_value=None)(x)


def _stencil(g, x, width, strict_right):
    """Index of the polynomial piece of x for pieces made of `width`+1 consecutive knots that
    share their end knots (pieces start at knots 0, width, 2*width, ...).  A point on a
    shared knot belongs to the piece on its LEFT; points left of the grid use piece 0.
    Quadratic (strict_right): a point right of the last knot (or NaN) is an IndexError in
    the reference (tp_inteprolants.py:120-134).  Cubic: it uses the last piece (:224-232)."""
    h = g[0::width]
    j = np.searchsorted(h, x, side="left") - 1          # #(h < x) - 1
    j = np.where(np.isnan(x), h.shape[0] - 1, j)
    j = np.maximum(j, 0)
    if strict_right:
        if h.shape[0] < 2 or np.any(j > h.shape[0] - 2):
            raise IndexError("pw-quadratic: point right of the last knot (reference "
                             "tp_inteprolants.py:131-133 indexes past the knot vector)")
    else:
        if h.shape[0] < 2:
            raise UnboundLocalError("pw-cubic needs at least 4 knots per active direction "
                                    "(reference tp_inteprolants.py:232)")
        j = np.minimum(j, h.shape[0] - 2)
    return j


def _local_lagrange_weights(z, xs):
    """Weights of the local Lagrange basis on the knots xs = [x0..xq] evaluated as in the
    reference: numerator and denominator are left-to-right products over the other knots
    in ascending order, then one division (tp_inteprolants.py:134-136, :246-253)."""
    out = []
    for a in range(len(xs)):
        num, den = None, None
        for b in range(len(xs)):
            if b == a:
                continue
            num = (z - xs[b]) if num is None else num * (z - xs[b])
            den = (xs[a] - xs[b]) if den is None else den * (xs[a] - xs[b])
        out.append(num / den)
    return out


def _tp_local(nodes, values, x, q):
    """Shared body of the pw-quadratic (q=2) and pw-cubic (q=3) operators."""
    axes = _as_axes(nodes)
    k = len(axes)
    P = x.shape[0]
    assert x.shape[1] == k
    start, wts = [], []
    for j in range(k):
        g = axes[j]
        s = _stencil(g, x[:, j], q, strict_right=(q == 2))
        if q == 3:
            g = g[:3 * ((g.shape[0] - 1) // 3) + 1]
        xs = [g[q * s + a] for a in range(q + 1)]
        start.append(q * s)
        wts.append(_local_lagrange_weights(x[:, j], xs))
    vals = values.reshape(tuple(a.shape[0] for a in axes) + (-1,))
    acc = np.zeros((P, vals.shape[-1]))
    for corner in product(range(q + 1), repeat=k):
        w = np.ones(P)
        for j, a in enumerate(corner):
            w = w * wts[j][a]
        idx = tuple(start[j] + corner[j] for j in range(k))
        acc = acc + vals[idx] * w[:, None]
    return acc


def tp_quadratic(nodes, values, x):
    return _tp_local(nodes, values, x, 2)


def tp_cubic(nodes, values, x):
    return _tp_local(nodes, values, x, 3)


def lagrange_lambdas(g):
    """lambda_i = 1 / prod_{j != i, ascending} (g_i - g_j) (tp_inteprolants.py:311-320)."""
    n = g.shape[0]
    lam = np.ones(n)
    for i in range(n):
        pr = 1.0
        for j in range(n):
            if j != i:
                pr = pr * (g[i] - g[j])
        lam[i] = pr
    return 1 / lam


def tp_lagrange(nodes, values, x):
    """Global tensor Lagrange interpolation in first barycentric form:
    l(x) = prod_j (x - g_j) (ascending), b_i = (l * lambda_i) / (x - g_i) -- NaN when x
    equals a knot --, tensor basis = ((1*b^0)*b^1)..., value = sum over ALL nodes in C order
    (tp_inteprolants.py:322-360)."""
    axes = _as_axes(nodes)
    k = len(axes)
    P = x.shape[0]
    assert x.shape[1] == k
    basis = []
    for j in range(k):
        g = axes[j]
        lam = lagrange_lambdas(g)
        diff = x[:, j][None, :] - g[:, None]                    # (n, P)
        ell = diff[0].copy()
        for r in range(1, g.shape[0]):
            ell = ell * diff[r]
        with np.errstate(divide="ignore", invalid="ignore"):
            basis.append((ell[None, :] * lam[:, None]) / diff)   # (n, P)
    shape = tuple(a.shape[0] for a in axes)
    vals = values.reshape(shape + (-1,))
    acc = np.zeros((P, vals.shape[-1]))
    for multi in product(*[range(s) for s in shape]):
        w = np.ones(P)
        for j, i in enumerate(multi):
            w = w * basis[j][i]
        acc = acc + vals[multi][None, :] * w[:, None]
    return acc


OPERATORS = {"linear": tp_linear, "linear_scipy": tp_linear_scipy,
             "quadratic": tp_quadratic, "cubic": tp_cubic, "lagrange": tp_lagrange}


def interpolate(plan, x_new, f_on_SG, kind="linear", squeeze=True):
    """out = sum_t c_t * TP_t(x[:, active dims of t]); a term without active dims
    contributes c_t * f_on_SG[map_t[0]] to every point; np.squeeze on return
    (sparse_grid_interpolant.py:268-279, tp_interpolant_wrapper.py:54-64)."""
    op = OPERATORS[kind]
    x_new = np.asarray(x_new, dtype=float)
    if x_new.ndim == 1:
        x_new = x_new.reshape(-1, 1)
    f_on_SG = np.asarray(f_on_SG, dtype=float)
    out = np.zeros((x_new.shape[0], f_on_SG.shape[1]))
    for c, dims, nodes, tmap in zip(plan.coeffs, plan.dims, plan.nodes, plan.maps):
        f_grid = f_on_SG[tmap, :]
        if dims.size == 0:
            tp = np.repeat(f_grid.reshape(1, -1), x_new.shape[0], axis=0)
        else:
            tp = op(nodes, f_grid, x_new[:, dims])
        out = out + c * tp
    return np.squeeze(out) if squeeze else out
