"""Host side of the opt-in hierarchical-surplus form (sgmethods_b200/hierarchical.py): the
hierarchisation batches and evaluation tables are checked on the CPU by a plain numpy
emulation of what the kernels do, against the oracle's combination-formula result."""
import numpy as np
import pytest

from oracle import sg_oracle
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.hierarchical import HierarchicalPlan
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant


def basis_table(hp, fam, xd):
    """numpy twin of fill_basis<SGM_HIER>: value of the family's one non-zero basis function at
    xd and the rank of its knot among the family's new knots."""
    from sgmethods_b200 import _lib
    g, nr, hb = hp.fam_knots_list[fam], hp.fam_newrank_list[fam], hp.fam_basis[fam]
    if hb == _lib.HB_HAT:
        i = np.clip(np.searchsorted(g, xd, side="right") - 1, 0, g.size - 2)
        y = (xd - g[i]) / (g[i + 1] - g[i])
        lo_new = nr[i] >= 0
        return np.where(lo_new, 1 - y, y), np.where(lo_new, nr[i], nr[i + 1])
    if hb == _lib.HB_QUAD_MID:
        H = (g.size + 1) // 2
        j = np.maximum(np.searchsorted(g[0::2], xd, side="left") - 1, 0)
        assert (j <= H - 2).all()                        # right of the last knot: IndexError path
        x0, x1, x2 = g[2 * j], g[2 * j + 1], g[2 * j + 2]
        return ((xd - x0) * (xd - x2)) / ((x1 - x0) * (x1 - x2)), nr[2 * j + 1]
    p = int(np.flatnonzero(nr >= 0)[0])
    num, den, on_knot = np.ones_like(xd), 1.0, np.zeros(xd.shape, dtype=bool)
    for i in range(g.size):
        on_knot |= xd == g[i]
        if i != p:
            num, den = num * (xd - g[i]), den * (g[p] - g[i])
    v = num / den
    if hb == _lib.HB_LAG_ONE:
        v = np.where(on_knot, np.nan, v)
    return v, np.zeros(xd.shape, dtype=np.int64)


def surpluses(hp, f):
    """numpy twin of sgm_hier_apply_batches."""
    alpha = np.array(f[hp.sel], dtype=float)
    for ch, pa, pb, wa, wb in hp.batches:
        coarse = wa[:, None] * alpha[pa]
        two = pb >= 0
        coarse[two] = coarse[two] + wb[two, None] * alpha[pb[two]]
        alpha[ch] -= coarse
    return alpha


def emulate(hp, x, f):
    """numpy twin of sgm_hier_apply + the SGM_HIER evaluation."""
    alpha = surpluses(hp, f)
    it = hp.it
    out = np.zeros((x.shape[0], f.shape[1]))
    for t in range(hp._term_fam.shape[0]):
        w = np.ones(x.shape[0])
        off = np.zeros(x.shape[0], dtype=np.int64)
        act = np.flatnonzero(hp._term_fam[t] >= 0)
        sizes = [int((hp.fam_newrank_list[hp._term_fam[t, d]] >= 0).sum()) for d in act]
        for a, d in enumerate(act):
            phi, j = basis_table(hp, hp._term_fam[t, d], x[:, d])
            w = w * phi
            off = off + j * int(np.prod(sizes[:a], dtype=np.int64))      # first direction fastest
        out += alpha[hp._maps_f[hp._map_off[t] + off]] * w[:, None]
    return out


def emulate_groups(hp, x, f):
    """numpy twin of eval_hier_groups_kernel (kernel H2) on the tables of
    sgm_host_hier_groups: same traversal, same operations."""
    from sgmethods_b200 import _lib
    alpha = surpluses(hp, f)
    tf = hp._term_fam
    T, N = tf.shape
    extent = np.array([(nr >= 0).sum() for nr in hp.fam_newrank_list], dtype=np.int32)
    groups, parents, steps, hmap, root, split = _lib.hier_groups(tf, extent, hp._map_off[:T], 64)
    assert split[0] == 0 and split[-1] == groups.shape[0] and np.all(np.diff(split) >= 0)
    # summation order: every group names its list and its position in it; positions count up
    # along the table (= execution) order, so a list's earlier groups are always taken earlier
    lists, pos = (groups[:, 1] >> 24) & 0x3f, groups[:, 11]
    for w in range(split.size - 1):
        assert np.array_equal(pos[lists == w], np.arange(split[w + 1] - split[w]))
    # the kernel's own value table: a gather of the surpluses through hmap, zeros in the padding
    used = hmap[hmap >= 0]
    assert np.array_equal(np.sort(used), np.arange(hp._map_off[T]))      # every surplus exactly once
    A = np.where((hmap >= 0)[:, None], alpha[hp._maps_f[np.maximum(hmap, 0)]], 0.0)
    # entries in order of first use (the plan's numbering) with their per-point tables
    ent = {}
    for t in range(T):
        for d in range(N):
            if tf[t, d] >= 0 and (d, tf[t, d]) not in ent:
                ent[(d, tf[t, d])] = len(ent)
    phi = np.zeros((len(ent), x.shape[0]))
    rank = np.zeros((len(ent), x.shape[0]), dtype=np.int64)
    for (d, fam), e in ent.items():
        phi[e], rank[e] = basis_table(hp, fam, x[:, d])
    out = np.zeros((x.shape[0], f.shape[1])) + (A[root] if root >= 0 else 0.0)
    leaves = 0
    for gr, pr in zip(groups, parents):
        n_steps, n_mem, n_par = gr[1] & 0xffff, (gr[1] >> 16) & 0xf, (gr[1] >> 20) & 0xf
        mu = gr.view(np.uint16)
        w_par, off_par, stride = np.ones(x.shape[0]), np.zeros(x.shape[0], dtype=np.int64), 1
        for j in range(n_par):
            e, ext = (int(pr[j]) & 0xffff) // 64, (int(pr[j]) & 0xffffffff) >> 16
            w_par = w_par * phi[e]
            off_par = off_par + rank[e] * stride
            stride *= ext
        assert stride == gr[2]
        acc = [np.zeros_like(out) for _ in range(n_mem)]
        w, off = [], []
        for m in range(7):
            e = int(mu[8 + m]) // 64             # the dummy entry (hat 1, rank 0) where unused
            assert int(mu[8 + m]) % 64 == 0 and (e < len(ent) or (e == len(ent) and (m >= n_mem or n_par == 0)))
            if m < n_mem:
                w.append(w_par * (phi[e] if e < len(ent) else 1.0))
                off.append(off_par + gr[2] * (rank[e] if e < len(ent) else 0))
        bounds = [int(b) for b in gr.view(np.uint16)[16:22]] + [n_steps]
        for si, (sbase, word) in enumerate(steps[gr[0]:gr[0] + n_steps]):
            e, slot = (int(word) & 0xffff) // 64, (int(word) & 0xffffffff) >> 16
            nv = 7 - next(i for i in range(7) if si < bounds[i])               # member count class
            assert 1 <= nv <= n_mem
            for m in range(nv):
                idx = sbase + m * slot + off[m] + gr[3] * rank[e]
                assert (idx // 16 == (sbase + m * slot) // 16).all() or slot > 16   # one 128-byte line
                acc[m] += A[idx] * phi[e][:, None]
                leaves += 1
        for m in range(n_mem):
            out += w[m][:, None] * acc[m]
    assert leaves == T - (1 if root >= 0 else 0)            # every non-root multi-index exactly once
    return out


CASES = ["cc", "gauss", "aniso", "quad_cc", "quad_aniso", "lag_leja", "lag_cc"]


@pytest.mark.parametrize("case", CASES)
def test_group_tables_of_the_grouped_kernel_reproduce_the_combination_formula(case):
    it, hp, x, f, ref = _case(case)
    got = emulate_groups(hp, x, f)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("case", CASES)
def test_hierarchical_tables_reproduce_the_combination_formula(case):
    it, hp, x, f, ref = _case(case)
    got = emulate(hp, x, f)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


def _case(case):
    rng = np.random.default_rng(3)
    op = None
    if case == "cc":
        S, knots = mis.smolyak_mid_set(4, 3), nodes_1d.cc_nodes
        lev = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
        x = rng.uniform(-1.3, 1.3, (200, 3))             # includes extrapolation
    elif case == "gauss":
        S, knots = mis.smolyak_mid_set(3, 5), nodes_1d.unbounded_nodes_nested
        lev = lambda n: 2 ** (n + 1) - 1  # noqa: E731
        x = rng.normal(0, 2.0, (200, 5))                 # beyond the outermost knots too
    elif case == "aniso":
        S, knots = mis.aniso_smolyak_mid_set(4, 4, [1, 1.5, 2, 4]), nodes_1d.unbounded_nodes_nested
        lev = lambda n: 2 ** (n + 1) - 1  # noqa: E731
        x = rng.normal(0, 1.0, (200, 4))
    elif case in ("quad_cc", "quad_aniso"):              # config 4's operator and knots
        from sgmethods_b200.tp_inteprolants import TPPwQuadraticInterpolator as op
        S = mis.smolyak_mid_set(4, 3) if case == "quad_cc" else mis.aniso_smolyak_mid_set(5, 4, [1, 1.5, 2, 3])
        knots = nodes_1d.cc_nodes
        lev = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
        x = rng.uniform(-1.3, 1.0, (200, S.shape[1]))    # extrapolates to the left only
    elif case == "lag_leja":                             # config 2's wording: Leja knots, m = nu + 1
        from sgmethods_b200.leja import leja_knots
        from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator as op
        import os
        with np.load(os.path.join(os.path.dirname(__file__), "golden", "lag_leja_gauss.npz")) as z:
            knots = leja_knots(z["leja_X"])
        S, lev = mis.smolyak_mid_set(4, 5), (lambda nu: nu + 1)
        x = rng.normal(0, 1.5, (200, 5))
    else:                                                # example/0 as shipped: Lagrange on cc knots
        from sgmethods_b200.tp_inteprolants import TPLagrangeInterpolator as op
        S, knots = mis.aniso_smolyak_mid_set(4, 4, [1, 2, 2, 4]), nodes_1d.cc_nodes
        lev = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
        x = rng.uniform(-1.0, 1.0, (200, 4))
    it = SGInterpolant(S, knots, lev, verbose=False) if op is None else \
        SGInterpolant(S, knots, lev, tp_interpolant=op, verbose=False)
    hp = HierarchicalPlan(it)
    f = rng.normal(size=(it.num_nodes, 2))
    ref = sg_oracle.interpolate(
        sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                   it.active_tp_nodes_list, it.map_tp_to_SG), x, f, squeeze=False,
        kind={0: "linear", 1: "quadratic", 3: "lagrange"}[it._kind])
    return it, hp, x, f, ref


def test_hierarchical_form_rejects_what_it_cannot_represent():
    S = mis.smolyak_mid_set(2, 2)
    with pytest.raises(NotImplementedError):              # cc knots with n+1 points: not nested
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: n + 1, verbose=False))
    with pytest.raises(NotImplementedError):              # lev2knots(0) != 1
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: 2 ** (n + 1) + 1, verbose=False))
    from sgmethods_b200.tp_inteprolants import TPPwCubicInterpolator, TPPwQuadraticInterpolator
    dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
    with pytest.raises(NotImplementedError):              # pw-cubic: patches of doubling families are not nested
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: 3 * 2 ** n + 1,
                                       tp_interpolant=TPPwCubicInterpolator, verbose=False))
    with pytest.raises(NotImplementedError):              # pw-quadratic needs 1, 3, 5, 9, ... knots
        HierarchicalPlan(SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1,
                                       tp_interpolant=TPPwQuadraticInterpolator, verbose=False))
    assert HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, dbl, tp_interpolant=TPPwQuadraticInterpolator,
                                          verbose=False)).n_terms == 11     # 6 multi-indices, level 1 = 2 families


def test_lagrange_hierarchical_form_is_nan_where_the_reference_is():
    """The reference's barycentric formula gives NaN when a coordinate equals a knot of an
    active direction (tp_inteprolants.py:346); the hierarchical tables reproduce that."""
    it, hp, x, f, _ = _case("lag_leja")
    x = x[:8].copy()
    x[0, 1] = 0.0                                       # the level-0 knot, a knot of every level
    x[1, 3] = hp.rank_knots[-1][0]                      # a knot of the finest level only
    x[2, 4] = hp.rank_knots[1][-1]
    ref = sg_oracle.interpolate(
        sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                   it.active_tp_nodes_list, it.map_tp_to_SG), x, f, squeeze=False,
        kind="lagrange")
    got = emulate(hp, x, f)
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(ref[:3]).all() and not np.isnan(ref[3:]).any()


@pytest.mark.parametrize("name", ["ml_lin_nf1", "ml_quad_nf3", "ml_lin_grow"])
def test_multilevel_hierarchical_tables_reproduce_the_reference_driver_outputs(name):
    """MLInterpolant in hierarchical-surplus form: one evaluation on the finest set whose
    surplus table is assembled level by level (multilevel_interpolant._hierarchical), emulated
    in numpy and compared with the outputs of the reference's own MLInterpolant.interpolate
    (tests/golden/ml_*.npz) -- also where the parametric dimension grows with the level."""
    from conftest import load_golden
    from golden_cases import LEV2KNOTS, ML_CASES, knot_family, ml_mid_set
    from sgmethods_b200 import tp_inteprolants as tpi
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    fam, l2k, kind, levels, NF, P, seed = ML_CASES[name]
    op = {"linear": tpi.TPPwLinearInterpolator, "quadratic": tpi.TPPwQuadraticInterpolator}[kind]
    g = load_golden(name)
    seq = [SGInterpolant(np.asarray(ml_mid_set(mis, name, k)), knot_family(fam, nodes_1d),
                         LEV2KNOTS[l2k], tp_interpolant=op, verbose=False) for k in range(levels)]
    mli = MLInterpolant(seq)
    hps, moves = mli._hierarchical()
    top = hps[-1]
    alpha = np.full((top.M, NF), np.nan)
    for k, hp in enumerate(hps):                       # block k = samples on the grid of seq[k]
        dst, src = moves[k]
        alpha[dst] = surpluses(hp, g[f"ml_{levels - 1 - k}"])[src]
    assert not np.isnan(alpha).any()                   # every basis tuple got its surpluses
    yy = g["yy"]
    out = np.zeros((yy.shape[0], NF))
    for t in range(top._term_fam.shape[0]):
        w, off = np.ones(yy.shape[0]), np.zeros(yy.shape[0], dtype=np.int64)
        act = np.flatnonzero(top._term_fam[t] >= 0)
        sizes = [int((top.fam_newrank_list[top._term_fam[t, d]] >= 0).sum()) for d in act]
        for a, d in enumerate(act):
            phi, j = basis_table(top, top._term_fam[t, d], yy[:, d])
            w, off = w * phi, off + j * int(np.prod(sizes[:a], dtype=np.int64))
        out += alpha[top._maps_f[top._map_off[t] + off]] * w[:, None]
    ref = np.asarray(g["out"]).reshape(out.shape)
    assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref))


def _random_downward_closed_set(rng, N, max_level, size):
    """A random downward closed multi-index set, lexicographically sorted (what the
    constructor requires): grown from the origin by admissible neighbours."""
    members = {(0,) * N}
    for _ in range(size * 4):
        if len(members) >= size:
            break
        base = list(sorted(members)[rng.integers(len(members))])
        d = int(rng.integers(N))
        base[d] += 1
        if base[d] > max_level:
            continue
        cand = tuple(base)
        ok = all(tuple(cand[j] - (1 if j == i else 0) for j in range(N)) in members
                 for i in range(N) if cand[i] > 0)
        if ok:
            members.add(cand)
    return np.array(sorted(members), dtype=np.int64)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("kind", ["linear", "quadratic", "lagrange"])
def test_random_downward_closed_sets_all_operators(kind, seed):
    """Random (not Smolyak-shaped) downward closed sets: the hierarchical tables and the group
    tables of the grouped kernel reproduce the oracle's combination formula to 1e-12."""
    from sgmethods_b200 import tp_inteprolants as tpi
    rng = np.random.default_rng(100 + seed)
    N = int(rng.integers(2, 6))
    S = _random_downward_closed_set(rng, N, 4 if kind != "lagrange" else 3, int(rng.integers(8, 40)))
    if kind == "linear":
        op, knots, lev = tpi.TPPwLinearInterpolator, nodes_1d.unbounded_nodes_nested, (lambda n: 2 ** (n + 1) - 1)
        x = rng.normal(0, 1.5, (120, N))
    elif kind == "quadratic":
        op, knots, lev = tpi.TPPwQuadraticInterpolator, nodes_1d.cc_nodes, (lambda n: np.where(n == 0, 1, 2 ** n + 1))
        x = rng.uniform(-1.2, 1.0, (120, N))
    else:
        op, knots, lev = tpi.TPLagrangeInterpolator, nodes_1d.cc_nodes, (lambda n: np.where(n == 0, 1, 2 ** n + 1))
        x = rng.uniform(-1.0, 1.0, (120, N))
    it = SGInterpolant(S, knots, lev, tp_interpolant=op, verbose=False)
    hp = HierarchicalPlan(it)
    f = rng.normal(size=(it.num_nodes, 2))
    ref = sg_oracle.interpolate(
        sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                   it.active_tp_nodes_list, it.map_tp_to_SG), x, f, squeeze=False, kind=kind)
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(emulate(hp, x, f) - ref)) <= 1e-12 * scale
    assert np.max(np.abs(emulate_groups(hp, x, f) - ref)) <= 1e-12 * scale


@pytest.mark.parametrize("kind", ["linear", "quadratic", "lagrange"])
def test_coarse_operator_weights_are_the_one_dimensional_operator(kind):
    """hierarchical._coarse_operator: sum_j w_j u(parent_j) is the level-(r-1) operator of the
    oracle applied to u at a new knot of level r (also outside the coarse knots for pw-linear,
    where the operator extrapolates from the end pair)."""
    from sgmethods_b200 import _lib
    from sgmethods_b200.hierarchical import _coarse_operator
    rng = np.random.default_rng(23)
    code = {"linear": _lib.KIND_LINEAR, "quadratic": _lib.KIND_QUADRATIC, "lagrange": _lib.KIND_LAGRANGE}[kind]
    knots = nodes_1d.unbounded_nodes_nested if kind == "linear" else nodes_1d.cc_nodes
    sizes = (3, 7, 15) if kind == "linear" else (3, 5, 9)
    for m_c, m_f in zip(sizes[:-1], sizes[1:]):
        gc, gf = np.asarray(knots(m_c), dtype=float), np.asarray(knots(m_f), dtype=float)
        old = np.array([int(np.argmin(np.abs(gf - v))) for v in gc])
        new = np.setdiff1d(np.arange(m_f), old)
        par, wts = _coarse_operator(code, gf, old, new)
        u = rng.normal(size=(m_f, 1))
        mine = sum(w[:, None] * u[p] for p, w in zip(par, wts))
        ref = sg_oracle.OPERATORS[kind]((gc,), u[old].reshape(m_c, 1), gf[new].reshape(-1, 1))
        assert np.max(np.abs(mine - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))


def test_multilevel_hierarchical_form_rejects_what_it_cannot_represent():
    from sgmethods_b200 import tp_inteprolants as tpi
    from sgmethods_b200.multilevel_interpolant import MLInterpolant
    dbl = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
    lin = [SGInterpolant(mis.smolyak_mid_set(k, 2), nodes_1d.cc_nodes, dbl, verbose=False) for k in range(3)]
    hps, moves = MLInterpolant(lin)._hierarchical()
    assert sum(m[0].size for m in moves) == hps[-1].M          # every surplus comes from exactly one level
    mixed = [lin[0], SGInterpolant(mis.smolyak_mid_set(1, 2), nodes_1d.cc_nodes, dbl,
                                   tp_interpolant=tpi.TPPwQuadraticInterpolator, verbose=False)]
    with pytest.raises(NotImplementedError):                   # levels with different operators
        MLInterpolant(mixed)._hierarchical()
    not_nested = [SGInterpolant(mis.smolyak_mid_set(2, 2), nodes_1d.cc_nodes, dbl, verbose=False),
                  SGInterpolant(mis.tensor_product_mid_set(1, 2), nodes_1d.cc_nodes, dbl, verbose=False)]
    with pytest.raises(NotImplementedError):                   # Lambda_0 is not a subset of Lambda_1
        MLInterpolant(not_nested)._hierarchical()
    other_knots = [lin[2], SGInterpolant(mis.smolyak_mid_set(3, 2), nodes_1d.equispaced_nodes, dbl, verbose=False)]
    with pytest.raises(NotImplementedError):                   # the levels do not share their knots
        MLInterpolant(other_knots)._hierarchical()
