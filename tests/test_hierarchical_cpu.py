"""Host side of the opt-in hierarchical-surplus form (sgmethods_b200/hierarchical.py): the
hierarchisation batches and evaluation tables are checked on the CPU by a plain numpy
emulation of what the kernels do, against the oracle's combination-formula result."""
import numpy as np
import pytest

from oracle import sg_oracle
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.hierarchical import HierarchicalPlan
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant


def emulate(hp, x, f):
    """numpy twin of sgm_hier_apply + the SGM_HIER_PW_LINEAR evaluation."""
    alpha = np.array(f[hp.sel], dtype=float)
    for ch, pa, pb, wa, wb in hp.batches:
        coarse = wa[:, None] * alpha[pa]
        two = pb >= 0
        coarse[two] = coarse[two] + wb[two, None] * alpha[pb[two]]
        alpha[ch] -= coarse
    it = hp.it
    out = np.zeros((x.shape[0], f.shape[1]))
    for t in range(hp._term_fam.shape[0]):
        w = np.ones(x.shape[0])
        off = np.zeros(x.shape[0], dtype=np.int64)
        act = np.flatnonzero(hp._term_fam[t] >= 0)
        sizes = [int((hp._newrank[hp._term_fam[t, d] + 1] >= 0).sum()) for d in act]
        for a, d in enumerate(act):
            r = hp._term_fam[t, d] + 1
            g, nr = hp.rank_knots[r], hp._newrank[r]
            i = np.clip(np.searchsorted(g, x[:, d], side="right") - 1, 0, g.size - 2)
            y = (x[:, d] - g[i]) / (g[i + 1] - g[i])
            lo_new = nr[i] >= 0
            w = w * np.where(lo_new, 1 - y, y)
            j = np.where(lo_new, nr[i], nr[i + 1])
            off = off + j * int(np.prod(sizes[:a], dtype=np.int64))      # first direction fastest
        out += alpha[hp._maps_f[hp._map_off[t] + off]] * w[:, None]
    return out


def emulate_groups(hp, x, f):
    """numpy twin of eval_hier_groups_kernel (kernel H2) on the tables of
    sgm_host_hier_groups: same traversal, same operations."""
    from sgmethods_b200 import _lib
    alpha = np.array(f[hp.sel], dtype=float)
    for ch, pa, pb, wa, wb in hp.batches:
        coarse = wa[:, None] * alpha[pa]
        two = pb >= 0
        coarse[two] = coarse[two] + wb[two, None] * alpha[pb[two]]
        alpha[ch] -= coarse
    tf = hp._term_fam
    T, N = tf.shape
    n_fam = len(hp.rank_knots) - 1
    extent = np.array([(hp._newrank[r] >= 0).sum() for r in range(1, n_fam + 1)], dtype=np.int32)
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
        g, nr = hp.rank_knots[fam + 1], hp._newrank[fam + 1]
        i = np.clip(np.searchsorted(g, x[:, d], side="right") - 1, 0, g.size - 2)
        y = (x[:, d] - g[i]) / (g[i + 1] - g[i])
        lo_new = nr[i] >= 0
        phi[e] = np.where(lo_new, 1 - y, y)
        rank[e] = np.where(lo_new, nr[i], nr[i + 1])
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


@pytest.mark.parametrize("case", ["cc", "gauss", "aniso"])
def test_group_tables_of_the_grouped_kernel_reproduce_the_combination_formula(case):
    it, hp, x, f, ref = _case(case)
    got = emulate_groups(hp, x, f)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("case", ["cc", "gauss", "aniso"])
def test_hierarchical_tables_reproduce_the_combination_formula(case):
    it, hp, x, f, ref = _case(case)
    got = emulate(hp, x, f)
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))


def _case(case):
    rng = np.random.default_rng(3)
    if case == "cc":
        S, knots = mis.smolyak_mid_set(4, 3), nodes_1d.cc_nodes
        lev = lambda n: np.where(n == 0, 1, 2 ** n + 1)  # noqa: E731
        x = rng.uniform(-1.3, 1.3, (200, 3))             # includes extrapolation
    elif case == "gauss":
        S, knots = mis.smolyak_mid_set(3, 5), nodes_1d.unbounded_nodes_nested
        lev = lambda n: 2 ** (n + 1) - 1  # noqa: E731
        x = rng.normal(0, 2.0, (200, 5))                 # beyond the outermost knots too
    else:
        S, knots = mis.aniso_smolyak_mid_set(4, 4, [1, 1.5, 2, 4]), nodes_1d.unbounded_nodes_nested
        lev = lambda n: 2 ** (n + 1) - 1  # noqa: E731
        x = rng.normal(0, 1.0, (200, 4))
    it = SGInterpolant(S, knots, lev, verbose=False)
    hp = HierarchicalPlan(it)
    f = rng.normal(size=(it.num_nodes, 2))
    ref = sg_oracle.interpolate(
        sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                   it.active_tp_nodes_list, it.map_tp_to_SG), x, f, squeeze=False)
    return it, hp, x, f, ref


def test_hierarchical_form_rejects_what_it_cannot_represent():
    S = mis.smolyak_mid_set(2, 2)
    with pytest.raises(NotImplementedError):              # cc knots with n+1 points: not nested
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: n + 1, verbose=False))
    with pytest.raises(NotImplementedError):              # lev2knots(0) != 1
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: 2 ** (n + 1) + 1, verbose=False))
