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


@pytest.mark.parametrize("case", ["cc", "gauss", "aniso"])
def test_hierarchical_tables_reproduce_the_combination_formula(case):
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
    got = emulate(hp, x, f)
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_hierarchical_form_rejects_what_it_cannot_represent():
    S = mis.smolyak_mid_set(2, 2)
    with pytest.raises(NotImplementedError):              # cc knots with n+1 points: not nested
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: n + 1, verbose=False))
    with pytest.raises(NotImplementedError):              # lev2knots(0) != 1
        HierarchicalPlan(SGInterpolant(S, nodes_1d.cc_nodes, lambda n: 2 ** (n + 1) + 1, verbose=False))
