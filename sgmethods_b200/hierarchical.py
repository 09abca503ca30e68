"""Work-reduced evaluation of a sparse-grid interpolant on NESTED knots: the
hierarchical-surplus form, for the pw-linear, pw-quadratic
and Lagrange operators.  ``SGInterpolant.interpolate`` uses it through ``method="auto"`` once
it has been validated against the exact kernel on the call at hand
(sparse_grid_interpolant.SGInterpolant._auto_candidate, _auto_verdict); ``method="combination"`` keeps the reference's
combination formula bit for bit.

Let U^0, U^1, ... be the reference's 1-D operators of one direction (level r has m_r knots,
m_0 = 1).  If the knots are nested and U^r reproduces the range of U^{r-1}, then
(U^r - U^{r-1}) u = sum_{new knots z of level r} s_z L^r_z, with L^r_z the level-r cardinal
function of z and s_z = (u - U^{r-1} u)(z) the hierarchical surplus, and the sparse-grid
interpolant of a downward closed set Lambda is

    I[u](x) = sum_{nu in Lambda} sum_{z in new(nu)} alpha_{nu,z} prod_{d: nu_d > 0} L^{nu_d}_{z_d}(x_d).

The new knots of a level are split into FAMILIES such that at most one member of a family
has a non-zero cardinal function at any x; a multi-index then expands into one "pseudo
multi-index" per tuple of families, and each of those costs ONE table read and at most k
multiplies per point (config 2: 4845 gathers per point instead of the 50 049 corner gathers
of the combination formula, reference loop sgmethods/sparse_grid_interpolant.py:269-278):

* pw-linear (tp_inteprolants.py:38-53): every bracket of level r holds exactly one new
  knot; one family per level, L = the hat of the new knot of the bracket of x (linear
  extrapolation outside the knots exactly like scipy's RegularGridInterpolator);
* pw-quadratic (tp_inteprolants.py:100-136) with m_r = 2 m_{r-1} - 1: the new knots are the
  middle knots of the 3-knot patches -- one family, L = the middle parabola of the patch of
  x; the first refinement 1 -> 3 knots adds BOTH ends of the single patch: two families of
  one knot each.  A level-(r-1) patch is the union of two level-r patches, so the ranges
  are nested (also for the reference's extrapolation to the left; a point right of the last
  knot sets the same flag as the exact kernel -> IndexError);
* Lagrange (tp_inteprolants.py:308-360): every cardinal polynomial is non-zero almost
  everywhere, so every new knot is its own family and L = prod_{i != z}(x - g_i)/(g_z - g_i)
  over the knots of its level (one family per level for Leja sequences with m_r = r + 1: 4845
  instead of 58 905 node contributions per point on config 2 with Gaussian-Leja knots).  The
  reference's barycentric formula returns NaN when a coordinate of x equals a knot of an
  active direction; the kernel reproduces that (the basis value is NaN on a knot of the
  family's level -- the finest level of every direction belongs to a multi-index of Lambda).
* pw-cubic (tp_inteprolants.py:206-253): NOT available.  Its 4-knot patches [g_3j, g_3j+3]
  are unions of finer patches only when the knot count triples (m_r = 3 m_{r-1} - 2); for the
  doubling families the reference is used with, U^r does not reproduce the range of U^{r-1}
  and the telescoping identity above does not hold.

``alpha`` are the hierarchical surpluses of the nodal values, obtained by the usual
dimension-wise hierarchisation (on the GPU, ``sgm_hier_apply_batches``: value minus the
coarser operator's interpolant at the new knot: 2 parents for pw-linear, 3 for pw-quadratic,
m_{r-1} for Lagrange).

The two forms are the same function in exact arithmetic (also where they extrapolate).  In
floating point this one avoids the large alternating combination sum, so it differs from
the reference by about the REFERENCE's own rounding error; the tests require 1e-12 against
the oracle on config-2-, 3-, 4- and 5-shaped interpolants with random and smooth data, and "not
worse than the oracle" against a long-double evaluation.

Not available (``NotImplementedError`` -> the exact kernels are used): pw-cubic and
user-defined operators, knot families that are not nested or whose levels do not have the
structure above (e.g. pw-linear with ``lev2knots = n + 1``), ``lev2knots(0) != 1``, sets that
are not downward closed.
"""
import numpy as np

from . import _lib
from .sparse_grid_interpolant import NODE_TOL


def _coarse_operator(kind, g, old, i):
    """The level-(r-1) operator of one direction applied at new knots of level r: ``g`` knots of
    level r, ``old`` positions in ``g`` of the level-(r-1) knots, ``i`` positions of the new knots
    (one per child).  Returns (parent positions in ``g``, weights), lists of arrays like ``i``:
    the coarser operator's interpolant at g[i] is sum_j weights[j] * u(g[parents[j]])."""
    n = i.size
    if old.size == 1:                                  # the coarser level is the constant
        return [np.full(n, old[0])], [np.ones(n)]
    G, x = g[old], g[i]
    if kind == _lib.KIND_LINEAR:
        k = np.clip(np.searchsorted(old, i), 1, old.size - 1)    # ends: extrapolate from the end pair
        ia, ib = old[k - 1], old[k]
        t = (x - g[ia]) / (g[ib] - g[ia])
        return [ia, ib], [1.0 - t, t]
    if kind == _lib.KIND_QUADRATIC:                    # patch of the coarser level (tp_inteprolants.py:111-136)
        H = (old.size + 1) // 2
        j = np.clip(np.searchsorted(G[0::2], x, side="left") - 1, 0, H - 2)
        x0, x1, x2 = G[2 * j], G[2 * j + 1], G[2 * j + 2]
        return [old[2 * j], old[2 * j + 1], old[2 * j + 2]], \
            [((x - x1) * (x - x2)) / ((x0 - x1) * (x0 - x2)),
             ((x - x0) * (x - x2)) / ((x1 - x0) * (x1 - x2)),
             ((x - x0) * (x - x1)) / ((x2 - x0) * (x2 - x1))]
    par, wts = [], []                                  # Lagrange: every coarser knot
    for a in range(old.size):
        w = np.ones(n)
        for b in range(old.size):
            if b != a:
                w = w * ((x - G[b]) / (G[a] - G[b]))
        par.append(np.full(n, old[a]))
        wts.append(w)
    return par, wts


class HierarchicalPlan:
    """Hierarchical-surplus evaluator of an ``SGInterpolant`` with nested knots (pw-linear,
    pw-quadratic or Lagrange operator)."""

    def __init__(self, interp):
        kind = interp._kind
        if kind not in (_lib.KIND_LINEAR, _lib.KIND_QUADRATIC, _lib.KIND_LAGRANGE):
            raise NotImplementedError("the hierarchical form does not exist for the pw-cubic operator "
                                      "on doubling knot families (its patches are not nested)")
        self.it = interp
        mid = np.asarray(interp.mid_set)
        card, N = mid.shape
        m_all = np.array([interp.lev2knots(row).astype(int) for row in mid], dtype=np.int64)
        ms = np.unique(m_all)
        if ms[0] != 1:
            raise NotImplementedError("hierarchical form needs lev2knots(0) == 1")
        # rank r = position of the node count in the increasing sequence 1 = m_0 < m_1 < ...;
        # one level per node count (lev2knots strictly increasing on the levels that occur)
        pairs = np.unique(m_all.ravel() * (int(mid.max()) + 1) + mid.ravel())
        if pairs.size != ms.size:
            raise NotImplementedError("lev2knots must be strictly increasing")
        ranks = np.searchsorted(ms, m_all).astype(np.int32)                   # (card, N)
        knots = [np.atleast_1d(np.asarray(interp.knots(np.int64(m)), dtype=np.float64)) for m in ms]
        if abs(knots[0][0]) > NODE_TOL:
            raise NotImplementedError("knots(1) must be 0")
        self.rank_knots = knots
        R = len(ms) - 1
        # nestedness, and the families of new knots of every rank >= 1: (positions in the rank's
        # knot vector, sgm_hier_basis); at most one member of a family is "active" at any x
        up, is_new_of = [None], [None]
        fam_rank, fam_pos, fam_basis = [], [], []
        first_fam, n_sub = np.zeros(R + 1, dtype=np.int64), np.ones(R + 1, dtype=np.int64)
        for r in range(1, R + 1):
            g, gp = knots[r], knots[r - 1]
            if np.any(np.diff(g) <= 0):
                raise NotImplementedError("knots must be strictly increasing")
            pos = np.searchsorted(g, gp)
            pos = np.clip(pos, 0, g.size - 1)
            pos = np.where((pos > 0) & (np.abs(g[np.maximum(pos - 1, 0)] - gp) < np.abs(g[pos] - gp)),
                           pos - 1, pos)
            if np.any(np.abs(g[pos] - gp) >= NODE_TOL) or np.unique(pos).size != gp.size:
                raise NotImplementedError("knot families are not nested")
            is_new = np.ones(g.size, dtype=bool)
            is_new[pos] = False
            new = np.flatnonzero(is_new)
            first_fam[r] = len(fam_rank)
            if kind == _lib.KIND_LINEAR:
                if np.any(is_new[:-1] == is_new[1:]):
                    raise NotImplementedError("every bracket must contain exactly one new knot")
                fams = [(new, _lib.HB_HAT)]
            elif kind == _lib.KIND_QUADRATIC:
                if g.size == 3 and gp.size == 1 and pos[0] == 1:
                    fams = [(new[:1], _lib.HB_QUAD_END), (new[1:], _lib.HB_QUAD_END)]
                elif g.size % 2 == 1 and gp.size >= 3 and np.array_equal(new, np.arange(1, g.size, 2)):
                    fams = [(new, _lib.HB_QUAD_MID)]
                else:
                    raise NotImplementedError("pw-quadratic hierarchical form needs knot counts 1, 3, "
                                              "5, 9, ... (m_r = 2 m_{r-1} - 1)")
            else:
                fams = [(new[i:i + 1], _lib.HB_LAG_ONE) for i in range(new.size)]
            n_sub[r] = len(fams)
            for p_, b_ in fams:
                fam_rank.append(r)
                fam_pos.append(p_.astype(np.int64))
                fam_basis.append(b_)
            up.append(pos)                       # knot i of rank r-1 = knot up[r][i] of rank r
            is_new_of.append(is_new)
        n_fam = len(fam_rank)
        fam_rank = np.array(fam_rank, dtype=np.int64)
        # identity of a family that does not depend on this plan's numbering (multilevel
        # interpolants match the basis tuples of different plans): node count + new positions
        self.fam_key = [(int(ms[r]), tuple(int(v) for v in p_)) for r, p_ in zip(fam_rank, fam_pos)]
        self.fam_rank, self.fam_basis = fam_rank, np.array(fam_basis, dtype=np.int32)
        self.fam_knots_list = [knots[r] for r in fam_rank]
        self.fam_newrank_list = []
        for r, p_ in zip(fam_rank, fam_pos):
            nr = np.full(knots[r].size, -1, dtype=np.int32)
            nr[p_] = np.arange(p_.size, dtype=np.int32)
            self.fam_newrank_list.append(nr)
        # entry q = family + 1 (0: inactive direction)
        q_rank = np.concatenate(([0], fam_rank)).astype(np.int64)
        q_ext = np.array([1] + [p_.size for p_ in fam_pos], dtype=np.int64)
        q_pos_off = np.concatenate(([0, 0], np.cumsum([p_.size for p_ in fam_pos]))).astype(np.int64)
        q_pos_cat = np.concatenate(fam_pos) if n_fam else np.zeros(0, dtype=np.int64)
        # position of every knot of every rank in the FINEST knot vector: an integer name of
        # the knot that is the same on all levels (the families are nested)
        fin = [None] * (R + 1)
        fin[R] = np.arange(knots[R].size, dtype=np.int64)
        for r in range(R - 1, -1, -1):
            fin[r] = fin[r + 1][up[r + 1]]
        fin_off = np.concatenate(([0], np.cumsum([f.size for f in fin]))).astype(np.int64)
        fin_cat = np.concatenate(fin)                        # fin_cat[fin_off[r] + i] = fin[r][i]
        c0 = int(fin[0][0])                                  # name of the one level-0 knot
        # a node is named by sum_d mult[d] * (finest position of its coordinate in direction d),
        # 64-bit wrap-around arithmetic with random odd multipliers: equal names <=> equal nodes
        # up to a 2^-64 collision chance per pair
        mult = np.random.default_rng(0x5eed).integers(1, 2 ** 62, N, dtype=np.int64) * 2 + 1

        # ---- downward closed: lowering any active direction by one level gives a member
        rt_idx, rd_idx = np.nonzero(ranks > 0)
        with np.errstate(over="ignore"):
            tkeys = (ranks.astype(np.int64) * mult).sum(axis=1)
            stk = np.sort(tkeys)
            lower = tkeys[rt_idx] - mult[rd_idx]
        hit = np.clip(np.searchsorted(stk, lower), 0, max(card - 1, 0))
        if card and not np.array_equal(stk[hit], lower):
            raise NotImplementedError("the multi-index set must be downward closed")

        # ---- pseudo multi-indices: one per tuple of families of a multi-index's levels, in
        # lexicographic order of the entry ids (last direction fastest within a multi-index)
        if np.all(n_sub == 1):
            pent = np.where(ranks > 0, first_fam[ranks] + 1, 0).astype(np.int64)
        else:
            cnt = n_sub[ranks].prod(axis=1)
            if cnt.sum() > (1 << 24):
                raise NotImplementedError("too many hierarchical basis tuples")
            term_of = np.repeat(np.arange(card), cnt)
            rem = np.arange(term_of.size, dtype=np.int64) - \
                np.concatenate(([0], np.cumsum(cnt)))[term_of]
            pent = np.zeros((term_of.size, N), dtype=np.int64)
            for d in range(N - 1, -1, -1):
                rk = ranks[term_of, d]
                radix = n_sub[rk]
                pent[:, d] = np.where(rk > 0, first_fam[rk] + rem % radix + 1, 0)
                rem //= radix
        self.n_terms = Tp = pent.shape[0]

        # ---- sparse form: active directions / entries per pseudo multi-index (padded)
        t_idx, d_idx = np.nonzero(pent > 0)                  # row-major: directions ascending
        k_t = np.bincount(t_idx, minlength=Tp)
        kmax = int(k_t.max()) if Tp else 0
        slot_of = np.arange(t_idx.size) - np.concatenate(([0], np.cumsum(k_t)))[t_idx]
        act_dim = np.zeros((Tp, max(kmax, 1)), dtype=np.int64)
        act_q = np.zeros((Tp, max(kmax, 1)), dtype=np.int64)            # 0 = padding
        act_dim[t_idx, slot_of] = d_idx
        act_q[t_idx, slot_of] = pent[t_idx, d_idx]
        ext = q_ext[act_q]                                              # new knots per slot

        # ---- hierarchical nodes: a pseudo multi-index has one node per tuple of new knots of
        # its families, numbered term by term in C order (pairwise distinct by construction)
        sizes = ext.prod(axis=1)
        map_off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
        self.M = int(map_off[-1])
        if self.M != interp.num_nodes:
            raise NotImplementedError("the hierarchical nodes do not coincide with the sparse grid")
        node_term = np.repeat(np.arange(Tp), sizes)
        local = np.arange(self.M, dtype=np.int64) - map_off[node_term]
        nidx = np.zeros((self.M, max(kmax, 1)), dtype=np.int64)         # new-knot index per slot
        rem = local.copy()
        for a in range(kmax - 1, -1, -1):
            e = ext[node_term, a]
            nidx[:, a] = rem % e
            rem //= e
        n_dim = act_dim[node_term]                                       # (M, kmax)
        n_q = act_q[node_term]
        n_rank = q_rank[n_q]
        n_pos = np.where(n_q > 0, q_pos_cat[np.minimum(q_pos_off[n_q] + nidx,
                                                       max(q_pos_cat.size - 1, 0))], 0) \
            if R else np.zeros_like(nidx)                                # index in knots[rank]
        n_fin = np.where(n_q > 0, fin_cat[fin_off[n_rank] + n_pos], c0)
        with np.errstate(over="ignore"):
            base_key = (np.full(N, c0, dtype=np.int64) * mult).sum()
        with np.errstate(over="ignore"):
            keys = base_key + ((n_fin - c0) * np.where(n_q > 0, mult[n_dim], 0)).sum(axis=1)
        order = np.argsort(keys, kind="stable")
        skeys = keys[order]
        if self.M > 1 and np.any(skeys[1:] == skeys[:-1]):
            raise RuntimeError("hierarchical nodes are not distinct")

        def lookup(probe):
            at = np.clip(np.searchsorted(skeys, probe), 0, self.M - 1)
            return np.where(skeys[at] == probe, order[at], -1)

        # ---- row of the sparse grid for every hierarchical node: the nodes of the interpolant's
        # own (nodal) tensor grids carry the same integer names
        tf, nn = interp._term_fam, interp._num_nodes
        rank_of_fam = np.searchsorted(ms, interp._fam_m)                 # family (node count) -> rank
        sg_key = np.full(interp.num_nodes, base_key, dtype=np.int64)
        Ta = tf.shape[0]
        at_idx, ad_idx = np.nonzero(tf >= 0)
        ka = np.bincount(at_idx, minlength=Ta)
        ka_max = int(ka.max()) if Ta else 0
        if ka_max:
            a_slot = np.arange(at_idx.size) - np.concatenate(([0], np.cumsum(ka)))[at_idx]
            a_dim = np.zeros((Ta, ka_max), dtype=np.int64)
            a_rank = np.zeros((Ta, ka_max), dtype=np.int64)
            a_ext = np.ones((Ta, ka_max), dtype=np.int64)
            a_dim[at_idx, a_slot] = ad_idx
            a_rank[at_idx, a_slot] = rank_of_fam[tf[at_idx, ad_idx]]
            a_ext[at_idx, a_slot] = nn[at_idx, ad_idx]
            a_sizes = np.diff(interp._map_off)
            e_term = np.repeat(np.arange(Ta), a_sizes)
            rem = np.arange(e_term.size, dtype=np.int64) - interp._map_off[e_term]
            with np.errstate(over="ignore"):
                ek = np.full(e_term.size, base_key, dtype=np.int64)
                for a in range(ka_max - 1, -1, -1):
                    e = a_ext[e_term, a]
                    i = rem % e
                    rem //= e
                    r = a_rank[e_term, a]
                    ek += np.where(r > 0, (fin_cat[fin_off[r] + i] - c0) * mult[a_dim[e_term, a]], 0)
            sg_key[interp._maps] = ek                                    # consistent across terms
        order_sg = np.argsort(sg_key, kind="stable")
        at = np.clip(np.searchsorted(sg_key[order_sg], keys), 0, max(interp.num_nodes - 1, 0))
        if not np.array_equal(sg_key[order_sg][at], keys):
            raise NotImplementedError("the hierarchical nodes do not coincide with the sparse grid")
        self.sel = order_sg[at].astype(np.int64)
        if np.unique(self.sel).size != self.M:
            raise NotImplementedError("the hierarchical nodes do not coincide with the sparse grid")

        # ---- hierarchisation batches: for every direction, finest rank first; the children of
        # a batch are the nodes whose knot in that direction is new on that rank, the parents the
        # nodes with the coarser operator's knots there (weights = its cardinal functions at the
        # child's knot).  A batch carries two parents per child: operators with more parents
        # emit several batches over the same children (their parents are not children of any
        # batch of this rank, so the order among them is free).
        self.batches = []
        if kmax:
            flat_dim, flat_rank = n_dim.ravel(), n_rank.ravel()
            code = np.where(flat_rank > 0, flat_dim * (R + 1) + flat_rank, -1)
            by_code = np.argsort(code, kind="stable")
            sc = code[by_code]
            for d in range(N):
                for r in range(int(ranks[:, d].max()), 0, -1):
                    lo, hi = np.searchsorted(sc, [d * (R + 1) + r, d * (R + 1) + r + 1])
                    if hi == lo:
                        continue
                    flat = by_code[lo:hi]
                    ch, a = flat // kmax, flat % kmax                    # node, its slot for d
                    ch_order = np.argsort(ch, kind="stable")
                    ch, a = ch[ch_order], a[ch_order]
                    g = knots[r]
                    old = np.flatnonzero(~is_new_of[r])
                    i = n_pos[ch, a]
                    par, wts = _coarse_operator(kind, g, old, i)         # lists of (n,) arrays
                    found = []
                    for ip in par:
                        with np.errstate(over="ignore"):
                            pn = lookup(keys[ch] + (fin[r][ip] - n_fin[ch, a]) * mult[d])
                        if np.any(pn < 0):
                            raise NotImplementedError("a hierarchical parent node is missing (set "
                                                      "not downward closed?)")
                        found.append(pn)
                    for j in range(0, len(found), 2):
                        two = j + 1 < len(found)
                        self.batches.append((
                            ch.astype(np.int32), found[j].astype(np.int32),
                            found[j + 1].astype(np.int32) if two else np.full(ch.size, -1, np.int32),
                            wts[j].astype(np.float64),
                            wts[j + 1].astype(np.float64) if two else np.zeros(ch.size)))

        # evaluation tables: one family per (rank, family of new knots) with the rank's full
        # knot vector (bracket / patch search) and the new-knot ranks; term tables are the
        # surplus blocks themselves
        self._fam_off = np.concatenate(([0], np.cumsum([k.size for k in self.fam_knots_list]))).astype(np.int64)
        self._fam_knots = np.concatenate(self.fam_knots_list) if n_fam else np.zeros(0)
        self._fam_newrank = np.concatenate(self.fam_newrank_list) if n_fam else np.zeros(0, np.int32)
        self._term_fam = (pent - 1).astype(np.int32)         # -1 where inactive
        self._map_off = map_off
        # the evaluation kernels address a term's surplus block with the FIRST active
        # direction fastest (offset prefixes are then shared by consecutive multi-indices);
        # the nodes are numbered with the last direction fastest: permutation
        stride_f = np.cumprod(np.concatenate((np.ones((Tp, 1), dtype=np.int64), ext[:, :-1]), axis=1),
                              axis=1)                        # F-order strides per slot
        f_index = (nidx * stride_f[node_term]).sum(axis=1)
        maps_f = np.empty(self.M, dtype=np.int32)
        maps_f[map_off[node_term] + f_index] = np.arange(self.M, dtype=np.int32)
        self._maps_f = maps_f
        self._device = {}

    # ------------------------------------------------------------------------------ device
    def _dev(self, device):
        from ._device import DevicePlan, torch_cuda
        torch = torch_cuda()
        if device not in self._device:
            dev = torch.device("cuda", device)
            plan = DevicePlan(_lib.KIND_HIER, self.it.N, self.M, np.ones(self._term_fam.shape[0]),
                              self._term_fam, self._map_off, self._maps_f,
                              self._fam_off, self._fam_knots, device=device,
                              fam_newrank=self._fam_newrank, fam_basis=self.fam_basis)
            # the batches concatenated: one ABI call hierarchises (sgm_hier_apply_batches)
            off = np.concatenate(([0], np.cumsum([b[0].size for b in self.batches]))).astype(np.int64)
            cat = [np.concatenate([b[i] for b in self.batches]) if self.batches else
                   np.zeros(0, dtype=np.int32 if i < 3 else np.float64) for i in range(5)]
            batches = {"off_host": off, "off_dev": torch.from_numpy(off).to(dev),
                       "arrays": [torch.from_numpy(a).to(dev) for a in cat]}
            sel = torch.from_numpy(self.sel).to(dev)
            self._device[device] = (plan, batches, sel)
        return self._device[device]

    def surpluses(self, f_dev):
        """Hierarchical surpluses (M, NF) of nodal values given in sparse-grid row order."""
        from ._device import torch_cuda
        torch = torch_cuda()
        plan, batches, sel = self._dev(f_dev.device.index)
        alpha = f_dev.index_select(0, sel).contiguous()
        stream = torch.cuda.current_stream(f_dev.device).cuda_stream
        ch, pa, pb, wa, wb = batches["arrays"]
        off = batches["off_host"]
        _lib.check(_lib.load().sgm_hier_apply_batches(
            alpha.data_ptr(), alpha.shape[1], alpha.stride(0), ch.data_ptr(), pa.data_ptr(),
            pb.data_ptr(), wa.data_ptr(), wb.data_ptr(), _lib.ptr(off, _lib.c_i64),
            batches["off_dev"].data_ptr(), off.size - 1, stream))
        return alpha

    def interpolate(self, x_new, f_on_SG):
        """Same call as ``SGInterpolant.interpolate(..., method="hierarchical")``."""
        return self.it.interpolate(x_new, f_on_SG, method="hierarchical")
