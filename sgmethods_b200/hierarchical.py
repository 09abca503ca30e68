"""Work-reduced evaluation of the pw-linear sparse-grid interpolant on NESTED knots:
the hierarchical-surplus form.  ``SGInterpolant.interpolate``
uses it through ``method="auto"`` once it has been validated against the exact kernel on the
interpolant at hand (sparse_grid_interpolant._resolve_auto); ``method="combination"`` keeps
the reference's combination formula bit for bit.

With nested knot families U^0 c U^1 c ... in every direction (one node on level 0, and on
every finer level exactly one NEW knot in each bracket), the interpolant is

    I[u](x) = sum_{nu in Lambda}  alpha_{nu, j(x)} * prod_{d: nu_d > 0} phi_{nu_d}(x_d),

one table entry and at most k multiplies per multi-index (config 2: 4845 gathers per point
instead of the 50 049 corner gathers of the combination formula, reference loop
sgmethods/sparse_grid_interpolant.py:269-278).  ``phi_l(x_d)`` is the value at x_d of the
level-l hat function of the one new knot adjacent to x_d (linear extrapolation outside the
knots, exactly like scipy's RegularGridInterpolator), and ``alpha`` are the hierarchical
surpluses of the nodal values, obtained by the usual dimension-wise hierarchisation (on the
GPU, ``sgm_hier_apply_batches``).

The two forms are the same function in exact arithmetic (also where they extrapolate).  In
floating point this one avoids the large alternating combination sum, so it differs from
the reference by about the REFERENCE's own rounding error; the tests require 1e-12 against
the oracle on config-2-, 3- and 5-shaped interpolants with random and smooth data, and "not
worse than the oracle" against a long-double evaluation.

Not available (``NotImplementedError`` -> the exact kernels are used): other operators than
pw-linear, knot families that are not nested or have brackets without / with several new
knots (e.g. ``lev2knots = n + 1``), ``lev2knots(0) != 1``, sets that are not downward closed.
"""
import numpy as np

from . import _lib
from .sparse_grid_interpolant import NODE_TOL


class HierarchicalPlan:
    """Hierarchical-surplus evaluator of a pw-linear ``SGInterpolant`` with nested knots."""

    def __init__(self, interp):
        if interp._kind != _lib.KIND_LINEAR:
            raise NotImplementedError("the hierarchical form is implemented for pw-linear only")
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
        # new knots of every rank >= 1 and the checks the one-basis-per-level form needs
        self.rank_knots = knots
        newrank, newpos, up = [None], [None], [None]
        for r in range(1, len(ms)):
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
            if np.any(is_new[:-1] == is_new[1:]):
                raise NotImplementedError("every bracket must contain exactly one new knot")
            nr = np.full(g.size, -1, dtype=np.int32)
            nr[is_new] = np.arange(is_new.sum(), dtype=np.int32)
            newrank.append(nr)
            newpos.append(np.flatnonzero(is_new))
            up.append(pos)                       # knot i of rank r-1 = knot up[r][i] of rank r
        self._newrank, self._newpos = newrank, newpos
        # position of every knot of every rank in the FINEST knot vector: an integer name of
        # the knot that is the same on all levels (the families are nested)
        R = len(ms) - 1
        fin = [None] * (R + 1)
        fin[R] = np.arange(knots[R].size, dtype=np.int64)
        for r in range(R - 1, -1, -1):
            fin[r] = fin[r + 1][up[r + 1]]
        fin_off = np.concatenate(([0], np.cumsum([f.size for f in fin]))).astype(np.int64)
        fin_cat = np.concatenate(fin)                        # fin_cat[fin_off[r] + i] = fin[r][i]
        c0 = int(fin[0][0])                                  # name of the one level-0 knot
        newpos_off = np.concatenate(([0, 0], np.cumsum([p.size for p in newpos[1:]]))).astype(np.int64)
        newpos_cat = np.concatenate(newpos[1:]) if R else np.zeros(0, dtype=np.int64)
        n_new = np.array([0] + [p.size for p in newpos[1:]], dtype=np.int64)
        # a node is named by sum_d mult[d] * (finest position of its coordinate in direction d),
        # 64-bit wrap-around arithmetic with random odd multipliers: equal names <=> equal nodes
        # up to a 2^-64 collision chance per pair
        mult = np.random.default_rng(0x5eed).integers(1, 2 ** 62, N, dtype=np.int64) * 2 + 1

        # ---- the multi-index set in sparse form: active directions / ranks per term (padded)
        t_idx, d_idx = np.nonzero(ranks > 0)                 # row-major: directions ascending
        k_t = np.bincount(t_idx, minlength=card)
        kmax = int(k_t.max()) if card else 0
        slot_of = np.arange(t_idx.size) - np.concatenate(([0], np.cumsum(k_t)))[t_idx]
        act_dim = np.zeros((card, max(kmax, 1)), dtype=np.int64)
        act_rank = np.zeros((card, max(kmax, 1)), dtype=np.int64)       # 0 = padding
        act_dim[t_idx, slot_of] = d_idx
        act_rank[t_idx, slot_of] = ranks[t_idx, d_idx]
        ext = np.where(act_rank > 0, n_new[act_rank], 1)                # new knots per slot
        with np.errstate(over="ignore"):
            # downward closed: lowering any active direction by one level gives a member
            tkeys = (ranks.astype(np.int64) * mult).sum(axis=1)
            stk = np.sort(tkeys)
            lower = tkeys[t_idx] - mult[d_idx]
        hit = np.clip(np.searchsorted(stk, lower), 0, max(card - 1, 0))
        if card and not np.array_equal(stk[hit], lower):
            raise NotImplementedError("the multi-index set must be downward closed")

        # ---- hierarchical nodes: term nu has one node per tuple of new knots, numbered term
        # by term in C order (they are pairwise distinct by construction)
        sizes = ext.prod(axis=1)
        map_off = np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
        self.M = int(map_off[-1])
        if self.M != interp.num_nodes:
            raise NotImplementedError("the hierarchical nodes do not coincide with the sparse grid")
        node_term = np.repeat(np.arange(card), sizes)
        local = np.arange(self.M, dtype=np.int64) - map_off[node_term]
        nidx = np.zeros((self.M, max(kmax, 1)), dtype=np.int64)         # new-knot index per slot
        rem = local.copy()
        for a in range(kmax - 1, -1, -1):
            e = ext[node_term, a]
            nidx[:, a] = rem % e
            rem //= e
        n_dim = act_dim[node_term]                                       # (M, kmax)
        n_rank = act_rank[node_term]
        n_pos = np.where(n_rank > 0, newpos_cat[np.minimum(newpos_off[n_rank] + nidx,
                                                           max(newpos_cat.size - 1, 0))], 0) \
            if R else np.zeros_like(nidx)                                # index in knots[rank]
        n_fin = np.where(n_rank > 0, fin_cat[fin_off[n_rank] + n_pos], c0)
        with np.errstate(over="ignore"):
            base_key = (np.full(N, c0, dtype=np.int64) * mult).sum()
        with np.errstate(over="ignore"):
            keys = base_key + ((n_fin - c0) * np.where(n_rank > 0, mult[n_dim], 0)).sum(axis=1)
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
        # a batch are the nodes whose knot in that direction is new on that rank
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
                    old = np.flatnonzero(newrank[r] < 0)
                    i = n_pos[ch, a]
                    if old.size == 1:                          # coarser level is the constant
                        ia, ib = np.full(i.size, old[0]), np.full(i.size, -1)
                        wa, wb = np.ones(i.size), np.zeros(i.size)
                    else:
                        k = np.searchsorted(old, i)            # olds left of i
                        k = np.clip(k, 1, old.size - 1)        # ends: extrapolate from the end pair
                        ia, ib = old[k - 1], old[k]
                        tpar = (g[i] - g[ia]) / (g[ib] - g[ia])
                        wa, wb = 1.0 - tpar, tpar
                    with np.errstate(over="ignore"):
                        pa = lookup(keys[ch] + (fin[r][ia] - n_fin[ch, a]) * mult[d])
                        pb = lookup(keys[ch] + (fin[r][ib] - n_fin[ch, a]) * mult[d]) \
                            if old.size > 1 else np.full(ch.size, -1, dtype=np.int64)
                    if np.any(pa < 0) or (old.size > 1 and np.any(pb < 0)):
                        raise NotImplementedError("a hierarchical parent node is missing (set not "
                                                  "downward closed?)")
                    self.batches.append((ch.astype(np.int32), pa.astype(np.int32), pb.astype(np.int32),
                                         wa.astype(np.float64), wb.astype(np.float64)))

        # evaluation tables: families = full knot vectors of ranks >= 1 (bracket search) with
        # the new-knot ranks; term tables are the surplus blocks themselves
        n_fam = R                                            # family f = rank f + 1
        self._fam_off = np.concatenate(([0], np.cumsum([k.size for k in knots[1:]]))).astype(np.int64)
        self._fam_knots = np.concatenate(knots[1:]) if n_fam else np.zeros(0)
        self._fam_newrank = np.concatenate(newrank[1:]) if n_fam else np.zeros(0, np.int32)
        self._term_fam = (ranks - 1).astype(np.int32)        # -1 where rank 0
        self._map_off = map_off
        # the evaluation kernels address a term's surplus block with the FIRST active
        # direction fastest (offset prefixes are then shared by consecutive multi-indices);
        # the nodes are numbered with the last direction fastest: permutation
        stride_f = np.cumprod(np.concatenate((np.ones((card, 1), dtype=np.int64), ext[:, :-1]), axis=1),
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
            plan = DevicePlan(_lib.KIND_HIER_LINEAR, self.it.N, self.M, np.ones(self._term_fam.shape[0]),
                              self._term_fam, self._map_off, self._maps_f,
                              self._fam_off, self._fam_knots, device=device,
                              fam_newrank=self._fam_newrank)
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
