"""Work-reduced evaluation of the pw-linear sparse-grid interpolant on NESTED knots:
the hierarchical-surplus form.  OPT-IN -- the default
``SGInterpolant.interpolate`` keeps the reference's combination formula bit for bit.

With nested knot families U^0 c U^1 c ... in every direction (one node on level 0, and on
every finer level exactly one NEW knot in each bracket), the interpolant is

    I[u](x) = sum_{nu in Lambda}  alpha_{nu, j(x)} * prod_{d: nu_d > 0} phi_{nu_d}(x_d),

one table entry and at most k multiplies per multi-index (config 2: 4845 gathers per point
instead of the 50 049 corner gathers of the combination formula).  ``phi_l(x_d)`` is the
value at x_d of the level-l hat function of the one new knot adjacent to x_d (linear
extrapolation outside the knots, exactly like scipy's RegularGridInterpolator), and
``alpha`` are the hierarchical surpluses of the nodal values, obtained by the usual
dimension-wise hierarchisation (on the GPU, ``sgm_hier_apply``).

The two forms are the same function in exact arithmetic (also where they extrapolate).  In
floating point this one avoids the large alternating combination sum, so it differs from
the reference by about the REFERENCE's own rounding error (measured <= 1e-12 relative on
the named configs with smooth data, ~1e-14 with random data); it is therefore not used by
default and its parity bar in the tests is 1e-10 against the oracle plus "not worse than the
oracle" against a long-double evaluation.
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
        # rank r = position of the node count in the increasing sequence 1 = m_0 < m_1 < ...
        levels_of_m = {}
        for lev, m in zip(mid.reshape(-1), m_all.reshape(-1)):
            levels_of_m.setdefault(int(m), set()).add(int(lev))
        if any(len(v) > 1 for v in levels_of_m.values()):
            raise NotImplementedError("lev2knots must be strictly increasing")
        rank_of_m = {int(m): r for r, m in enumerate(ms)}
        ranks = np.vectorize(rank_of_m.get)(m_all).astype(np.int32)           # (card, N)
        if not _is_downward_closed(ranks):
            raise NotImplementedError("the multi-index set must be downward closed")
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

        # hierarchical nodes: term nu has one node per tuple of new knots (C order)
        n_fam = len(ms) - 1                                  # family f = rank f + 1
        fam_off_new = np.concatenate(([0], np.cumsum([p.size for p in newpos[1:]]))).astype(np.int64)
        new_knots = np.concatenate([knots[r][newpos[r]] for r in range(1, len(ms))]) \
            if n_fam else np.zeros(0)
        term_fam = (ranks - 1).astype(np.int32)              # -1 where rank 0
        grid = _lib.HostGrid(None, None, fam_off_new, new_knots, NODE_TOL, term_fam=term_fam)
        map_off, maps = grid.maps()
        self.M = grid.num_nodes
        if self.M != maps.size or not np.array_equal(maps, np.arange(self.M, dtype=np.int32)):
            raise RuntimeError("hierarchical nodes are not distinct")
        nodes = grid.nodes_dense()
        grid.close()
        sel = _lib.match_rows(nodes, interp.SG, NODE_TOL)
        if self.M != interp.num_nodes or np.any(sel < 0):
            raise NotImplementedError("the hierarchical nodes do not coincide with the sparse grid")
        self.sel = sel.astype(np.int64)

        # level (rank) and knot position of every node in every direction
        node_rank = np.zeros((self.M, N), dtype=np.int32)
        node_pos = np.zeros((self.M, N), dtype=np.int32)     # index in knots[rank]
        for t in range(card):
            lo, hi = map_off[t], map_off[t + 1]
            act = np.flatnonzero(ranks[t] > 0)
            shape = [newpos[ranks[t, d]].size for d in act]
            node_rank[lo:hi] = ranks[t]
            if act.size:
                grids = np.indices(shape).reshape(act.size, -1)
                for a, d in enumerate(act):
                    node_pos[lo:hi, d] = newpos[ranks[t, d]][grids[a]]
        # integer key of a node: the finest-level positions of its coordinates, hashed with
        # random odd multipliers (collisions are excluded by comparing the rows afterwards)
        node_fin = np.empty((self.M, N), dtype=np.int64)
        for r in range(R + 1):
            sel_r = node_rank == r
            node_fin[sel_r] = fin[r][node_pos[sel_r]]
        mult = np.random.default_rng(0x5eed).integers(1, 2 ** 62, N, dtype=np.int64) * 2 + 1
        with np.errstate(over="ignore"):
            keys = (node_fin * mult).sum(axis=1)
        order = np.argsort(keys, kind="stable")
        skeys = keys[order]

        def find(ch, d, fin_pos):
            """Rows of the nodes that equal node ch except for knot `fin_pos` in direction d."""
            with np.errstate(over="ignore"):
                probe = keys[ch] + (fin_pos - node_fin[ch, d]) * mult[d]
            at = np.clip(np.searchsorted(skeys, probe), 0, self.M - 1)
            par = order[at]
            ok = skeys[at] == probe
            if ok.all():                                       # exact check of the hashed match
                want = node_fin[ch].copy()
                want[:, d] = fin_pos
                ok = (node_fin[par] == want).all(axis=1)
            if not ok.all():
                raise NotImplementedError("a hierarchical parent node is missing (set not "
                                          "downward closed?)")
            return par

        # hierarchisation batches: for every direction, finest rank first
        self.batches = []
        for d in range(N):
            for r in range(int(ranks[:, d].max()), 0, -1):
                ch = np.flatnonzero(node_rank[:, d] == r)
                if ch.size == 0:
                    continue
                g = knots[r]
                old = np.flatnonzero(newrank[r] < 0)
                i = node_pos[ch, d]
                if old.size == 1:                              # coarser level is the constant
                    ia, ib = np.full(i.size, old[0]), np.full(i.size, -1)
                    wa, wb = np.ones(i.size), np.zeros(i.size)
                else:
                    k = np.searchsorted(old, i)                # olds left of i
                    k = np.clip(k, 1, old.size - 1)            # ends: extrapolate from the end pair
                    ia, ib = old[k - 1], old[k]
                    tpar = (g[i] - g[ia]) / (g[ib] - g[ia])
                    wa, wb = 1.0 - tpar, tpar
                pa = find(ch, d, fin[r][ia])
                pb = find(ch, d, fin[r][ib]) if old.size > 1 else np.full(ch.size, -1, dtype=np.int64)
                self.batches.append((ch.astype(np.int32), pa.astype(np.int32), pb.astype(np.int32),
                                     wa.astype(np.float64), wb.astype(np.float64)))

        # evaluation tables: families = full knot vectors of ranks >= 1 (bracket search) with
        # the new-knot ranks; term tables are the surplus blocks themselves (identity map)
        self._fam_off = np.concatenate(([0], np.cumsum([k.size for k in knots[1:]]))).astype(np.int64)
        self._fam_knots = np.concatenate(knots[1:]) if n_fam else np.zeros(0)
        self._fam_newrank = np.concatenate(newrank[1:]) if n_fam else np.zeros(0, np.int32)
        self._term_fam = term_fam
        self._map_off = map_off
        # the evaluation kernels address a term's surplus block with the FIRST active
        # direction fastest (offset prefixes are then shared by consecutive multi-indices);
        # the grid builder numbers nodes with the last direction fastest: permutation
        maps_f = np.arange(self.M, dtype=np.int32)
        for t in range(card):
            lo, hi = map_off[t], map_off[t + 1]
            shape = [newpos[ranks[t, d]].size for d in np.flatnonzero(ranks[t] > 0)]
            if len(shape) > 1:
                maps_f[lo:hi] = lo + np.arange(hi - lo, dtype=np.int32).reshape(shape).ravel(order="F")
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
            batches = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in self.batches]
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
        lib = _lib.load()
        for ch, pa, pb, wa, wb in batches:
            _lib.check(lib.sgm_hier_apply(alpha.data_ptr(), alpha.shape[1], alpha.stride(0),
                                          ch.data_ptr(), pa.data_ptr(), pb.data_ptr(),
                                          wa.data_ptr(), wb.data_ptr(), ch.numel(), stream))
        return alpha

    def interpolate(self, x_new, f_on_SG):
        """Same call as ``SGInterpolant.interpolate`` (numpy or CUDA tensors in, same kind out)."""
        from ._device import torch_cuda
        from .sparse_grid_interpolant import _to_device
        torch = torch_cuda()
        as_torch = isinstance(x_new, torch.Tensor)
        device = x_new.device if as_torch and x_new.is_cuda else \
            torch.device("cuda", torch.cuda.current_device())
        x = _to_device(torch, x_new, device)
        f = _to_device(torch, f_on_SG, device)
        if x.dim() == 1:
            x = x.reshape(-1, 1)
        if x.shape[1] < self.it.N:
            x = torch.nn.functional.pad(x, (0, self.it.N - x.shape[1]))
        plan, _, _ = self._dev(device.index)
        out = plan.eval(x, self.surpluses(f[:self.M]))
        return out.squeeze() if as_torch else np.squeeze(out.cpu().numpy())


def _is_downward_closed(ranks):
    rows = {tuple(r) for r in ranks.tolist()}
    for r in rows:
        for n, v in enumerate(r):
            if v > 0 and r[:n] + (v - 1,) + r[n + 1:] not in rows:
                return False
    return True
