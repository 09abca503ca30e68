"""Multi-GPU evaluation of one interpolant: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch on the B200 box, gloo in the CPU tests) for the plumbing.

The path shards naturally:

  * ``mode="points"`` (default): every evaluation point is independent, so each rank
    evaluates a contiguous block of rows of ``x_new`` with the full (replicated) tables and
    nodal values.  There is no data-path collective; the only exchange is the optional final
    gather of the (P/G, NF) output slabs (``gather="all"`` -> NCCL all-gather,
    ``gather="root"`` -> gather to rank 0, ``gather=None`` -> the result stays sharded).
    Results are bit-identical to the single-GPU run with the same ``method`` (the exact
    kernels and the hierarchical kernels both compute a point's value independently of the
    batch it is in, with a fixed summation order).
  * ``mode="terms"``: for few points and a huge combination set each rank evaluates a
    contiguous range of terms (balanced by corner count, reference order inside the range)
    at all points and the (P, NF) partial sums are combined with an NCCL all-reduce.  This
    re-associates the sum over terms (partial sums per rank), so it is NOT bit-identical:
    the deviation is bounded by eps * sum_t |c_t TP_t| (DESIGN.md "Arithmetic order").

The reference has no distributed code at all; this module only adds the sharding around the
same ``sgm_eval`` kernels.
"""
import numpy as np


def shard_bounds(n, world):
    """Contiguous near-equal blocks: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(int(n), int(world))
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


def balanced_term_ranges(cost, world):
    """Contiguous term ranges with near-equal total cost (prefix-sum split)."""
    cost = np.asarray(cost, dtype=np.float64)
    csum = np.concatenate(([0.0], np.cumsum(cost)))
    targets = csum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(csum, targets, side="left")
    return np.concatenate(([0], cuts, [cost.size])).astype(np.int64)


class ShardedInterpolant:
    """Evaluate ``interpolant`` (an SGInterpolant) on all ranks of a process group."""

    def __init__(self, interpolant, group=None, mode="points", _evaluator=None, method=None):
        """``method``: as in ``SGInterpolant.interpolate`` (points mode; None = its default,
        the validated work-reduced form); the term-sharded mode always evaluates the
        combination formula."""
        import torch.distributed as dist
        assert mode in ("points", "terms")
        self.it = interpolant
        self.group = group
        self.mode = mode
        self.method = method
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # test hook: a callable (x, f, t0, t1) -> (P, NF) tensor standing in for the CUDA
        # kernels so that the sharding / collective logic can be exercised with gloo on CPU
        self._evaluator = _evaluator
        self._term_plan = None
        if mode == "terms":
            radix = {0: 2, 1: 3, 2: 4}.get(interpolant._kind)
            nn = interpolant._num_nodes
            if radix is None:
                cost = np.prod(nn, axis=1, dtype=np.float64)
            else:
                cost = float(radix) ** (nn > 1).sum(axis=1)
            self.term_ranges = balanced_term_ranges(cost, self.world)

    # ------------------------------------------------------------------------------ pieces
    def _eval(self, x, f, t0=None, t1=None):
        if self._evaluator is not None:
            return self._evaluator(x, f, t0, t1)
        if t0 is None:
            out = self.it.interpolate(x, f, method=self.method)     # all checks inside
            return out.reshape(x.shape[0], f.shape[1])
        plan = self._terms_plan(x.device.index, t0, t1)
        # same checks as SGInterpolant.interpolate: operator constraints before the launch, the
        # row count of f against the grid, the pw-quadratic condition flag after it
        self.it._check_operator_constraints()
        if f.shape[0] < self.it.num_nodes:
            raise IndexError(f"f_on_SG has {f.shape[0]} rows, the sparse grid {self.it.num_nodes}")
        out = plan.eval(x, f)
        if plan.kind == 1 and plan.take_flags() & 1:
            raise IndexError("pw-quadratic interpolation: a point lies right of the last knot "
                             "(or is NaN)")
        return out

    def _terms_plan(self, device, t0, t1):
        from ._device import DevicePlan
        if self._term_plan is None:
            it = self.it
            lo, hi = it._map_off[t0], it._map_off[t1]
            self._term_plan = DevicePlan(
                it._kind, it.N, it.num_nodes, it._coeffs[t0:t1].astype(np.float64),
                it._term_fam[t0:t1], it._map_off[t0:t1 + 1] - lo, it._maps[lo:hi],
                it._fam_off, it._fam_knots, device=device)
        return self._term_plan

    # -------------------------------------------------------------------------- evaluation
    def interpolate(self, x_new, f_on_SG, gather="all"):
        """x_new: the FULL (P, >= N) point set (tensor on this rank's device; every rank
        passes the same array) -- or use ``interpolate_local`` when each rank already holds
        only its block.  Returns (P, NF) on every rank (gather="all"), on rank 0 only
        (gather="root", other ranks get None) or the local block (gather=None)."""
        import torch
        import torch.distributed as dist
        P = x_new.shape[0]
        if self.mode == "terms":
            t0, t1 = (int(v) for v in self.term_ranges[self.rank:self.rank + 2])
            part = self._eval(x_new, f_on_SG, t0, t1) if t1 > t0 else \
                torch.zeros((P, f_on_SG.shape[1]), dtype=torch.float64, device=x_new.device)
            if self.world > 1:
                if gather == "root":
                    dist.reduce(part, dst=0, op=dist.ReduceOp.SUM, group=self.group)
                    return part if self.rank == 0 else None
                dist.all_reduce(part, op=dist.ReduceOp.SUM, group=self.group)
            return part
        b = shard_bounds(P, self.world)
        lo, hi = int(b[self.rank]), int(b[self.rank + 1])
        local = self._eval(x_new[lo:hi], f_on_SG)
        return self.gather_blocks(local, P, gather)

    def interpolate_local(self, x_local, f_on_SG):
        """Each rank evaluates the rows it holds; no communication."""
        return self._eval(x_local, f_on_SG)

    def gather_blocks(self, local, P, gather="all"):
        """Assemble the per-rank row blocks (sizes from shard_bounds(P, world))."""
        import torch
        import torch.distributed as dist
        if gather is None or self.world == 1:
            return local
        b = shard_bounds(P, self.world)
        NF = local.shape[1]
        width = int((b[1:] - b[:-1]).max())
        padded = torch.zeros((width, NF), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        if gather == "root":
            bufs = [torch.empty_like(padded) for _ in range(self.world)] if self.rank == 0 else None
            dist.gather(padded, bufs, dst=0, group=self.group)
            if self.rank != 0:
                return None
            return torch.cat([bufs[r][:int(b[r + 1] - b[r])] for r in range(self.world)])
        full = torch.empty((self.world * width, NF), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(full, padded, group=self.group)
        if width * self.world == P:
            return full
        return torch.cat([full[r * width:r * width + int(b[r + 1] - b[r])]
                          for r in range(self.world)])
