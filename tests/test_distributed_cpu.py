"""World-size-2 tests of the sharding / collective logic on CPU (gloo).  The CUDA kernels
are replaced by the oracle through the `_evaluator` test hook of ShardedInterpolant, so
what is tested here is the host logic: block bounds, term ranges, gather and reduce."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgmethods_b200.distributed import balanced_term_ranges, shard_bounds


def test_shard_bounds_and_term_ranges():
    assert shard_bounds(10, 4).tolist() == [0, 3, 6, 8, 10]
    assert shard_bounds(3, 8).tolist() == [0, 1, 2, 3, 3, 3, 3, 3, 3]
    assert shard_bounds(0, 2).tolist() == [0, 0, 0]
    r = balanced_term_ranges([16, 1, 1, 1, 1, 8, 8], 2)
    assert r.tolist() == [0, 3, 7] or r.tolist() == [0, 2, 7] or r[0] == 0 and r[-1] == 7
    r8 = balanced_term_ranges(np.ones(5), 8)
    assert r8[0] == 0 and r8[-1] == 5 and np.all(np.diff(r8) >= 0)


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sg_oracle
        from sgmethods_b200 import multi_index_sets as mis, nodes_1d
        from sgmethods_b200.distributed import ShardedInterpolant
        from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
        it = SGInterpolant(mis.smolyak_mid_set(3, 4), nodes_1d.unbounded_nodes_nested,
                           lambda n: 2 ** (n + 1) - 1, verbose=False)
        rng = np.random.default_rng(0)
        x = rng.normal(size=(101, 4))
        f = rng.normal(size=(it.num_nodes, 3))

        def oracle_eval(xt, ft, t0, t1):
            sl = slice(t0, t1)
            plan = sg_oracle.plan_from_tables(it.N, it.combination_coeffs[sl],
                                              it.active_tp_dims[sl], it.active_tp_nodes_list[sl],
                                              it.map_tp_to_SG[sl])
            return torch.from_numpy(sg_oracle.interpolate(plan, xt.numpy(), ft.numpy(),
                                                          squeeze=False))
        sh = ShardedInterpolant(it, mode=mode, _evaluator=oracle_eval)
        xt, ft = torch.from_numpy(x), torch.from_numpy(f)
        full = sh.interpolate(xt, ft, gather="all")
        root = sh.interpolate(xt, ft, gather="root")
        ref = sg_oracle.interpolate(
            sg_oracle.plan_from_tables(it.N, it.combination_coeffs, it.active_tp_dims,
                                       it.active_tp_nodes_list, it.map_tp_to_SG),
            x, f, squeeze=False)
        err = float(np.max(np.abs(full.numpy() - ref)) / np.max(np.abs(ref)))
        ok_root = (root is None) if rank else bool(np.max(np.abs(root.numpy() - ref)) <= 1e-12 * np.max(np.abs(ref)))
        local = None
        if mode == "points":
            local = tuple(sh.interpolate(xt, ft, gather=None).shape)
        q.put((rank, err, ok_root, local))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["points", "terms"])
def test_world_size_two_gloo(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + (0 if mode == "points" else 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, ok_root, local in res:
        assert err == 0.0 if mode == "points" else err <= 1e-12
        assert ok_root
    if mode == "points":
        assert res[0][3] == (51, 3) and res[1][3] == (50, 3)
