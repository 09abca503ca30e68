"""2+ GPU check of ShardedInterpolant (NCCL): points mode is bit-identical to one GPU, terms
mode agrees to rounding.  Run: torchrun --nproc-per-node 2 tools/dist_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.distributed import ShardedInterpolant
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
rank = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
it = SGInterpolant(mis.smolyak_mid_set(4, 8), nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
gen = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn((100003, 8), dtype=torch.float64, generator=gen).to(dev)
f = torch.randn((it.num_nodes, 3), dtype=torch.float64, generator=gen).to(dev)
single_exact = it.interpolate(x, f, method="combination").reshape(x.shape[0], -1)
single_auto = it.interpolate(x, f).reshape(x.shape[0], -1)
scale = float(single_exact.abs().max())
for mode, method, single, bitwise in (("points", None, single_auto, True), ("points", "combination", single_exact, True),
                                      ("terms", None, single_exact, False)):
    sh = ShardedInterpolant(it, mode=mode, method=method)
    full = sh.interpolate(x, f, gather="all")
    root = sh.interpolate(x, f, gather="root")
    err = float((full - single).abs().max() / scale)
    ok_root = (root is None) if dist.get_rank() else bool(torch.allclose(root, single, rtol=0, atol=1e-12 * scale))
    print(f"rank {dist.get_rank()} mode {mode} method {method}: rel dev from single-GPU {err:.2e}, root ok {ok_root}", flush=True)
    assert ok_root and (err == 0.0 if bitwise else err < 1e-12)
dist.barrier(); dist.destroy_process_group()
