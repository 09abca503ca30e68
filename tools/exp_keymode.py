"""Locality key A/B in one process: SGM_KEY_MODE=0 (greedy multi-level key) against 1 (one bit per
direction first, the round-1 key), read when a plan is created."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def build(name):
    if name == "s68":
        return SGInterpolant(mis.smolyak_mid_set(6, 8), nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False), "normal"
    it, _, _, _, sampler = build_workload(name)
    return (it.interp_seq[-1] if name == "c4" else it), sampler


for name, P, NF in (("c3full", 100000, 1000), ("c4", 100000, 1000), ("c2", 10 ** 6, 1), ("c5", 10 ** 6, 1), ("s68", 10 ** 6, 1)):
    res = {}
    outs = {}
    for mode in ("0", "1"):
        os.environ["SGM_KEY_MODE"] = mode
        it, sampler = build(name)
        gen = torch.Generator(device=dev).manual_seed(5)
        x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64) if sampler == "normal" else \
            torch.rand((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 2 - 1
        f = torch.randn((it.num_nodes, NF), generator=gen, device=dev, dtype=torch.float64)
        hp = it.hierarchical()
        hplan = hp._dev(0)[0]
        alpha = hp.surpluses(f)
        out = torch.empty((P, NF), dtype=torch.float64, device=dev)
        res[mode] = timed(lambda: hplan.eval(x, alpha, out))
        if NF == 1:
            eplan = it.device_plan(0)
            res[mode + " exact"] = timed(lambda: eplan.eval(x, f), 2) if name != "c5" else float("nan")
        outs[mode] = out
    print(name, P, NF, {k: round(v, 3) for k, v in res.items()}, "identical:", torch.equal(outs["0"], outs["1"]), flush=True)
