"""Exact kernel on small batches (the validation sample of method='auto': 256 points) and on
config 5 / config 2 at full batches, A/B of SGM_DEV_FLAGS bit 2 (4 = no record prefetch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)


def timed(fn, reps=9):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for name, sizes in (("c2", (256, 2048, 10 ** 6)), ("c5", (256, 10 ** 5)), ("c2leja", (256, 10 ** 5))):
    it, desc, _, NF, _ = build_workload(name)
    f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
    plan = it.device_plan(0)
    for P in sizes:
        x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
        res = {}
        for flags in (0, 4, 0, 4):
            os.environ["SGM_DEV_FLAGS"] = str(flags)
            plan.reload_knobs()
            res.setdefault(flags, []).append(round(timed(lambda: plan.eval(x, f), 5 if P > 10000 else 9), 4))
        print(name, P, res, flush=True)
