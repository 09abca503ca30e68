"""End-to-end (pinned host -> device -> host) and device-resident step of config 2, A/B in one
process: validation on a side stream or on the main stream; dynamic or static tiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from sgmethods_b200 import sparse_grid_interpolant as sgi
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
it, desc, P, NF, _ = build_workload("c2")
x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
xh, fh = x.cpu().pin_memory(), f.cpu().pin_memory()
oh = torch.empty((P, 1), dtype=torch.float64).pin_memory()


def timed(fn, reps=9):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def e2e():
    oh.copy_(it.interpolate(xh, fh).reshape(P, 1), non_blocking=True)


for rnd in range(3):
    print(f"device step {timed(lambda: it.interpolate(x, f)):.3f} ms, e2e step {timed(e2e):.3f} ms", flush=True)
