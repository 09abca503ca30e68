"""method='auto' with the validation sample on a side stream or on the main stream, config 5 /
config 3 / config 2-Leja shapes, A/B in one process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from sgmethods_b200 import sparse_grid_interpolant as sgi
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)


def timed(fn, reps=5):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for name, P, NF in (("c5", 10 ** 6, 1), ("c3full", 10 ** 5, 1000), ("c2leja", 10 ** 6, 1)):
    it, desc, _, _, sampler = build_workload(name)
    x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
    f = torch.randn((it.num_nodes, NF), generator=gen, device=dev, dtype=torch.float64)
    for rnd in range(2):
        for side in (True, False):
            sgi.AUTO_SIDE_STREAM = side
            print(f"{name}: side stream {side}: {timed(lambda: it.interpolate(x, f)):.3f} ms "
                  f"({it.auto_validation['enabled']})", flush=True)
