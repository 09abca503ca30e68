"""Grouped hierarchical kernel with / without the locality sort of the points (SGM_DEV_FLAGS bit
4 = 16: no sort, dynamic tiles), A/B in one process, several point distributions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)


def timed(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for name in (("c2", "c5", "c2leja", "c2small") if __name__ == "__main__" else ()):
    it, desc, P, NF, _ = build_workload(name)
    P = 10 ** 6
    f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
    hp = it.hierarchical(); hplan = hp._dev(0)[0]; alpha = hp.surpluses(f)
    for dist_name, x in (("normal", torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)),
                         ("uniform[-2,2]", torch.rand((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 4 - 2),
                         ("clustered", torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 0.1 + 0.3)):
        res = {}
        for flags in ("0", "16", "0", "16"):
            os.environ["SGM_DEV_FLAGS"] = flags
            hplan.reload_knobs()
            res.setdefault(flags, []).append(round(timed(lambda: hplan.eval(x, alpha)), 3))
        print(name, dist_name, res, flush=True)
