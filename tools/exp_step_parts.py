"""Where a method='auto' step of config 2 goes (CUDA events around the pieces, L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator(device=dev).manual_seed(1)
it, desc, P, NF, _ = build_workload("c2")
x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
hp = it.hierarchical()
hplan = hp._dev(0)[0]


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


alpha = hp.surpluses(f)
print(f"whole step (interpolate, auto): {timed(lambda: it.interpolate(x, f)):.3f} ms")
print(f"surpluses (index_select + hierarchisation): {timed(lambda: hp.surpluses(f)):.3f} ms")
print(f"plan.eval (pregather + sort + grouped kernel): {timed(lambda: hplan.eval(x, alpha)):.3f} ms")
os.environ["SGM_NO_SORT"] = "1"; hplan.reload_knobs()
print(f"plan.eval without the locality sort: {timed(lambda: hplan.eval(x, alpha)):.3f} ms")
del os.environ["SGM_NO_SORT"]; hplan.reload_knobs()
xs = x[::3906][:256].contiguous()
print(f"validation pieces: exact 256 {timed(lambda: it.device_plan(0).eval(xs, f)):.3f} ms, hier 256 {timed(lambda: hplan.eval(xs, alpha)):.3f} ms")
