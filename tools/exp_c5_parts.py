"""Where a method='auto' call of config 5 (10^6 points) goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from tools.exp_nosort import timed, dev, gen
it, desc, P, NF, _ = build_workload("c5")
x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)
f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
hp = it.hierarchical(); hplan = hp._dev(0)[0]
alpha = hp.surpluses(f)
print(f"interpolate auto: {timed(lambda: it.interpolate(x, f), 3):.2f} ms")
print(f"interpolate hierarchical (no validation): {timed(lambda: it.interpolate(x, f, method='hierarchical'), 3):.2f} ms")
print(f"surpluses: {timed(lambda: hp.surpluses(f), 3):.2f} ms")
print(f"plan.eval: {timed(lambda: hplan.eval(x, alpha), 3):.2f} ms")
xs = x[::3906][:256].contiguous()
print(f"exact 256: {timed(lambda: it.device_plan(0).eval(xs, f), 3):.2f} ms; hier 256: {timed(lambda: hplan.eval(xs, alpha), 3):.2f} ms")
t0 = time.perf_counter(); it.interpolate(x, f); torch.cuda.synchronize(); print("wall one call", round(1e3 * (time.perf_counter() - t0), 2), "ms")
