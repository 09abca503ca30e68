"""Exact path (locality sort with warp-aggregated atomics) on clustered points, config 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from tools.exp_nosort import timed, dev, gen
it, desc, P, NF, _ = build_workload("c2")
f = torch.randn((it.num_nodes, 1), generator=gen, device=dev, dtype=torch.float64)
plan = it.device_plan(0)
for name, x in (("normal", torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64)),
                ("clustered", torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 0.1 + 0.3)):
    a = plan.eval(x, f)
    os.environ["SGM_NO_SORT"] = "1"; plan.reload_knobs()
    b = plan.eval(x, f); t_ns = timed(lambda: plan.eval(x, f), 3)
    del os.environ["SGM_NO_SORT"]; plan.reload_knobs()
    print(name, "exact with sort", round(timed(lambda: plan.eval(x, f), 3), 3), "ms, without", round(t_ns, 3), "ms, identical", torch.equal(a, b), flush=True)
