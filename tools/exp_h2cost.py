"""Scan of the cost model that deals the groups of the hierarchical kernel to the warps
(SGM_H2_COST = "group,parent entry,step,leaf", read when a plan is created)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import build_workload
from sgmethods_b200._device import DevicePlan
from sgmethods_b200 import _lib

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
it, desc, P, NF, sampler = build_workload(name)
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
hp = it.hierarchical()
alpha = hp.surpluses(f)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
for cost in ("800,60,250,10", "8,3,5,2.5", "400,60,250,10", "1600,60,250,10", "800,60,250,40", "800,60,150,40",
             "800,60,400,10", "800,200,250,10", "1200,100,200,30", "0,0,1,0", "0,0,0,1"):
    os.environ["SGM_H2_COST"] = cost
    plan = DevicePlan(_lib.KIND_HIER_LINEAR, it.N, hp.M, np.ones(hp._term_fam.shape[0]), hp._term_fam,
                      hp._map_off, hp._maps_f, hp._fam_off, hp._fam_knots, device=0,
                      fam_newrank=hp._fam_newrank)
    out = plan.eval(x, alpha)
    torch.cuda.synchronize()
    ts = []
    for i in range(5):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.eval(x, alpha, out); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"cost {cost:22s} {np.median(ts):8.3f} ms")
    plan.close()
