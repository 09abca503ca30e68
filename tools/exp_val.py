import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from bench import build_workload
from tools.exp_hier_ops import timed
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(1)
ml, desc, P, NF, _ = build_workload("c4")
it = ml.interp_seq[-1]
P = 10 ** 5
x = torch.rand((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 2 - 1
f = torch.randn((it.num_nodes, NF), generator=gen, device=dev, dtype=torch.float64)
hp = it.hierarchical()
plan = it.device_plan(0)
hplan = hp._dev(0)[0]
xs = x[::390][:256].contiguous()
print("exact 256:", timed(lambda: plan.eval(xs, f))[0])
al = hp.surpluses(f)
print("surpluses:", timed(lambda: hp.surpluses(f))[0], len(hp.batches))
print("hier 256:", timed(lambda: hplan.eval(xs, al))[0])
print("hier 1e5:", timed(lambda: hplan.eval(x, al))[0])
for n in (1024, 4096, 16384):
    xx = x[:n].contiguous()
    print("exact", n, timed(lambda: plan.eval(xx, f))[0], "hier", timed(lambda: hplan.eval(xx, al))[0])
