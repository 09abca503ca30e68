"""Hierarchical-surplus scalar kernels on config 2: H1 (one record per multi-index) against the
launch shapes of the grouped kernel H2, with the deviation from the exact combination path.
Usage: python tools/exp_h2.py [workload] [points]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import build_workload
from sgmethods_b200._device import reload_knobs

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
it, desc, P, NF, sampler = build_workload(name)
if len(sys.argv) > 2:
    P = int(sys.argv[2])
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
exact = it.device_plan(0).eval(x, f)
hp = it.hierarchical()
plan, _, _ = hp._dev(0)
alpha = hp.surpluses(f)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(label):
    reload_knobs()
    out = plan.eval(x, alpha)
    torch.cuda.synchronize()
    ts = []
    for i in range(7):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.eval(x, alpha, out); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    dev_rel = float((out - exact).abs().max() / exact.abs().max())
    print(f"{label:28s} {np.median(ts):8.3f} ms  (min {min(ts):.3f})  max rel dev from exact {dev_rel:.2e}")


os.environ["SGM_NO_H2"] = "1"
timed("H1 (per multi-index)")
del os.environ["SGM_NO_H2"]
for shape, label in ((0, "16w x2"), (1, "16w x2 contiguous tiles"), (2, "8w x4"),
                     (3, "8w x3"), (4, "16w x1"), (5, "16w x2 unroll4")):
    os.environ["SGM_H2_SHAPE"] = str(shape)
    timed("H2 " + label)
del os.environ["SGM_H2_SHAPE"]
reload_knobs()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record(); alpha = hp.surpluses(f); b.record(); torch.cuda.synchronize()
print(f"hierarchisation of the nodal values: {a.elapsed_time(b):.3f} ms ({len(hp.batches)} batches)")
