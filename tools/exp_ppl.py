"""Experiment: prefix kernel with 1 or 2 points per lane on a bench workload.  Dev tool, not
part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload
from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(min(ts))


for wl in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["c2"]):
    it, desc, P, NF, sampler = build_workload(wl)
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
    f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
    plan = it.device_plan(0)
    outs = {}
    os.environ["SGM_FORCE_KERNEL"] = "split"
    ref = torch.empty((P, 1), dtype=torch.float64, device=dev)
    med, mn = timed(lambda: plan.eval(x, f, ref))
    print(f"{wl} split kernel: median {med:.2f} ms  min {mn:.2f} ms", flush=True)
    for ppl in ("1", "2"):
        os.environ["SGM_FORCE_KERNEL"] = "prefix"
        os.environ["SGM_PREFIX_PPL"] = ppl
        out = torch.empty((P, 1), dtype=torch.float64, device=dev)
        med, mn = timed(lambda: plan.eval(x, f, out))
        print(f"{wl} prefix kernel, {ppl} point(s) per lane: median {med:.2f} ms  "
              f"min {mn:.2f} ms  bit-identical to split: {bool(torch.equal(ref, out))}", flush=True)
    for k in ("SGM_FORCE_KERNEL", "SGM_PREFIX_PPL"):
        os.environ.pop(k)
