"""Experiment: prefix kernel (consecutive terms per warp, reused weight prefixes) against the
split kernel on a bench workload.  Dev tool, not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload

it, desc, P, NF, sampler = build_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
if len(sys.argv) > 2:
    P = int(sys.argv[2])
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(label):
    out = torch.empty((P, 1), dtype=torch.float64, device=dev)
    for _ in range(3): plan.eval(x, f, out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.eval(x, f, out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{label}: median {np.median(ts):.2f} ms  min {min(ts):.2f} ms", flush=True)
    return out

os.environ["SGM_FORCE_KERNEL"] = "split"
o0 = timeit("split kernel")
os.environ["SGM_FORCE_KERNEL"] = "prefix"
for cfg in sys.argv[3].split(",") if len(sys.argv) > 3 else ("8x4", "8x8u", "8x8"):
    unb = cfg.endswith("u")
    warps, run = cfg.rstrip("u").split("x")
    os.environ.pop("SGM_PREFIX_UNBALANCED", None)
    if unb:
        os.environ["SGM_PREFIX_UNBALANCED"] = "1"
    os.environ["SGM_PREFIX_RUN"] = run
    o1 = timeit(f"prefix kernel, {warps} warps, run {run}{' (uniform runs)' if unb else ''}")
    print("   bit-identical to split:", bool(torch.equal(o0, o1)))
