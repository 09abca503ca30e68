"""Experiment: how much does point locality (sorting by the sign pattern of the coordinates)
help the gather-bound scalar kernel?  Dev tool, not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_workload

it, desc, P, NF, sampler = build_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0)

def timeit(xx, label):
    out = torch.empty((P, 1), dtype=torch.float64, device=dev)
    for _ in range(3): plan.eval(xx, f, out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): plan.eval(xx, f, out)
    b.record(); torch.cuda.synchronize()
    print(f"{label}: {a.elapsed_time(b)/5:.2f} ms")
    return out

o0 = timeit(x, "random order")
for nb in (8, 16):
    bits = (x[:, :nb] >= 0).to(torch.int64)
    key = (bits * (2 ** torch.arange(nb - 1, -1, -1, device=dev))).sum(1)
    perm = torch.argsort(key)
    xs = x[perm].contiguous()
    o1 = timeit(xs, f"sorted by sign pattern of first {nb} dims")
    assert torch.equal(o1, o0[perm])
# finer key: 2 bits per dim (quartile cell of the 7-knot level) for the first 16 dims
g7 = torch.from_numpy(it._fam_knots_list[1]).to(dev)
cell = torch.bucketize(x, g7[1:-1].contiguous())          # 0..5
key = torch.zeros(P, dtype=torch.int64, device=dev)
for d in range(16):
    key = key * 8 + cell[:, d]
perm = torch.argsort(key)
o2 = timeit(x[perm].contiguous(), "sorted by 7-knot cell of all dims (lexicographic)")
assert torch.equal(o2, o0[perm])
# floor without gather divergence: all points identical -> every gather is a broadcast
xsame = x[:1].expand(P, -1).contiguous()
timeit(xsame, "all points identical (broadcast gathers: issue/FP64/smem floor)")
