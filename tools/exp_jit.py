"""example/1 shape at 10^7 points: generic kernel A0 against the plan-specialised kernel
(built synchronously here), several launch shapes via SGM_JIT_SHAPE."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SGM_JIT_SYNC"] = "1"
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
dev = torch.device("cuda"); P = int(sys.argv[1]) if len(sys.argv) > 1 else 10 ** 7
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
out = torch.empty((P, 1), dtype=torch.float64, device=dev)
def timeit(plan):
    for _ in range(3): plan.eval(x, f, out)
    torch.cuda.synchronize(); ts = []
    for i in range(9):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.eval(x, f, out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))
os.environ["SGM_NO_JIT"] = "1"
plan = it.device_plan(0); ms0 = timeit(plan); ref = out.clone()
print(f"generic A0        {ms0:.4f} ms {88 * P / ms0 / 1e6:7.0f} GB/s ({88 * P / ms0 / 1e6 / 6563.2:.3f})")
del os.environ["SGM_NO_JIT"]; plan.reload_knobs()
ms = timeit(plan); print(plan.jit_state())
print(f"specialised       {ms:.4f} ms {88 * P / ms / 1e6:7.0f} GB/s ({88 * P / ms / 1e6 / 6563.2:.3f})  bit-identical: {torch.equal(out, ref)}")
