"""example/1 as shipped at 10^7 points (the HBM-bound regime of the interpolant): kernel A0
against kernel A, with the algorithmic HBM rate."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200._device import reload_knobs
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
S = mis.aniso_smolyak_mid_set(5, 10, 2 ** (np.linspace(0, 9, 10)))
it = SGInterpolant(S, nodes_1d.unbounded_nodes_nested, lambda n: 2 ** (n + 1) - 1, verbose=False)
dev = torch.device("cuda"); P = 10 ** 7
gen = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen)
f = torch.randn((it.num_nodes, 1), dtype=torch.float64, device=dev, generator=gen)
plan = it.device_plan(0); out = torch.empty((P, 1), dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
res = {}
for label, env in (("A0 small-plan kernel", None), ("A thread-per-point kernel", "1")):
    if env: os.environ["SGM_NO_SMALL"] = env
    else: os.environ.pop("SGM_NO_SMALL", None)
    reload_knobs()
    for _ in range(3): plan.eval(x, f, out)
    torch.cuda.synchronize()
    ts = []
    for i in range(7):
        flush.fill_(float(i))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.eval(x, f, out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts)); res[label] = out.clone()
    print(f"{label:28s} {ms:7.3f} ms  {88 * P / ms / 1e6:7.0f} GB/s algorithmic ({88 * P / ms / 1e6 / 6563.2:.2f} of HBM peak)")
print("bit-identical:", torch.equal(*res.values()))
