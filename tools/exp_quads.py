"""Experiment: wide column kernel with two (256-column tiles) or four (128-column tiles) points
per warp; exact kinds must agree bit for bit.  Dev tool, not part of the product."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgmethods_b200 import multi_index_sets as mis, nodes_1d
from sgmethods_b200.sparse_grid_interpolant import SGInterpolant
from sgmethods_b200.tp_inteprolants import TPPwQuadraticInterpolator
from sgmethods_b200.utils import compute_level_lc
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


a = np.array([compute_level_lc(n)[0] + 1 for n in range(64)], dtype=float)
c3 = SGInterpolant(mis.aniso_smolyak_mid_set(10, 64, a), lambda n: nodes_1d.optimal_gaussian_nodes(n, p=2),
                   lambda n: 2 ** (n + 1) - 1, verbose=False)
c4 = SGInterpolant(mis.smolyak_mid_set(3, 8), nodes_1d.cc_nodes, lambda n: np.where(n == 0, 1, 2 ** n + 1),
                   tp_interpolant=TPPwQuadraticInterpolator, verbose=False)
cases = [("c3 w=10 exact", c3, c3.device_plan(0), None, "normal"),
         ("c4 lvl3 quadratic", c4, c4.device_plan(0), None, "uniform")]
hp = c3.hierarchical()
cases.append(("c3 w=10 hierarchical", c3, hp._dev(0)[0], hp, "normal"))
hp4 = c4.hierarchical()
cases.append(("c4 lvl3 hierarchical", c4, hp4._dev(0)[0], hp4, "uniform"))
if len(sys.argv) > 1:
    cases = [c for c in cases if sys.argv[1] in c[0]]
for name, it, plan, hier, sampler in cases:
    gen = torch.Generator(device=dev).manual_seed(0)
    for NF in (1000, 4096):
        f = torch.randn((it.num_nodes, NF), dtype=torch.float64, device=dev, generator=gen)
        fe = hier.surpluses(f) if hier is not None else f
        for P in (16384, 100000) + ((1000000,) if name.startswith('c4') and NF == 1000 else ()):
            x = torch.randn((P, it.N), dtype=torch.float64, device=dev, generator=gen) if sampler == "normal" \
                else torch.rand((P, it.N), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
            res, outs = {}, {}
            for pts in ("2", "4") + (("8",) if hier is not None else ()):
                os.environ["SGM_WIDE_PTS"] = pts
                plan.reload_knobs()
                out = torch.empty((P, NF), dtype=torch.float64, device=dev)
                res[pts] = timed(lambda: plan.eval(x, fe, out))
                outs[pts] = out
            same = torch.equal(outs["2"], outs["4"])
            dev_rel = float((outs["2"] - outs["4"]).abs().max() / outs["2"].abs().max())
            extra = ""
            if "8" in res:
                extra = (f"; 4 pts/warp x 256 columns {res['8']:.2f} ms (x{res['2'] / res['8']:.2f}) "
                         f"identical={torch.equal(outs['2'], outs['8'])}")
            print(f"{name} P={P} NF={NF}: 2 pts/warp {res['2']:.2f} ms, 4 pts/warp {res['4']:.2f} ms "
                  f"(x{res['2'] / res['4']:.2f}) identical={same} rel dev {dev_rel:.1e}{extra}", flush=True)
            del outs
