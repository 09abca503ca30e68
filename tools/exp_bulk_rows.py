"""Experiment: row segments of the wide hierarchical kernel staged by bulk asynchronous copies
(SGM_WIDE_BULK=1, eval_cols_wide_bulk_kernel) against the L1-cached LDG.128 kernels (one point
per warp: SGM_WIDE_PAIRS=0; two points per warp: default for config 3; four: default for config 4)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import build_workload  # noqa: E402
from tools.exp_hier_ops import timed  # noqa: E402

dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(3)
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name in ("c4", "c3full"):
    if only and only != name:
        continue
    it, desc, _, _, sampler = build_workload(name)
    if name == "c4":
        it = it.interp_seq[-1]
    P, NF = 100000, 1000
    x = torch.randn((P, it.N), generator=gen, device=dev, dtype=torch.float64) if sampler == "normal" else \
        torch.rand((P, it.N), generator=gen, device=dev, dtype=torch.float64) * 2 - 1
    f = torch.randn((it.num_nodes, NF), generator=gen, device=dev, dtype=torch.float64)
    hp = it.hierarchical()
    plan = hp._dev(0)[0]
    alpha = hp.surpluses(f)
    res, outs = {}, {}
    for label, env in (("default", {}), ("one point per warp (LDG)", {"SGM_WIDE_PAIRS": "0"}),
                       ("one point per warp (bulk copies)", {"SGM_WIDE_BULK": "1"})):
        for k in ("SGM_WIDE_PAIRS", "SGM_WIDE_BULK"):
            os.environ.pop(k, None)
        os.environ.update(env)
        plan.reload_knobs()
        out = torch.empty((P, NF), dtype=torch.float64, device=dev)
        res[label] = timed(lambda: plan.eval(x, alpha, out))[0]
        outs[label] = out
    same = torch.equal(outs["one point per warp (LDG)"], outs["one point per warp (bulk copies)"])
    dev_rel = float((outs["default"] - outs["one point per warp (bulk copies)"]).abs().max() / outs["default"].abs().max())
    print(f"{name} hierarchical P={P} NF={NF}: " + ", ".join(f"{k} {v:.2f} ms" for k, v in res.items()) +
          f"; bulk == LDG bit for bit: {same}; vs default rel dev {dev_rel:.1e}", flush=True)
