"""GPU fuzz of the Levy-Ciesielski kernel (bulk-copy and cp.async flavours: contiguous rows /
strided row views, odd and even d, ragged row counts) against the CPU oracle, bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import brownian_oracle
from sgmethods_b200.parametric_expansions_brownian_motion import lc_brownian_device

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
t_end = time.time() + budget
dev = torch.device("cuda")
n = bad = 0
while time.time() < t_end:
    rng = np.random.default_rng(300 + n)
    d = int(rng.choice([1, 2, 3, 7, 8, 10, 16, 31, 32, 33, 64, 65, 100, 128, 200, 256, 300]))
    nt = int(rng.choice([1, 31, 32, 33, 100, 511, 512, 513, 1000]))
    P = int(rng.choice([1, 5, 31, 32, 33, 64, 97, 257, 1000, 4097]))
    T = float(rng.choice([1.0, 2.5]))
    pad = int(rng.choice([0, 0, 1, 2, 3]))
    ybig = rng.normal(size=(P, d + pad))
    tt = np.sort(rng.uniform(0, T, nt))
    try:
        y_d = torch.from_numpy(ybig).to(dev)[:, :d]          # strided rows when pad > 0
        got = lc_brownian_device(torch.from_numpy(tt).to(dev), y_d, T).cpu().numpy()
        ref = np.stack([brownian_oracle.lc_brownian(tt, ybig[p, :d], T) for p in range(min(P, 40))])
        assert np.array_equal(got[:ref.shape[0]], ref), float(np.max(np.abs(got[:ref.shape[0]] - ref)))
        if P > 40:                                           # the other rows: same kernel on a contiguous copy
            again = lc_brownian_device(torch.from_numpy(tt).to(dev), torch.from_numpy(np.ascontiguousarray(ybig[:, :d])).to(dev), T).cpu().numpy()
            assert np.array_equal(got, again), "strided vs contiguous rows differ"
    except Exception as exc:  # noqa: BLE001
        bad += 1
        print(f"CASE {n} d={d} nt={nt} P={P} pad={pad}: {type(exc).__name__}: {exc}", flush=True)
    n += 1
print(f"brownian fuzz: {n} cases, {bad} failures", flush=True)
