"""Cache the tables of the default bench workload for the `--impl reference` arm.

    python tools/make_ref_tables.py [workload ...]        (default: c2)

`bench.py --impl reference` times the reference's CPU path (oracle/, the numpy/scipy port of
its term loop).  The reference's own constructor is O(M^2) in the number of nodes -- ten
minutes for config 2 -- so the arm takes the per-term tables (coefficients, active
directions, node counts, node maps, knots) from this file instead: bench_data/<workload>.npz,
written HERE by this repo's table builder, whose output is pinned bit for bit to the
reference constructor on every fixture (tests/test_tables.py).  With the cache the arm is
numpy/scipy only: it never loads libsgm_b200.so.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from bench import build_workload  # noqa: E402


def main(names):
    os.makedirs(os.path.join(ROOT, "bench_data"), exist_ok=True)
    for name in names:
        it, desc, P, NF, sampler = build_workload(name)
        T = len(it._coeffs)
        counts = it._num_nodes
        dims_off = np.concatenate(([0], np.cumsum((counts > 1).sum(axis=1)))).astype(np.int64)
        tt, dd = np.nonzero(counts > 1)
        ms = np.asarray(it._fam_m, dtype=np.int64)
        path = os.path.join(ROOT, "bench_data", name + ".npz")
        np.savez_compressed(
            path, N=np.int64(it.N), num_nodes=np.int64(it.num_nodes), kind=np.int64(it._kind),
            NF=np.int64(NF), P=np.int64(P), sampler=np.array(sampler), desc=np.array(desc),
            coeffs=np.asarray(it._coeffs, dtype=np.int64), dims_off=dims_off,
            dims=dd.astype(np.int32), shapes=counts[tt, dd].astype(np.int32),
            map_off=np.asarray(it._map_off, dtype=np.int64), maps=np.asarray(it._maps, dtype=np.int32),
            fam_m=ms, fam_off=np.asarray(it._fam_off, dtype=np.int64),
            fam_knots=np.asarray(it._fam_knots, dtype=np.float64),
            knots_one=np.atleast_1d(np.asarray(it.knots(np.int64(1)), dtype=np.float64)))
        print(f"{path}: {T} terms, {it.num_nodes} nodes, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main(sys.argv[1:] or ["c2"])
