"""CPU oracle for the sparse-grid interpolant hot path -- TEST INFRASTRUCTURE ONLY.

This package is a plain numpy restatement of the reference algorithm
(andreascaglioni/SGMethods, read-only at /root/reference in the build container) for the
path `SGInterpolant.__init__` tables -> `SGInterpolant.interpolate`, the four tensor-product
operators, the Levy-Ciesielski expansion and the (intended) multilevel / error-indicator
drivers.  Every function cites the reference file:line it follows.

Rules:
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl
    reference` legs may import this package, and only as the checker / the timed CPU
    baseline -- never as part of the product path (`sgmethods_b200/` must not import it);
  * parity status: PINNED.  `tests/test_oracle_golden.py` checks every function here
    against fixtures in `tests/golden/*.npz` that were produced by running the unmodified
    reference (`tests/golden/make_golden.py`), against the golden integer arrays of the
    reference's own tests, and (pw-linear) against scipy's RegularGridInterpolator, which
    is the third-party code the reference calls for that operator
    (sgmethods/tp_inteprolants.py:49-53; pinned scipy==1.10.1 in requirements.txt:24,
    1.18.1 installed here).
"""
