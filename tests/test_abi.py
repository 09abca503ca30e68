"""The C-ABI shared library loads without a GPU and exports every symbol that
include/sgm_b200.h declares (no compute entry point is called here)."""
import ctypes
import re

import numpy as np
import pytest

from sgmethods_b200 import _lib


def declared_symbols():
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgm_[a-z_0-9]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert lib.sgm_version() >= 100


def test_error_reporting_without_gpu():
    lib = _lib.load()
    out = np.zeros(1, dtype=np.int64)
    bad = np.array([[1, 0]], dtype=np.int64)
    rc = lib.sgm_host_combination_coeffs(_lib.ptr(bad, _lib.c_i64), 1, 2, _lib.ptr(out, _lib.c_i64))
    assert rc == _lib.SGM_ERR_INVALID and "leading entry 0" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    # null plan -> invalid argument, no CUDA call is made
    rc = lib.sgm_eval(None, None, 4, 2, None, 1, 1, 1, None, 1, None, 0, None)
    assert rc == _lib.SGM_ERR_INVALID
    flags = ctypes.c_uint32(0)
    assert lib.sgm_plan_take_flags(None, None, ctypes.byref(flags)) == _lib.SGM_ERR_INVALID
    assert lib.sgm_plan_table_bytes(None) == 0


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under sgmethods_b200/ may import it."""
    import os
    pkg = os.path.dirname(_lib.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_match_rows_semantics():
    rng = np.random.default_rng(0)
    old = rng.normal(size=(50, 3))
    new = np.vstack((old[[7, 3]] + 3e-11, rng.normal(size=(2, 3))))
    assert _lib.match_rows(new, old).tolist() == [7, 3, -1, -1]
    dup = np.vstack((old, old[:1] + 1e-12))
    with pytest.raises(AssertionError):
        _lib.match_rows(old[:1], dup)
    # values chained closer than the tolerance: falls back to the exact all-pairs search
    chain = np.array([[0.0], [6e-11], [1.2e-10], [1.8e-10]])
    assert _lib.match_rows(np.array([[1.8e-10 + 9e-11]]), chain).tolist() == [3]


def test_reference_arm_of_the_bench_is_numpy_scipy_only():
    """`bench.py --impl reference` (the CPU baseline the driver runs) takes its tables from
    bench_data/c2.npz and never loads the product's native library."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                           "--warmup", "0", "--cpu-sample", "32"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    line = json.loads(proc.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["native_library_loaded"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0
