"""ctypes binding of libsgm_b200.so (C ABI declared in include/sgm_b200.h).

The library is built in-tree (``sgmethods_b200/csrc/libsgm_b200.so``) by ``make`` in that
directory (nvcc, sm_100a).  If the file is missing it is built on first use; if that is
impossible the import of the native side raises -- there is no Python/CPU fallback for any
entry point.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libsgm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sgm_b200.h")

SGM_OK = 0
SGM_ERR_INVALID, SGM_ERR_CUDA, SGM_ERR_OVERFLOW, SGM_ERR_NODE_TOL, SGM_ERR_WORKSPACE = \
    -1, -2, -3, -4, -5
KIND_LINEAR, KIND_QUADRATIC, KIND_CUBIC, KIND_LAGRANGE = 0, 1, 2, 3
KIND_HIER_LINEAR = 5          # SGM_HIER: one-basis-per-entry hierarchical form
KIND_HIER = 5
HB_HAT, HB_QUAD_MID, HB_LAG_ONE, HB_QUAD_END = 0, 1, 2, 3     # sgm_hier_basis
FLAG_QUADRATIC_RIGHT = 1

c_i32, c_i64, c_f64, c_vp, c_sz = (ctypes.c_int32, ctypes.c_int64, ctypes.c_double,
                                   ctypes.c_void_p, ctypes.c_size_t)
p_i32, p_i64, p_f64 = (ctypes.POINTER(c_i32), ctypes.POINTER(c_i64), ctypes.POINTER(c_f64))


class PlanDesc(ctypes.Structure):
    """struct sgm_plan_desc"""
    _fields_ = [("kind", c_i32), ("N", c_i32), ("T", c_i64), ("M", c_i64),
                ("coeffs", p_f64), ("term_fam", p_i32), ("map_off", p_i64), ("maps", p_i32),
                ("n_fam", c_i32), ("fam_off", p_i64), ("fam_knots", p_f64),
                ("fam_lambda", p_f64), ("fam_newrank", p_i32), ("fam_basis", p_i32),
                ("n_seg", c_i32), ("seg_off", p_i64)]


# name -> (restype, argtypes); must list every symbol of include/sgm_b200.h
SIGNATURES = {
    "sgm_last_error": (ctypes.c_char_p, []),
    "sgm_version": (ctypes.c_int, []),
    "sgm_host_combination_coeffs": (ctypes.c_int, [p_i64, c_i64, c_i32, p_i64]),
    "sgm_grid_build": (ctypes.c_int, [c_i64, c_i32, p_i32, c_i32, p_i32, p_i64, p_f64, c_f64,
                                      ctypes.POINTER(c_vp)]),
    "sgm_grid_build_families": (ctypes.c_int, [c_i64, c_i32, p_i32, c_i32, p_i64, p_f64, c_f64,
                                               ctypes.POINTER(c_vp)]),
    "sgm_hier_apply": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                                      c_vp]),
    "sgm_hier_apply_batches": (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, p_i64,
                                              c_vp, c_i32, c_vp]),
    "sgm_max_rel_deviation": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "sgm_grid_num_nodes": (c_i64, [c_vp]),
    "sgm_grid_map_size": (c_i64, [c_vp]),
    "sgm_grid_nnz": (c_i64, [c_vp]),
    "sgm_grid_copy_maps": (ctypes.c_int, [c_vp, p_i64, p_i32]),
    "sgm_grid_copy_nodes_dense": (ctypes.c_int, [c_vp, p_f64]),
    "sgm_grid_copy_nodes_csr": (ctypes.c_int, [c_vp, p_i64, p_i32, p_f64]),
    "sgm_grid_destroy": (None, [c_vp]),
    "sgm_host_hier_groups": (ctypes.c_int, [c_i64, c_i32, p_i32, c_i32, p_i32, p_i64, c_i32, p_i64,
                                            p_i64, p_i64, p_i32, p_i32, p_i32, p_i32, p_i32, p_i32]),
    "sgm_host_match_rows": (ctypes.c_int, [p_f64, c_i64, p_f64, c_i64, c_i32, c_f64, p_i64]),
    "sgm_plan_create": (ctypes.c_int, [ctypes.POINTER(PlanDesc), ctypes.c_int,
                                       ctypes.POINTER(c_vp)]),
    "sgm_plan_destroy": (None, [c_vp]),
    "sgm_plan_reload_knobs": (ctypes.c_int, [c_vp]),
    "sgm_plan_jit_state": (ctypes.c_int, [c_vp, ctypes.c_char_p, c_sz]),
    "sgm_plan_jit_wait": (ctypes.c_int, [c_vp]),
    "sgm_host_small_kernel_source": (ctypes.c_int, [ctypes.POINTER(PlanDesc), c_i32, c_i32, c_i32,
                                                    ctypes.c_char_p, c_sz, ctypes.POINTER(c_sz)]),
    "sgm_eval_segments": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp,
                                         c_i64, c_i64, c_vp, c_sz, c_vp]),
    "sgm_segment_sumsq": (ctypes.c_int, [c_vp, c_i32, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "sgm_surplus_norms_workspace_bytes": (c_sz, [c_vp, c_i64, c_i64]),
    "sgm_surplus_norms": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp,
                                         c_vp, c_sz, c_vp]),
    "sgm_plan_table_bytes": (c_i64, [c_vp]),
    "sgm_plan_corners_per_point": (c_i64, [c_vp]),
    "sgm_plan_take_flags": (ctypes.c_int, [c_vp, c_vp, ctypes.POINTER(ctypes.c_uint32)]),
    "sgm_eval_workspace_bytes": (c_sz, [c_vp, c_i64, c_i64]),
    "sgm_eval": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp,
                                c_i64, c_vp, c_sz, c_vp]),
    "sgm_eval_signed_sum": (ctypes.c_int, [ctypes.POINTER(c_vp), p_f64, c_i32, c_vp, c_i64,
                                           c_i64, ctypes.POINTER(c_vp), p_i64, p_i64, c_i64,
                                           c_vp, c_i64, c_vp, c_sz, c_vp]),
    "sgm_lc_brownian": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_f64, c_f64,
                                       c_vp, c_i32, c_vp, c_i64, c_vp]),
    "sgm_kl_brownian": (ctypes.c_int, [c_vp, c_i64, c_vp, c_i64, c_i32, c_i64, c_vp, c_vp,
                                       c_vp, c_i64, c_vp]),
}

_lock = threading.Lock()
_lib = None


def build(force=False, verbose=False):
    """Compile libsgm_b200.so in-tree with nvcc for sm_100a (``make`` in csrc/)."""
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    proc = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0 or not os.path.exists(LIB_PATH):
        raise RuntimeError("building libsgm_b200.so failed (nvcc for sm_100a is required):\n"
                           + proc.stderr[-2000:])
    return LIB_PATH


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
            if f.endswith((".cu", ".cuh", ".cpp", ".h"))] + [HEADER_PATH]
    return any(os.path.getmtime(s) > built for s in srcs)


def load():
    """The loaded library with typed entry points (built first if missing or out of date)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            try:
                build()
            except Exception as exc:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        "sgmethods_b200 needs its CUDA library libsgm_b200.so and there is no "
                        "CPU fallback; build it with `make -C sgmethods_b200/csrc`") from exc
                import warnings
                warnings.warn("libsgm_b200.so is older than its sources and could not be rebuilt "
                              f"({exc}); loading the existing library", RuntimeWarning)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return _lib


class SgmError(RuntimeError):
    pass


def last_error():
    msg = load().sgm_last_error()
    return msg.decode() if msg else ""


def check(rc, invalid=ValueError):
    """Turn a negative sgm_status into the matching Python exception."""
    if rc == SGM_OK:
        return
    msg = last_error()
    if rc == SGM_ERR_INVALID:
        raise invalid(msg)
    if rc == SGM_ERR_OVERFLOW:
        raise OverflowError(msg)
    if rc == SGM_ERR_NODE_TOL:
        raise ValueError(msg)
    if rc == SGM_ERR_WORKSPACE:
        raise MemoryError(msg)
    raise SgmError(f"sgm status {rc}: {msg}")


def ptr(arr, ctype):
    """Typed pointer to a C-contiguous numpy array (None for an empty array)."""
    if arr is None or arr.size == 0:
        return ctypes.cast(None, ctypes.POINTER(ctype))
    assert arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(ctypes.POINTER(ctype))


# ---------------------------------------------------------------------------------- host side
def combination_coeffs(mid_set):
    """int64 combination coefficient of every row of a lexsorted multi-index set."""
    mid = np.ascontiguousarray(mid_set, dtype=np.int64)
    if mid.ndim != 2:
        raise ValueError("mid_set must be a 2-D integer array")
    out = np.zeros(mid.shape[0], dtype=np.int64)
    check(load().sgm_host_combination_coeffs(ptr(mid, c_i64), mid.shape[0], mid.shape[1],
                                             ptr(out, c_i64)))
    return out


class HostGrid:
    """Owner of an sgm_grid handle (sparse-grid node numbering)."""

    def __init__(self, num_nodes, fam_m, fam_off, fam_knots, tol=1.e-10, term_fam=None):
        """Either ``num_nodes`` (T, N) with the families keyed by node count ``fam_m``, or
        ``term_fam`` (T, N) naming a family per (term, direction), -1 = single node at 0."""
        lib = load()
        fam_off = np.ascontiguousarray(fam_off, dtype=np.int64)
        fam_knots = np.ascontiguousarray(fam_knots, dtype=np.float64)
        handle = c_vp()
        if term_fam is not None:
            term_fam = np.ascontiguousarray(term_fam, dtype=np.int32)
            self.T, self.N = term_fam.shape
            check(lib.sgm_grid_build_families(self.T, self.N, ptr(term_fam, c_i32),
                                              fam_off.size - 1, ptr(fam_off, c_i64),
                                              ptr(fam_knots, c_f64), tol, ctypes.byref(handle)))
        else:
            num_nodes = np.ascontiguousarray(num_nodes, dtype=np.int32)
            self.T, self.N = num_nodes.shape
            fam_m = np.ascontiguousarray(fam_m, dtype=np.int32)
            check(lib.sgm_grid_build(self.T, self.N, ptr(num_nodes, c_i32), fam_m.size,
                                     ptr(fam_m, c_i32), ptr(fam_off, c_i64),
                                     ptr(fam_knots, c_f64), tol, ctypes.byref(handle)))
        self._h = handle
        self.num_nodes = int(lib.sgm_grid_num_nodes(handle))
        self.map_size = int(lib.sgm_grid_map_size(handle))
        self.nnz = int(lib.sgm_grid_nnz(handle))

    def maps(self):
        map_off = np.zeros(self.T + 1, dtype=np.int64)
        maps = np.zeros(self.map_size, dtype=np.int32)
        check(load().sgm_grid_copy_maps(self._h, ptr(map_off, c_i64), ptr(maps, c_i32)))
        return map_off, maps

    def nodes_dense(self):
        SG = np.zeros((self.num_nodes, self.N), dtype=np.float64)
        if SG.size:
            check(load().sgm_grid_copy_nodes_dense(self._h, ptr(SG, c_f64)))
        return SG

    def nodes_csr(self):
        row_off = np.zeros(self.num_nodes + 1, dtype=np.int64)
        cols = np.zeros(self.nnz, dtype=np.int32)
        vals = np.zeros(self.nnz, dtype=np.float64)
        check(load().sgm_grid_copy_nodes_csr(self._h, ptr(row_off, c_i64), ptr(cols, c_i32),
                                             ptr(vals, c_f64)))
        return row_off, cols, vals

    def close(self):
        if getattr(self, "_h", None):
            load().sgm_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def match_rows(xx, old_xx, tol=1.e-10):
    """For each row of xx the index of the unique row of old_xx within `tol` (1-norm), -1 if
    none; AssertionError if several match (as the reference asserts)."""
    xx = np.ascontiguousarray(xx, dtype=np.float64)
    old_xx = np.ascontiguousarray(old_xx, dtype=np.float64)
    assert xx.ndim == 2 and old_xx.ndim == 2 and xx.shape[1] == old_xx.shape[1]
    match = np.full(xx.shape[0], -1, dtype=np.int64)
    check(load().sgm_host_match_rows(ptr(xx, c_f64), xx.shape[0], ptr(old_xx, c_f64),
                                     old_xx.shape[0], xx.shape[1], tol, ptr(match, c_i64)),
          invalid=AssertionError)
    return match


def hier_groups(term_fam, fam_extent, val_off, n_lists=64):
    """Tables of the grouped hierarchical kernel (sgm_host_hier_groups): groups (n, 12),
    parents (n, 8), steps (m, 2), hmap (slots,) int32 arrays, the root's slot and the
    n_lists + 1 bounds of the warps' group lists."""
    term_fam = np.ascontiguousarray(term_fam, dtype=np.int32)
    fam_extent = np.ascontiguousarray(fam_extent, dtype=np.int32)
    val_off = np.ascontiguousarray(val_off, dtype=np.int64)
    T, N = term_fam.shape
    ng, ns, nv, root = c_i64(0), c_i64(0), c_i64(0), c_i32(-1)
    none32 = ctypes.cast(None, p_i32)
    args = (T, N, ptr(term_fam, c_i32), fam_extent.size, ptr(fam_extent, c_i32), ptr(val_off, c_i64),
            n_lists, ctypes.byref(ng), ctypes.byref(ns), ctypes.byref(nv), ctypes.byref(root))
    check(load().sgm_host_hier_groups(*args, none32, none32, none32, none32, none32))
    groups = np.zeros((max(ng.value, 1), 12), dtype=np.int32)
    parents = np.zeros((max(ng.value, 1), 8), dtype=np.int32)
    steps = np.zeros((max(ns.value, 1), 2), dtype=np.int32)
    hmap = np.zeros(max(nv.value, 1), dtype=np.int32)
    split = np.zeros(n_lists + 1, dtype=np.int32)
    check(load().sgm_host_hier_groups(*args, ptr(groups, c_i32), ptr(parents, c_i32),
                                      ptr(steps, c_i32), ptr(hmap, c_i32), ptr(split, c_i32)))
    return groups[:ng.value], parents[:ng.value], steps[:ns.value], hmap[:nv.value], root.value, split
