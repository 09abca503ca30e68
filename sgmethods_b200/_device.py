"""Device-side plumbing: torch is used for device memory, streams and device selection only;
all arithmetic runs in the kernels of libsgm_b200.so.  There is no CPU fallback: every
function here raises if CUDA is unavailable."""
import ctypes
import weakref

import numpy as np

from . import _lib


def torch_cuda():
    """torch, after checking that a CUDA device is usable."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("sgmethods_b200 evaluates on a CUDA device (B200, sm_100a) and has "
                           "no CPU fallback: torch.cuda.is_available() is False")
    return torch


def lagrange_lambdas(g):
    """lambda_i = 1 / prod_{j != i, ascending} (g_i - g_j) with the reference's operation
    order (tp_inteprolants.py:311-320); host side because it is part of operator set-up."""
    g = np.asarray(g, dtype=np.float64)
    lam = np.ones(g.shape[0])
    for i in range(g.shape[0]):
        keep = np.ones(g.shape[0], dtype=bool)
        keep[i] = False
        lam[i] = np.prod(g[i] - g, where=keep)
    return 1 / lam


_LIVE_PLANS = weakref.WeakSet()


def reload_knobs():
    """Re-read the SGM_* development switches for every live plan (they are read once, at plan
    creation).  For tests and the experiment scripts under tools/."""
    for plan in list(_LIVE_PLANS):
        if getattr(plan, "_h", None):
            _lib.check(_lib.load().sgm_plan_reload_knobs(plan._h))


def make_plan_desc(kind, N, M, coeffs, term_fam, map_off, maps, fam_off, fam_knots,
                   fam_newrank=None, seg_off=None, fam_basis=None):
    """``sgm_plan_desc`` for the packed tables of an interpolant, plus the arrays it points into
    (keep them alive as long as the descriptor is used)."""
    kind, N = int(kind), int(N)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    term_fam = np.ascontiguousarray(term_fam, dtype=np.int32).reshape(-1, N)
    map_off = np.ascontiguousarray(map_off, dtype=np.int64)
    maps = np.ascontiguousarray(maps, dtype=np.int32)
    fam_off = np.ascontiguousarray(fam_off, dtype=np.int64)
    n_fam = fam_off.size - 1
    fam_knots = np.ascontiguousarray(fam_knots, dtype=np.float64)
    lam = None
    if kind == _lib.KIND_LAGRANGE:
        lam = np.concatenate([lagrange_lambdas(fam_knots[fam_off[i]:fam_off[i + 1]])
                              for i in range(n_fam)]) if n_fam else np.zeros(0)
    newrank = None if fam_newrank is None else np.ascontiguousarray(fam_newrank, dtype=np.int32)
    basis = None if fam_basis is None else np.ascontiguousarray(fam_basis, dtype=np.int32)
    seg = None if seg_off is None else np.ascontiguousarray(seg_off, dtype=np.int64)
    desc = _lib.PlanDesc(
        kind=kind, N=N, T=coeffs.shape[0], M=int(M),
        coeffs=_lib.ptr(coeffs, _lib.c_f64), term_fam=_lib.ptr(term_fam, _lib.c_i32),
        map_off=_lib.ptr(map_off, _lib.c_i64), maps=_lib.ptr(maps, _lib.c_i32),
        n_fam=n_fam, fam_off=_lib.ptr(fam_off, _lib.c_i64),
        fam_knots=_lib.ptr(fam_knots, _lib.c_f64),
        fam_lambda=_lib.ptr(lam, _lib.c_f64),
        fam_newrank=_lib.ptr(newrank, _lib.c_i32), fam_basis=_lib.ptr(basis, _lib.c_i32),
        n_seg=0 if seg is None else seg.size - 1, seg_off=_lib.ptr(seg, _lib.c_i64))
    return desc, (coeffs, term_fam, map_off, maps, fam_off, fam_knots, lam, newrank, seg, basis)


def small_kernel_source(interp, threads=256, points_per_thread=1, compile_cc=0):
    """CUDA source of the plan-specialised kernel that ``sgm_eval`` builds for few-term pw-linear
    interpolants evaluated at many points (csrc/sgm_jit.cpp); host only.  ``compile_cc=100``
    also compiles it with NVRTC for sm_100a.  ValueError when the plan does not qualify."""
    desc, keep = make_plan_desc(interp._kind, interp.N, interp.num_nodes, interp._coeffs.astype(np.float64),
                                interp._term_fam, interp._map_off, interp._maps, interp._fam_off,
                                interp._fam_knots)
    lib = _lib.load()
    need = _lib.c_sz(0)
    _lib.check(lib.sgm_host_small_kernel_source(ctypes.byref(desc), threads, points_per_thread, 0, None, 0,
                                                ctypes.byref(need)))
    buf = ctypes.create_string_buffer(need.value)
    _lib.check(lib.sgm_host_small_kernel_source(ctypes.byref(desc), threads, points_per_thread, compile_cc,
                                                buf, need.value, None))
    del keep
    return buf.value.decode()


class DevicePlan:
    """Packed tables of one interpolant resident on one GPU (sgm_plan)."""

    def __init__(self, kind, N, M, coeffs, term_fam, map_off, maps, fam_off, fam_knots,
                 device=None, fam_newrank=None, seg_off=None, fam_basis=None):
        """term_fam: (T, N) knot-family index per direction, -1 where the term has one node;
        fam_off / fam_knots: CSR list of the families' ascending knot vectors; seg_off: term
        offsets of the segments of a segmented plan (``eval_segments`` / ``surplus_norms``)."""
        torch = torch_cuda()
        lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.kind, self.N, self.M = int(kind), int(N), int(M)
        desc, self._desc_arrays = make_plan_desc(kind, N, M, coeffs, term_fam, map_off, maps, fam_off,
                                                 fam_knots, fam_newrank, seg_off, fam_basis)
        self.T, self.n_seg = int(desc.T), int(desc.n_seg)
        handle = _lib.c_vp()
        _lib.check(lib.sgm_plan_create(ctypes.byref(desc), self.device, ctypes.byref(handle)))
        self._h = handle
        self.table_bytes = int(lib.sgm_plan_table_bytes(handle))
        self.corners_per_point = int(lib.sgm_plan_corners_per_point(handle))
        self._workspaces = {}
        # measurement hook (bench.py): a list to which every eval appends (P, NF, start, end)
        # CUDA events recorded around the sgm_eval launches on the launching stream
        self.timing = None
        _LIVE_PLANS.add(self)

    # -- evaluation -------------------------------------------------------------------------
    def workspace(self, P, NF, need=None):
        """Scratch for one launch; one buffer per CUDA stream, so that launches of the same
        plan on different streams never share scratch."""
        torch = torch_cuda()
        if need is None:
            need = int(_lib.load().sgm_eval_workspace_bytes(self._h, P, NF))
        dev = torch.device("cuda", self.device)
        key = torch.cuda.current_stream(dev).cuda_stream
        ws = self._workspaces.get(key)
        if ws is None or ws.numel() < need:
            ws = self._workspaces[key] = torch.empty(need, dtype=torch.uint8, device=dev)
        return ws

    def eval(self, x, f, out=None):
        """x (P, >=N), f (M, NF), both fp64 CUDA tensors on this plan's device -> (P, NF)."""
        torch = torch_cuda()
        assert x.is_cuda and f.is_cuda and x.dtype == torch.float64 and f.dtype == torch.float64
        assert x.dim() == 2 and f.dim() == 2
        P, NF = x.shape[0], f.shape[1]
        if out is None:
            out = torch.empty((P, NF), dtype=torch.float64, device=x.device)
        if P == 0 or NF == 0:
            return out
        assert (x.shape[1] == 1 or x.stride(1) == 1) and (NF == 1 or f.stride(1) == 1)
        ws = self.workspace(P, NF)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        if self.timing is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(torch.cuda.current_stream(x.device))
        _lib.check(_lib.load().sgm_eval(
            self._h, x.data_ptr(), P, x.stride(0) if P > 1 else max(x.shape[1], 1),
            f.data_ptr(), f.shape[0], NF, f.stride(0) if f.shape[0] > 1 else max(NF, 1),
            out.data_ptr(), out.stride(0) if P > 1 else max(NF, 1), ws.data_ptr(), ws.numel(),
            stream))
        if self.timing is not None:
            t1.record(torch.cuda.current_stream(x.device))
            self.timing.append((P, NF, t0, t1))
        return out

    def _check_inputs(self, x, f):
        torch = torch_cuda()
        assert x.is_cuda and f.is_cuda and x.dtype == torch.float64 and f.dtype == torch.float64
        assert x.dim() == 2 and f.dim() == 2
        assert (x.shape[1] == 1 or x.stride(1) == 1) and (f.shape[1] == 1 or f.stride(1) == 1)

    def eval_segments(self, x, f, out=None):
        """Segmented plan: (n_seg, P, NF) tensor, out[s] = sum over the terms of segment s
        (sgm_eval_segments, one launch for all segments)."""
        torch = torch_cuda()
        self._check_inputs(x, f)
        P, NF = x.shape[0], f.shape[1]
        if out is None:
            out = torch.empty((self.n_seg, P, NF), dtype=torch.float64, device=x.device)
        if P == 0 or self.n_seg == 0:
            return out
        ws = self.workspace(P, NF)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().sgm_eval_segments(
            self._h, x.data_ptr(), P, x.stride(0) if P > 1 else max(x.shape[1], 1),
            f.data_ptr(), f.shape[0], NF, f.stride(0) if f.shape[0] > 1 else max(NF, 1),
            out.data_ptr(), NF, P * NF, ws.data_ptr(), ws.numel(), stream))
        return out

    def surplus_sumsq(self, x, f):
        """Segmented plan: (n_seg,) tensor of sum_{p, c} out[s, p, c]^2, the surpluses never
        leaving the GPU workspace (sgm_surplus_norms)."""
        torch = torch_cuda()
        self._check_inputs(x, f)
        P, NF = x.shape[0], f.shape[1]
        sumsq = torch.zeros(self.n_seg, dtype=torch.float64, device=x.device)
        if P == 0 or self.n_seg == 0:
            return sumsq
        need = int(_lib.load().sgm_surplus_norms_workspace_bytes(self._h, P, NF))
        ws = self.workspace(P, NF, need)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(_lib.load().sgm_surplus_norms(
            self._h, x.data_ptr(), P, x.stride(0) if P > 1 else max(x.shape[1], 1),
            f.data_ptr(), f.shape[0], NF, f.stride(0) if f.shape[0] > 1 else max(NF, 1),
            sumsq.data_ptr(), ws.data_ptr(), ws.numel(), stream))
        return sumsq

    def reload_knobs(self):
        _lib.check(_lib.load().sgm_plan_reload_knobs(self._h))

    def jit_wait(self):
        """Wait for a background build of the plan-specialised kernel; returns its state."""
        return int(_lib.load().sgm_plan_jit_wait(self._h))

    def jit_state(self):
        """(state, description) of the plan-specialised kernel: 1 in use, 0 not built, -1 unavailable."""
        buf = ctypes.create_string_buffer(512)
        state = int(_lib.load().sgm_plan_jit_state(self._h, buf, 512))
        return state, buf.value.decode()

    def take_flags(self):
        torch = torch_cuda()
        flags = ctypes.c_uint32(0)
        stream = torch.cuda.current_stream(torch.device("cuda", self.device)).cuda_stream
        _lib.check(_lib.load().sgm_plan_take_flags(self._h, stream, ctypes.byref(flags)))
        return flags.value

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().sgm_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def signed_sum(plans, signs, x, fs, out=None):
    """out = sum_i signs[i] * plans[i](x; fs[i]) in one ABI call (sgm_eval_signed_sum)."""
    torch = torch_cuda()
    n = len(plans)
    P, NF = x.shape[0], fs[0].shape[1]
    if out is None:
        out = torch.empty((P, NF), dtype=torch.float64, device=x.device)
    need = max(int(_lib.load().sgm_eval_workspace_bytes(p._h, P, NF)) for p in plans)
    ws = torch.empty(need, dtype=torch.uint8, device=x.device)
    handles = (_lib.c_vp * n)(*[p._h for p in plans])
    fptr = (_lib.c_vp * n)(*[f.data_ptr() for f in fs])
    Ms = (_lib.c_i64 * n)(*[f.shape[0] for f in fs])
    ldf = (_lib.c_i64 * n)(*[f.stride(0) if f.shape[0] > 1 else max(NF, 1) for f in fs])
    sg = (_lib.c_f64 * n)(*[float(s) for s in signs])
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.load().sgm_eval_signed_sum(
        handles, sg, n, x.data_ptr(), P, x.stride(0) if P > 1 else max(x.shape[1], 1), fptr, Ms,
        ldf, NF, out.data_ptr(), out.stride(0) if P > 1 else max(NF, 1), ws.data_ptr(),
        ws.numel(), stream))
    return out
