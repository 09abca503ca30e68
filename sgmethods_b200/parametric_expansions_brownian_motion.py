"""Parametric expansions of the Wiener process, evaluated on the GPU for batches of
parameter vectors.

Same function names and arguments as the reference's
sgmethods/parametric_expansions_brownian_motion.py (param_LC_Brownian_motion :17-71,
param_KL_Brownian_motion :73-91).  A 1-D ``yy`` gives the reference's 1-D result; a 2-D
``yy`` (one parameter vector per row, which is what the reference's docstrings describe but
its code flattens) is evaluated row by row in ONE kernel launch and returns (P, nt).
numpy in -> numpy out, CUDA tensors in -> CUDA tensor out.  No CPU fallback.

The Levy-Ciesielski kernel evaluates one Faber-Schauder hat per level (the one containing
t/T) instead of all 2^(l-1); every other hat contributes an exact zero in the reference's
sum, so the results are bit-identical (tests/test_gpu_parity.py against reference outputs).
"""
import warnings
from math import ceil, log, pi, sqrt

import numpy as np

from . import _lib


def _prep(tt, yy):
    from ._device import torch_cuda
    torch = torch_cuda()
    as_torch = isinstance(yy, torch.Tensor) or isinstance(tt, torch.Tensor)
    dev = None
    for a in (yy, tt):
        if isinstance(a, torch.Tensor) and a.is_cuda:
            dev = a.device
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device())

    def to_dev(a):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    t_d, y_d = to_dev(tt), to_dev(yy)
    single = y_d.dim() == 1
    if single:
        y_d = y_d.reshape(1, -1)
    return torch, as_torch, dev, t_d.reshape(-1), y_d, single


_SCALE_CACHE = {}


def lc_brownian_device(t_d, y_d, T, out=None):
    """Launch-only form: ``t_d`` (nt,) and ``y_d`` (P, d) are fp64 CUDA tensors; returns the
    (P, nt) CUDA tensor of expansions.  No host synchronisation, no range check."""
    from ._device import torch_cuda
    torch = torch_cuda()
    dev = y_d.device
    d = y_d.shape[1]
    L = ceil(log(d, 2))
    key = (L, dev.index)
    if key not in _SCALE_CACHE:        # 2^((l-1)/2) as the reference computes it (python floats)
        scale = np.array([2 ** ((lev - 1) / 2) for lev in range(1, L + 1)] or [0.0])
        _SCALE_CACHE[key] = torch.from_numpy(scale).to(dev)
    P, nt = y_d.shape[0], t_d.numel()
    W = out if out is not None else torch.empty((P, nt), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.load().sgm_lc_brownian(
        t_d.data_ptr(), nt, y_d.data_ptr(), P, d, y_d.stride(0) if P > 1 else d, float(T),
        float(np.sqrt(T)), _SCALE_CACHE[key].data_ptr(), L, W.data_ptr(),
        W.stride(0) if P > 1 else max(nt, 1), stream))
    return W


def param_LC_Brownian_motion(tt, yy, T):
    """Levy-Ciesielski expansion W(y, t) = sqrt(T) [ y_0 s + sum_{l,j} y_{l,j} 2^((l-1)/2)
    eta_{l,j}(s) ], s = t/T, with the Faber-Schauder hats eta_{l,j} on [0, 1]; the number of
    levels is ceil(log2 d) and the last level may be incomplete (zero padding)."""
    torch, as_torch, dev, t_d, y_d, single = _prep(tt, yy)
    if t_d.numel():
        lo, hi = (np.amin(tt), np.amax(tt)) if isinstance(tt, np.ndarray) else \
            (float(v) for v in torch.aminmax(t_d))
        if hi > T or lo < 0:
            warnings.warn("Warning...........tt not within [0,T]")
    W = lc_brownian_device(t_d, y_d, T)
    if single:
        W = W.reshape(-1)
    return W if as_torch else W.cpu().numpy()


def param_KL_Brownian_motion(tt, yy):
    """Karhunen-Loeve sum as shipped by the reference:
    W = sum_n pi (n + 1/2) sqrt(2) sin((n + 1/2) pi t) y_n."""
    torch, as_torch, dev, t_d, y_d, single = _prep(tt, yy)
    d = y_d.shape[1]
    factor = np.array([pi * (n + 0.5) * sqrt(2) for n in range(d)], dtype=np.float64)
    freq = np.array([(n + 0.5) * pi for n in range(d)], dtype=np.float64)
    P, nt = y_d.shape[0], t_d.numel()
    W = torch.empty((P, nt), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    factor_d = torch.from_numpy(factor).to(dev)      # keep alive until the launch is queued
    freq_d = torch.from_numpy(freq).to(dev)
    _lib.check(_lib.load().sgm_kl_brownian(
        t_d.data_ptr(), nt, y_d.data_ptr(), P, d, y_d.stride(0) if P > 1 else max(d, 1),
        factor_d.data_ptr(), freq_d.data_ptr(), W.data_ptr(), max(nt, 1), stream))
    if single:
        W = W.reshape(-1)
    return W if as_torch else W.cpu().numpy()
