"""Integer helpers for multi-indices plus the small convergence-fit helpers.

Mirrors the public names of the reference's sgmethods/utils.py (compute_level_lc :11,
coord_unit_vector :44, lexic_sort :60, find_mid :77, find_idx_in_margin :103,
checkIfElement :111, mid_is_in_reduced_margin :121, is_downward :154, rate :178,
ls_fit_loglog :199, plot_2d_midset :217).  matplotlib is imported lazily (only
plot_2d_midset needs it), so the package imports on machines without it.
Deviation: ``is_downward`` is implemented (the reference ships a stub returning False).
"""
from math import exp, floor, log2

import numpy as np


def compute_level_lc(i):
    """Linear index of a Levy-Ciesielski basis function -> (level, index in level).

    i = 0 -> (0, 1), i = 1 -> (1, 1), otherwise l = floor(log2 i) + 1, j = i - 2^(l-1)
    (j is 0-based from i = 2 on, exactly as the reference returns it)."""
    if i < 2:
        return i, 1
    lev = floor(log2(i)) + 1
    return lev, i - 2 ** (lev - 1)


def coord_unit_vector(dim, entry):
    """Integer vector of length ``dim`` with a single 1 at position ``entry``."""
    e = np.zeros(dim, dtype=int)
    e[entry] = 1
    return e


def lexic_sort(mid_set):
    """Rows of ``mid_set`` in lexicographic order (first column most significant)."""
    mid_set = np.asarray(mid_set)
    keys = tuple(mid_set[:, c] for c in range(mid_set.shape[1] - 1, -1, -1))
    return mid_set[np.lexsort(keys)]


def find_mid(mid, mid_set):
    """(found, position of the first occurrence or -1) of ``mid`` among the rows."""
    if mid_set.size == 0:
        return False, -1
    assert mid_set.shape[1] == mid.size
    hits = np.flatnonzero(np.all(mid_set == mid, axis=1))
    assert hits.size < 2
    if hits.size == 0:
        return False, -1
    return True, hits[0]


def find_idx_in_margin(margin, mid):
    """Position of ``mid`` in ``margin`` (IndexError if absent), kept for compatibility."""
    return np.where(np.all(margin == mid, axis=1))[0][0]


def checkIfElement(mid, mid_set):
    """True iff ``mid`` is a row of ``mid_set`` (kept for compatibility)."""
    assert mid.size == mid_set.shape[1]
    assert len(mid.shape) == 1
    assert len(mid_set.shape) == 2
    return (mid_set == mid).all(axis=1).any()


def mid_is_in_reduced_margin(mid, mid_set):
    """True iff for every n: mid[n] == 0 or mid - e_n is in ``mid_set``."""
    N = mid_set.shape[1]
    assert N == mid.size
    for n in range(N):
        if mid[n] != 0 and not checkIfElement(mid - coord_unit_vector(N, n), mid_set):
            return False
    return True


def is_downward(mid_set):
    """True iff the set is downward closed (nu in set, nu_n > 0 => nu - e_n in set)."""
    mid_set = np.asarray(mid_set)
    rows = {tuple(r) for r in mid_set.tolist()}
    for r in rows:
        for n, v in enumerate(r):
            if v > 0 and r[:n] + (v - 1,) + r[n + 1:] not in rows:
                return False
    return True


def rate(err, n_dofs):
    """-log(err[i+1]/err[i]) / log(n[i+1]/n[i]) for consecutive pairs."""
    err = np.array(err)
    n_dofs = np.array(n_dofs)
    return -np.log(err[1:] / err[:-1]) / np.log(n_dofs[1:] / n_dofs[:-1])


def ls_fit_loglog(x, y):
    """Least-squares line through (log x, log y): returns (slope, exp(offset))."""
    lx = np.log(x)
    A = np.vstack([lx, np.ones(len(lx))]).T
    slope, offset = np.linalg.lstsq(A, np.log(y), rcond=None)[0]
    return slope, exp(offset)


def plot_2d_midset(midset):
    """Scatter plot of a 1-D or 2-D multi-index set (needs matplotlib)."""
    import matplotlib.pyplot as plt
    assert midset.shape[1] <= 2
    if midset.shape[1] == 1:
        plt.plot(midset, np.zeros(midset.shape[0]), 's')
    else:
        plt.plot(midset[:, 0], midset[:, 1], 's')
