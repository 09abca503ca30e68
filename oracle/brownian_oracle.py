"""numpy restatement of the reference's Brownian-motion expansions (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/sgmethods/parametric_expansions_brownian_motion.py:
  * lc_brownian   -- :17-71 (Levy-Ciesielski / Faber-Schauder sum, one parameter vector)
  * kl_brownian   -- :73-91 (Karhunen-Loeve sum as shipped)
and /root/reference/sgmethods/utils.py:11-40 (compute_level_lc).

lc_brownian is the literal O(d * nt) double loop over every hat function; the CUDA kernel
evaluates one hat per level instead and must reproduce these numbers bit for bit.

Parity status: PINNED against tests/golden/brownian.npz (reference outputs).
"""
from math import ceil, floor, log, log2, pi, sqrt

import numpy as np


def level_lc(i):
    """Linear index -> (level, index in level): (0,1), (1,1), then l = floor(log2 i) + 1,
    j = i - 2^(l-1) (0-based from i = 2 on), utils.py:31-40."""
    if i == 0:
        return 0, 1
    if i == 1:
        return 1, 1
    lev = floor(log2(i)) + 1
    return lev, i - 2 ** (lev - 1)


def lc_brownian(tt, yy, T):
    """W(t) = sqrt(T) * [ y_0 s + sum_{l=1..L} sum_{j=1..2^(l-1)} y_{2^(l-1)+j-1} 2^((l-1)/2)
    eta_{l,j}(s) ],  s = t/T, L = ceil(log2 len(y)), y zero-padded to 2^L; eta_{l,j} is the hat
    on [(2j-2)/2^l, 2j/2^l] with height 2^-l, both half-intervals closed and the right half
    written last (:47-69)."""
    tt = np.asarray(tt, dtype=float)
    yy = np.asarray(yy, dtype=float).reshape(-1)
    s = tt / T
    L = ceil(log(len(yy), 2))
    yy = np.append(yy, np.zeros(2 ** L - len(yy)))
    W = yy[0] * s
    for lev in range(1, L + 1):
        for j in range(1, 2 ** (lev - 1) + 1):
            left, mid, right = (2 * j - 2) / 2 ** lev, (2 * j - 1) / 2 ** lev, (2 * j) / 2 ** lev
            hat = 0 * s
            up = (s >= left) & (s <= mid)
            hat[up] = s[up] - left
            down = (s >= mid) & (s <= right)
            hat[down] = -s[down] + right
            W = W + yy[2 ** (lev - 1) + j - 1] * 2 ** ((lev - 1) / 2) * hat
    return W * np.sqrt(T)


def kl_brownian(tt, yy):
    """W = sum_n pi (n + 1/2) sqrt(2) sin((n + 1/2) pi t) y_n  -- as shipped, i.e. multiplying
    by the eigenvalue factor (:88-91)."""
    tt = np.asarray(tt, dtype=float)
    W = 0 * tt
    for n in range(len(yy)):
        W = W + pi * (n + 0.5) * sqrt(2) * np.sin((n + 0.5) * pi * tt) * yy[n]
    return W
