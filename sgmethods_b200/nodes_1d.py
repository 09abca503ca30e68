"""1-D knot families (host callables; their outputs are uploaded as fp64 knot tables).

Same names, signatures and return conventions as the reference module
sgmethods/nodes_1d.py (equispaced_nodes :13, equispaced_interior_nodes :30, cc_nodes :49,
hermite_nodes :69, optimal_gaussian_nodes :84, unbounded_nodes_nested :108), including the
python scalar ``0`` returned for a single node by the first three families.  The floating
point expressions are evaluated in the same order, so the knots are bit-identical
(tests/test_host_api.py compares with tests/golden/nodes_1d.npz).
"""
from math import pi, sqrt

import numpy as np
from numpy.polynomial.hermite import hermroots
from scipy.special import erfinv

__all__ = ["equispaced_nodes", "equispaced_interior_nodes", "cc_nodes", "hermite_nodes",
           "optimal_gaussian_nodes", "unbounded_nodes_nested"]


def equispaced_nodes(n):
    """n equispaced nodes on [-1, 1], end points included (0 for n == 1)."""
    return 0 if n == 1 else np.linspace(-1, 1, n)


def equispaced_interior_nodes(n):
    """The n interior nodes of the n+2 equispaced nodes on [-1, 1] (0 for n == 1)."""
    return 0 if n == 1 else np.linspace(-1, 1, n + 2)[1:-1]


def cc_nodes(n):
    """Clenshaw-Curtis nodes x_i = -cos(i pi / (n-1)), i = 0..n-1 (0 for n == 1)."""
    if n == 1:
        return 0
    i = np.linspace(0, n - 1, n)
    return -np.cos(pi * i / (n - 1))


def hermite_nodes(n):
    """Roots of the n-th (physicists') Hermite polynomial."""
    coeffs = np.zeros(n + 1)
    coeffs[n] = 1
    return hermroots(coeffs)


def optimal_gaussian_nodes(n, p=2):
    """n (odd) symmetric nodes c * erfinv(j / (m+1)), j = -m..m, m = (n-1)/2,
    c = sqrt(2 (2p+1)): optimal for degree p-1 piecewise polynomials in L^2 w.r.t. the
    Gaussian measure."""
    assert n % 2 == 1
    m = int((n - 1) / 2)
    grid = np.linspace(-m, m, n) / (m + 1)
    return sqrt(2 * (2 * p + 1)) * erfinv(grid)


def unbounded_nodes_nested(n, p=2):
    """optimal_gaussian_nodes for the nested sizes n = 2^(i+1) - 1.

    The reference's nestedness assertion (nodes_1d.py:127) reduces, by operator precedence,
    to ``n >= 1``; only oddness is enforced (by optimal_gaussian_nodes).  Same here."""
    assert n >= 1
    return optimal_gaussian_nodes(n, p)
