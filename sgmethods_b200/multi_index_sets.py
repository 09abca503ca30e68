"""Generators of downward-closed multi-index sets (host side, integer work, bit-exact).

Same public functions and argument meaning as the reference's sgmethods/multi_index_sets.py
(tensor_product_mid_set :9, aniso_smolyak_mid_set_free_N :47, aniso_smolyak_mid_set :84,
smolyak_mid_set :131, mid_set_profit_free_dim :154, compute_mid_set_fast :197).  Every
function returns an ``int`` ndarray of shape (card, N) whose rows are in lexicographic order.

The rows are produced by depth-first enumeration into a Python list (one allocation at the
end) instead of the reference's repeated ``np.concatenate`` of sub-blocks, but every
floating-point budget / profit operation is performed with the same operands in the same
order (``floor(w / a[0])``, ``w - comp * a[0]``, ``Pmin / profit(nu)``), and the user's
profit callable receives exactly the same sequence of arguments (integer vectors of
shrinking length and finally python ints), because those decide which indices are kept.
Checked against the golden arrays of the reference's tests and tests/golden/index_sets.npz.
"""
from math import floor

import numpy as np

from .utils import coord_unit_vector, lexic_sort

__all__ = ["tensor_product_mid_set", "aniso_smolyak_mid_set_free_N", "aniso_smolyak_mid_set",
           "smolyak_mid_set", "mid_set_profit_free_dim", "compute_mid_set_fast"]


def tensor_product_mid_set(w, N):
    """The full box {0..w}^N (this is what the reference builds, multi_index_sets.py:30-44,
    although its docstring describes a simplex)."""
    assert N >= 1
    assert w >= 0
    if w == 0:
        return np.zeros([1, N], dtype=int)
    axes = np.meshgrid(*([np.arange(w + 1, dtype=int)] * N), indexing="ij")
    return np.stack(axes, axis=-1).reshape(-1, N)


def _aniso_rows(budget, a, pos, prefix, rows):
    """Append, in lexicographic order, all completions of ``prefix`` whose remaining weighted
    sum fits in ``budget`` (same case split and float ops as multi_index_sets.py:108-128)."""
    rest = a.size - pos
    assert budget >= 0
    if budget == 0:
        rows.append(prefix + [0] * rest)
        return
    top = floor(budget / a[pos])
    if rest == 1:
        for comp in range(top + 1):
            rows.append(prefix + [comp])
        return
    for comp in range(top + 1):
        _aniso_rows(budget - comp * a[pos], a, pos + 1, prefix + [comp], rows)


def aniso_smolyak_mid_set(w, N, a):
    """{nu in N_0^N : sum_i a_i nu_i <= w} for a positive weight vector ``a`` of length N."""
    assert N >= 1
    assert w >= 0
    if type(a) == list:  # noqa: E721  (same test as the reference)
        a = np.array(a)
    assert a.size == N
    assert np.all(a > 0)
    rows = []
    _aniso_rows(w, a, 0, [], rows)
    return np.array(rows, dtype=int).reshape(-1, N)


def aniso_smolyak_mid_set_free_N(w, a):
    """Anisotropic Smolyak set where ``a(N)`` returns the weight vector of length N and N is
    the largest dimension whose unit vector still fits: a(N+1)[-1] <= w  =>  N += 1."""
    N = 1
    while a(N + 1)[-1] <= w:
        N += 1
    return aniso_smolyak_mid_set(w, N, a(N))


def smolyak_mid_set(w, N):
    """{nu in N_0^N : |nu|_1 <= w}."""
    assert N >= 1
    return aniso_smolyak_mid_set(w, N, np.repeat(1., N))


def mid_set_profit_free_dim(profit_fun, minimum_profit):
    """{nu : profit(nu) >= Pmin} with the dimension found by probing unit vectors
    e_d of length d+1 until profit(e_d) <= Pmin (multi_index_sets.py:188-194)."""
    d_max = 1
    while profit_fun(coord_unit_vector(d_max + 1, d_max)) > minimum_profit:
        d_max += 1
    return compute_mid_set_fast(profit_fun, minimum_profit, d_max)


def compute_mid_set_fast(profit_fun, minimum_profit, N):
    """{nu in N_0^N : profit(nu) > Pmin} for a multiplicative, monotone profit.

    Recursion on the LAST component (multi_index_sets.py:228-258): its largest admissible
    value is found with unit-direction probes, and for each value v the first N-1 components
    solve the same problem with threshold Pmin / profit((0,..,0,v))."""
    assert N >= 0
    if N == 0:
        return np.array([0], dtype=int)
    if N == 1:
        nu = 0
        assert profit_fun(nu) > 0
        while profit_fun(nu + 1) > minimum_profit:
            nu += 1
        return np.arange(nu + 1, dtype=int).reshape(-1, 1)
    probe = coord_unit_vector(N, N - 1)
    while profit_fun(probe) > minimum_profit:
        probe[-1] += 1
    blocks = []
    for last in range(probe[-1]):
        nu = np.zeros(N, dtype=int)
        nu[-1] = last
        head = compute_mid_set_fast(profit_fun, minimum_profit / profit_fun(nu), N - 1)
        blocks.append(np.hstack((head, np.full((head.shape[0], 1), last))))
    if not blocks:
        return np.zeros((0, N), dtype=int)
    return lexic_sort(np.vstack(blocks))
