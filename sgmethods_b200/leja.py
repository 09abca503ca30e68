"""Knot callables for nested weighted-Leja sequences.

The reference ships two 50-point weighted-Leja sequences as MATLAB files
(``sgmethods/knots_weighted_leja_2.mat``, ``sgmethods/knots_weighted_leja_guass.mat``: ``X``
1 x 50 fp64 in Leja order, ``W`` unused) but no code that reads them.
BASELINE config 2 mentions "Gaussian-Leja knots"; this module is the missing glue:

    X = load_leja_sequence("/path/to/knots_weighted_leja_guass.mat")
    interp = SGInterpolant(mid_set, leja_knots(X), lambda nu: nu + 1, TPLagrangeInterpolator)

``leja_knots(X)(n)`` returns the first ``n`` points of the sequence in ascending order (the
operators need ascending knots); the sets are nested for every ``n`` by construction.
"""
import numpy as np


def load_leja_sequence(path):
    """The 1-D array ``X`` of a Leja .mat file (needs scipy.io)."""
    from scipy.io import loadmat
    return np.asarray(loadmat(path)["X"], dtype=np.float64).reshape(-1)


def leja_knots(sequence):
    """Knot callable ``n -> sorted(sequence[:n])`` for SGInterpolant / tp_knots."""
    sequence = np.asarray(sequence, dtype=np.float64).reshape(-1)

    def knots(n):
        n = int(n)
        if n > sequence.size:
            raise ValueError(f"the Leja sequence has only {sequence.size} points, {n} requested")
        return np.sort(sequence[:n])
    return knots
