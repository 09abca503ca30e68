"""The Gaussian-Leja knot family of the fixtures (tests/golden/lag_leja_gauss.npz)."""
import os

import numpy as np


def leja():
    from sgmethods_b200.leja import leja_knots
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with np.load(os.path.join(root, "tests", "golden", "lag_leja_gauss.npz")) as z:
        return leja_knots(z["leja_X"])
