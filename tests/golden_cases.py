"""How each interpolant fixture in tests/golden/ was set up (mirrors make_golden.py) but
expressed with THIS repo's host API, so the same case can be rebuilt by the oracle and by
the CUDA path on the GPU box (where /root/reference does not exist)."""
import numpy as np

# name -> (knot family, lev2knots name, operator kind)
CASES = {
    "lin_deg1": ("equi", "n+1", "linear"),
    "cub_deg3": ("equi", "3n+1", "cubic"),
    "lin_exp_cc": ("cc", "n+1", "linear"),
    "lin_gauss_rf": ("optg", "2^(n+1)-1", "linear"),
    "lin_cc_nested": ("cc", "doubling", "linear"),
    "quad_cc_nested": ("cc", "doubling", "quadratic"),
    "quad_aniso5": ("equi_arr", "doubling", "quadratic"),
    "cub_smolyak3": ("equi_arr", "3n+1", "cubic"),
    "lag_example0": ("cc", "example0", "lagrange"),
    "lin_example0": ("cc", "example0", "linear"),
    "lag_hermite": ("hermite", "n+1", "lagrange"),
    "lin_example1": ("unb", "2^(n+1)-1", "linear"),
    "lin_c2_small": ("unb", "2^(n+1)-1", "linear"),
    "lag_leja_gauss": ("leja", "n+1", "lagrange"),
}

LEV2KNOTS = {
    "n+1": lambda n: n + 1,
    "3n+1": lambda n: 3 * n + 1,
    "2^(n+1)-1": lambda n: 2 ** (n + 1) - 1,
    "doubling": lambda n: np.where(n == 0, 1, 2 ** n + 1),
    "example0": lambda n: np.where(0, 1, 2 ** n + 1),
}


def knot_family(name, nodes_module):
    n1d = nodes_module
    if name == "leja":          # the first points of the reference's Gaussian Leja sequence
        import os
        from sgmethods_b200.leja import leja_knots
        here = os.path.dirname(os.path.abspath(__file__))
        with np.load(os.path.join(here, "golden", "lag_leja_gauss.npz")) as z:
            return leja_knots(z["leja_X"])
    return {
        "equi": n1d.equispaced_nodes,
        "equi_arr": lambda n: np.atleast_1d(np.asarray(n1d.equispaced_nodes(n), dtype=float)),
        "cc": n1d.cc_nodes,
        "optg": lambda n: n1d.optimal_gaussian_nodes(n, p=2),
        "unb": n1d.unbounded_nodes_nested,
        "hermite": n1d.hermite_nodes,
    }[name]


def unpack_plan(g):
    """Lists (dims, shapes, maps) from the packed arrays written by make_golden.pack_plan."""
    T = g["coeffs"].shape[0]
    dims = [g["dims"][g["dims_off"][t]:g["dims_off"][t + 1]] for t in range(T)]
    shapes = [tuple(int(v) for v in g["shapes"][g["shape_off"][t]:g["shape_off"][t + 1]])
              for t in range(T)]
    maps = [g["maps"][g["map_off"][t]:g["map_off"][t + 1]].reshape(shapes[t])
            for t in range(T)]
    return dims, shapes, maps


# ------------------------------------------------------------------ composite drivers
# Multilevel cases (make_golden.drivers): name -> (knot family, lev2knots, kind, levels, NF,
# points, seed).  Level k of a case uses the multi-index set ml_mid_set(mis, name, k).
ML_CASES = {
    "ml_lin_nf1": ("cc", "doubling", "linear", 4, 1, 160, 101),
    "ml_quad_nf3": ("equi_arr", "doubling", "quadratic", 4, 3, 120, 102),
    # the parametric dimension grows with the level (N = 2, 3, 4): the restriction to the
    # coarser grid drops the rows of the finer grid whose tail is non-zero
    # (sparse_grid_interpolant.py:205-212); the newest directions carry level <= 1 so that
    # the reference's int-truncated tail test is exact
    "ml_lin_grow": ("cc", "doubling", "linear", 3, 2, 96, 103),
}


def ml_mid_set(mis, name, k):
    if name == "ml_lin_grow":
        return mis.aniso_smolyak_mid_set(k + 1, k + 2, [1, 1, 2, 3][:k + 2])
    if name == "ml_quad_nf3":
        return mis.smolyak_mid_set(k, 3)
    return mis.smolyak_mid_set(k, 4)


def ML_F_APPROX(y, k):
    """Level-k 'finite element' approximation used for MLInterpolant.sample."""
    return np.array([np.sin(np.sum(y)) + 2.0 ** (-k), float(k) + y[0]])


# Indicator cases: name -> (knot family, lev2knots, kind, script, points, seed); the script is
# a list of adaptive steps, each a list of MidSet operations ("e", margin index) / ("d", None)
# applied before the indicators of that step are computed (old reduced margin = previous
# step's, so steps >= 1 exercise the recycling and the dimension growth).
GG_CASES = {
    "gg_lin_unb": ("unb", "2^(n+1)-1", "linear",
                   [[("e", 0), ("d", None), ("e", 1), ("e", 0), ("d", None), ("e", 2)],
                    [("e", 0)], [("d", None), ("e", 3)]], 300, 201),
    "gg_quad_equi": ("equi_arr", "doubling", "quadratic",
                     [[("e", 0), ("d", None), ("e", 1), ("e", 0)], [("d", None), ("e", 1)]],
                     200, 202),
}


def GG_F(y):
    return np.array([np.exp(-0.3 * np.sum(y * y)), np.sum(y)])


def GG_NORM(a, b):
    return float(np.sqrt(np.mean(np.sum((a - b) ** 2, axis=1))))
