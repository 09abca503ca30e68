import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
from conftest import load_golden
from sgmethods_b200.parametric_expansions_brownian_motion import param_KL_Brownian_motion
g = load_golden("brownian")
W = param_KL_Brownian_motion(g["kl_tt"], g["kl_yy"])
d = np.abs(W-g["kl_W"]); print(W.shape, d.max(), np.abs(g["kl_W"]).max(), np.unravel_index(d.argmax(), d.shape))
