"""sgmethods_b200 -- B200-native evaluation engine for combination-technique sparse-grid
interpolants, behind the host API of andreascaglioni/SGMethods (module and symbol names are
the reference's, e.g. ``sgmethods_b200.sparse_grid_interpolant.SGInterpolant``).

Importing the package never needs a GPU; evaluating an interpolant does (there is no CPU
fallback: the CUDA library ``sgmethods_b200/csrc/libsgm_b200.so`` must be built and a
CUDA device must be present, otherwise the call raises)."""

__version__ = "0.1.0"
