"""Adaptive (growable) downward-closed multi-index set -- host side, integer work.

Same class name, constructor, attributes (``mid_set``, ``margin``, ``reduced_margin``,
``max_n``, ``track_reduced_margin``) and methods as the reference's sgmethods/mid_set.py
(MidSet :9, getters :46-91, enlarge_from_margin :93-155, increase_dim :157-216).  All three
arrays are kept as lexicographically sorted ``int`` matrices; the golden sequences of
tests/golden/midset.npz (outputs of the reference) and of the reference's own
tests/test_MidSet.py are reproduced exactly.

Documented deviations (both are defects of the reference):
  * after adding missing backward neighbours recursively, the reference removes the row at
    the *original* position ``idx_margin`` from a margin that has changed meanwhile
    (mid_set.py:129); here the added index is removed by value, which is what the
    reference does whenever its position happens to be unchanged;
  * reaching ``max_n`` prints a message and returns (the reference's message concatenates
    str and number and raises TypeError, mid_set.py:171-172).
"""
import numpy as np
from numpy import inf

from .utils import coord_unit_vector, find_mid, lexic_sort, mid_is_in_reduced_margin


def _sorted_unique(rows):
    return lexic_sort(np.unique(rows, axis=0))


class MidSet():
    """Multi-index set initialised to {0} that grows by indices taken from its margin
    (``enlarge_from_margin``) or by opening a new dimension (``increase_dim``)."""

    def __init__(self, max_n=inf, track_reduced_margin=False):
        self.max_n = max_n
        self.track_reduced_margin = track_reduced_margin
        self.mid_set = np.zeros((1, 1), dtype=int)
        self.margin = np.ones((1, 1), dtype=int)
        self.reduced_margin = np.ones((1, 1), dtype=int)

    # -- sizes ------------------------------------------------------------------------------
    def get_dim(self):
        """Number of dimensions (0 while the set is still {0})."""
        return 0 if self.mid_set.size == 1 else self.mid_set.shape[1]

    def get_cardinality_mid_set(self):
        return self.mid_set.shape[0]

    def get_cardinality_margin(self):
        return self.margin.shape[0]

    def get_cardinality_reduced_margin(self):
        assert self.track_reduced_margin, "Reduced margin not tracked"
        return self.reduced_margin.shape[0]

    # -- growth -----------------------------------------------------------------------------
    def enlarge_from_margin(self, idx_margin):
        """Move margin row ``idx_margin`` into the set; backward neighbours that are still
        missing are added first (recursively) so that the set stays downward closed."""
        assert 0 <= idx_margin < self.margin.shape[0], "idx_margin out of bound"
        nu = np.array(self.margin[idx_margin, :])

        for n in range(self.get_dim()):
            if nu[n] == 0:
                continue
            below = nu - coord_unit_vector(self.get_dim(), n)
            if not find_mid(below, self.mid_set)[0]:
                found, pos = find_mid(below, self.margin)
                if not found:
                    raise IndexError("backward neighbour neither in the set nor in its margin")
                self.enlarge_from_margin(pos)

        self.mid_set = lexic_sort(np.vstack((self.mid_set, nu)))
        dim = self.get_dim()

        keep = ~np.all(self.margin == nu, axis=1)
        forward = nu + np.identity(dim, dtype=int)
        self.margin = _sorted_unique(np.vstack((self.margin[keep], forward)))

        if self.track_reduced_margin:
            found, pos = find_mid(nu, self.reduced_margin)
            assert found, "mid not in reduced margin, error and exit"
            rm = np.delete(self.reduced_margin, pos, 0)
            fresh = [fw for fw in forward if mid_is_in_reduced_margin(fw, self.mid_set)]
            if fresh:
                rm = np.vstack((rm, np.array(fresh, dtype=int)))
            self.reduced_margin = _sorted_unique(rm)

    def increase_dim(self):
        """Append a dimension: add the new unit vector to the set and update the margins."""
        dim = self.get_dim()
        if dim >= self.max_n:
            print("The multi-index set reached the maximum dimension maxN=" + str(self.max_n)
                  + ", cannot increase it further")
            return
        if dim == 0:
            # the reference fails here too (row_stack shape mismatch, mid_set.py:184):
            # the initial set {0} has to be enlarged once before a dimension can be added
            raise ValueError("increase_dim() needs a set with at least one dimension; call "
                             "enlarge_from_margin first")
        e_new = coord_unit_vector(dim + 1, dim)

        def pad(rows):
            return np.column_stack((rows, np.zeros((rows.shape[0], 1), dtype=int)))

        self.mid_set = lexic_sort(np.vstack((pad(self.mid_set), e_new)))
        lifted = self.mid_set[1:].copy()          # every index but 0, moved up by one
        lifted[:, -1] += 1
        self.margin = lexic_sort(np.vstack((pad(self.margin), lifted)))
        if self.track_reduced_margin:
            forward = e_new + np.identity(dim + 1, dtype=int)
            self.reduced_margin = lexic_sort(np.vstack((pad(self.reduced_margin), forward)))
