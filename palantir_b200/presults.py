"""Result container of run_palantir (mirror of src/palantir/presults.py:24-77; the gene-trend
and clustering functions of that module are downstream consumers and out of scope)."""
from __future__ import annotations

import pickle


class PResults(object):
    """Pseudotime, entropy, fate probabilities and waypoints of one run."""

    def __init__(self, pseudotime, entropy, branch_probs, waypoints, *, _clipped: bool = False):
        self._pseudotime = pseudotime
        self._entropy = entropy
        self._branch_probs = branch_probs
        # presults.py:34 -- probabilities below 1 % are zeroed in place (run_palantir has already
        # done it on the device and says so: pandas' boolean-frame assignment costs 20 ms at 1M cells)
        if not _clipped:
            self._branch_probs[self._branch_probs < 0.01] = 0
        self._waypoints = waypoints

    @property
    def pseudotime(self):
        return self._pseudotime

    @property
    def branch_probs(self):
        return self._branch_probs

    @branch_probs.setter
    def branch_probs(self, value):
        self._branch_probs = value

    @property
    def entropy(self):
        return self._entropy

    @entropy.setter
    def entropy(self, value):
        self._entropy = value

    @property
    def waypoints(self):
        return self._waypoints

    @classmethod
    def load(cls, pkl_file):
        with open(pkl_file, "rb") as f:
            d = pickle.load(f)
        return cls(d["_pseudotime"], d["_entropy"], d["_branch_probs"], d["_waypoints"])

    def save(self, pkl_file: str):
        # the reference passes the file NAME to pickle.dump (presults.py:77), which raises;
        # opening the file is the evident intent
        with open(pkl_file, "wb") as f:
            pickle.dump(vars(self), f)
