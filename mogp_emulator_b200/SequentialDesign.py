"""``MICEFastGP`` over the GPU emulator -- the leave-one-out predictor the MICE sequential design evaluates for every
candidate point (reference: mogp_emulator/SequentialDesign.py:683-748).

The reference corrects a GP fitted to all n candidates for the removal of candidate ``index`` with the Woodbury identity and
returns the predictive variance at that candidate; each call rebuilds K^-1 from the factor with two n x n triangular
solves (O(n^3) per index).  The quantity itself is the Schur complement of K + nugget I with respect to its ``index``-th
row and column, i.e. ``1 / (K^-1)[index, index]`` -- so ONE L^-1 gives all n values: ``mogp_loo_variance`` runs the dataflow
TRSM on an identity right-hand side (the gradient's kernel) and one row-norm kernel, and ``fast_predict`` reads the cached
vector.  The design loop itself (``MICEDesign``, candidate generation, the simulator callback) is host-side orchestration
of the reference and is not rebuilt here.
"""
import numpy as np

from .GaussianProcessGPU import GaussianProcessGPU


class MICEFastGP(GaussianProcessGPU):
    """``GaussianProcessGPU`` with ``fast_predict(index)``: variance of the prediction at training point ``index`` by the GP
    fitted to all the other training points (hyperparameters unchanged)."""

    def fit(self, theta):
        self._loo = None
        super().fit(theta)

    def loo_variances(self):
        """All n leave-one-out variances, clipped at zero (one device call per fit, cached)."""
        if not self.theta.data_has_been_set():
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        if getattr(self, "_loo", None) is None:
            self._loo = np.maximum(self._handle.loo_variance(0), 0.0)
        return self._loo

    def fast_predict(self, index):
        """SequentialDesign.py:705-748: shape ``(1,)`` like the reference's return value."""
        index = int(index)
        assert index >= 0 and index < self.n, "index must be 0 <= index < n"
        return self.loo_variances()[index:index + 1].copy()

    def __getstate__(self):
        state = super().__getstate__()
        state["_loo"] = None
        return state
