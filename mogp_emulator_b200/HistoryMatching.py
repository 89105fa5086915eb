"""History matching over the GPU emulators -- the consumer of ``predict`` on large query sets
(reference: mogp_emulator/HistoryMatching.py:5-703; SURVEY section 8f rank 4, "HistoryMatching on top").

The implausibility of a query point compares observations z_k with the emulator's posterior there,

    I_k(x) = |z_k - E[f_k(x)]| / sqrt( Var[f_k(x)] + model discrepancy_k + observation variance_k ),

and scores a point by the (rank+1)-th largest I_k over the outputs (HistoryMatching.py:197-289).  The expensive part is the
posterior at ``ncoords`` query points for every output: ``gp.predict(coords, unc=True, deriv=False)`` -- for the many query
points history matching uses, that is exactly the many-right-hand-side predict of libmogp_b200 (int8 tcgen05 TRSM).  What
is left for the host is elementwise arithmetic and one ``np.partition``.

Same public surface as the reference class (constructor keywords, ``get_implausibility``, ``get_NROY``, ``get_RO``,
``set_*`` / ``check_*``, ``update``, ``status``, attributes ``gp obs coords expectations ndim ncoords threshold I NROY RO``)
with the reference's exception types.  Fixes, not drift: ``__str__`` no longer fails once ``NROY`` / ``RO`` are lists
(the reference concatenates ``str`` and ``int``, HistoryMatching.py:683-690); predictive derivatives are not requested.
"""
import numpy as np

from .GaussianProcessGPU import GaussianProcessGPU, PredictResult
from .MultiOutputGP_GPU import MultiOutputGP_GPU


class HistoryMatching(object):
    def __init__(self, gp=None, obs=None, coords=None, expectations=None, threshold=3.0):
        self.gp = None
        self.obs = None
        self.coords = None
        self.expectations = None
        self.ndim = None
        self.ncoords = None
        self.threshold = None
        self.I = None
        self.NROY = None
        self.RO = None
        if self.check_gp(gp):
            self.set_gp(gp)
        if self.check_obs(obs):
            self.set_obs(obs)
        if self.check_coords(coords):
            self.set_coords(coords)
        if self.check_expectations(expectations):
            self.set_expectations(expectations)
        if self.check_threshold(threshold):
            self.set_threshold(threshold)
        self.update()

    # -- the calculation -----------------------------------------------------------------------------
    def get_n_obs(self):
        """Number of observed quantities (= number of emulator outputs compared)."""
        return len(self.obs[0])

    def _select_expectations(self):
        """Posterior to score: the one given explicitly, or a prediction of ``gp`` at ``coords`` -- exactly one of the two
        must be set (HistoryMatching.py:155-195)."""
        from_gp = self.check_coords(self.coords) and self.check_gp(self.gp)
        given = self.check_expectations(self.expectations)
        if from_gp and given:
            raise ValueError("Multiple valid parameter combinations are set. Previously set parameters can be removed "
                             "by setting them to None")
        if not from_gp and not given:
            raise ValueError("Expectations are not provided, nor is a GP and coordinates. Must set one in order to "
                             "perform History Matching")
        if self.ncoords is None:
            raise ValueError("ncoords is not set despite a valid parameter combination being found.")
        if given:
            return self.expectations
        return self.gp.predict(self.coords, unc=True, deriv=False)

    def get_implausibility(self, discrepancy=0.0, rank=1):
        """Implausibility of every query point, shape ``(ncoords,)``; ``rank`` = how many of the largest per-output values
        are ignored (0: the maximum; forced to 0 for a single output), ``discrepancy`` = extra variance (scalar or one per
        output)."""
        if not self.check_obs(self.obs):
            raise ValueError("implausibility calculation requires that the observation value is set. This can be done "
                             "using the set_obs method.")
        assert np.all(np.asarray(discrepancy) >= 0.0), "Model discrepancy variance cannot be negative"
        discrepancy = np.atleast_1d(np.asarray(discrepancy, dtype=np.float64))
        post = self._select_expectations()
        mean = np.atleast_2d(post[0])
        var = np.atleast_2d(post[1])
        n_obs = self.get_n_obs()
        assert n_obs == mean.shape[0] and n_obs == var.shape[0]
        if n_obs == 1:
            rank = 0
        assert rank >= 0, "rank must be a non-negative integer"
        assert rank < n_obs, "rank must be less than the number of observations"
        total_var = var + discrepancy[:, np.newaxis] + self.obs[1][:, np.newaxis]
        scores = np.abs(self.obs[0][:, np.newaxis] - mean) / np.sqrt(total_var)
        kth = n_obs - rank - 1
        self.I = np.partition(scores, kth, axis=0)[kth]
        return self.I

    def get_NROY(self, discrepancy=0.0, rank=1):
        """Indices of the query points that are Not Ruled Out Yet (I <= threshold)."""
        if self.I is None:
            self.get_implausibility(discrepancy, rank)
        self.NROY = list(np.where(self.I <= self.threshold)[0])
        return self.NROY

    def get_RO(self, discrepancy=0.0, rank=1):
        """Indices of the query points that are Ruled Out (I > threshold)."""
        if self.I is None:
            self.get_implausibility(discrepancy, rank)
        self.RO = list(np.where(self.I > self.threshold)[0])
        return self.RO

    # -- setters ---------------------------------------------------------------------------------------
    def set_gp(self, gp):
        if not self.check_gp(gp):
            raise TypeError("bad input for set_gp - expects a GaussianProcessGPU or MultiOutputGP_GPU object.")
        self.gp = gp

    def set_obs(self, obs):
        """A number (no observation error), or ``[value(s)]`` / ``[value(s), variance(s)]``."""
        if not self.check_obs(obs):
            raise TypeError("bad input for set_obs")
        if isinstance(obs, (float, int, np.floating, np.integer)):
            self.obs = [np.array([float(obs)]), np.array([0.0])]
            return
        parts = [np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in obs]
        if len(parts) == 1:
            parts.append(np.array([0.0]))
        self.obs = parts

    def set_coords(self, coords):
        if coords is not None and not self.check_coords(coords):
            raise TypeError("bad input for set_coords - expected coords in the form of a 1D or 2D ndarray of numerical values")
        if coords is None:
            self.coords = None
        elif coords.ndim == 1:
            self.coords = np.reshape(coords, (-1, 1))
        else:
            self.coords = coords
        self.update()

    def set_expectations(self, expectations):
        if expectations is not None and not self.check_expectations(expectations):
            raise TypeError("bad input for set_expectations - expected a PredictResult holding mean and variance arrays.")
        self.expectations = expectations
        self.update()

    def set_threshold(self, threshold):
        if not self.check_threshold(threshold):
            raise TypeError("bad input for set_threshold - expected a non-negative float")
        self.threshold = float(threshold)

    # -- checks (True when the argument can be used; None is "not set") -------------------------------------
    def check_gp(self, gp):
        return isinstance(gp, (GaussianProcessGPU, MultiOutputGP_GPU))

    def check_obs(self, obs):
        if obs is None:
            return False
        if isinstance(obs, np.ndarray):
            if obs.ndim > 2:
                raise ValueError("bad input for HistoryMatching, the obs parameter cannot be an array of more than 2D")
            assert obs.shape[0] == 2, "first dimension of observations must have length 2"
        elif isinstance(obs, (list, tuple)):
            if len(obs) > 2:
                raise ValueError("bad input type for HistoryMatching - the specified observation parameter cannot contain "
                                 "more than 2 entries (value, [variance])")
        else:
            try:
                float(obs)
            except (TypeError, ValueError):
                raise TypeError("bad input type for HistoryMatching - the specified observation parameter must contain "
                                "only numerical values")
            return True
        if len(obs) == 2:
            assert np.all(np.asarray(obs[1]) >= 0.0), "variance in observations cannot be negative"
        return True

    def check_coords(self, coords):
        return isinstance(coords, np.ndarray) and coords.ndim <= 2

    def check_expectations(self, expectations):
        if expectations is None or not isinstance(expectations, PredictResult):
            return False
        mean, unc, deriv = expectations[0], expectations[1], expectations[2]
        if not (isinstance(mean, np.ndarray) and isinstance(unc, np.ndarray) and (deriv is None or isinstance(deriv, np.ndarray))):
            raise TypeError("bad input type for HistoryMatching - expected expectation values in the form of a "
                            "PredictResult object with mean and uncertainty set.")
        if mean.shape != unc.shape:
            raise ValueError("bad input for HistoryMatching - mean and variance expectations do not match")
        assert np.all(unc >= 0.0), "all variances must be non-negative"
        return True

    def check_threshold(self, threshold):
        if threshold is None:
            return False
        try:
            value = float(threshold)
        except TypeError:
            return False
        assert value >= 0.0, "threshold must be non-negative"
        return True

    def update(self):
        """Derive ``ndim`` / ``ncoords`` from whatever is set."""
        if self.check_coords(self.coords):
            self.ncoords, self.ndim = self.coords.shape
        elif self.check_expectations(self.expectations):
            self.ncoords = self.expectations[0].shape[-1] if self.expectations[0].ndim > 1 else self.expectations[0].shape[0]

    def status(self):
        print(str(self))

    def __str__(self):
        def shape_of(a):
            return None if a is None else a.shape

        def size_of(lst):
            return None if lst is None else "%s of length %d" % (type(lst), len(lst))

        exp = None
        if self.expectations is not None:
            exp = "%s of len %d containing arrays of shape: %s" % (type(self.expectations), len(self.expectations),
                                                                    self.expectations[0].shape)
        fields = [("Gaussian Process", self.gp), ("Observations", self.obs), ("Coords", shape_of(self.coords)),
                  ("Expectations", exp), ("No. of Input Dimensions", self.ndim),
                  ("No. of Descrete Expectation Values", self.ncoords), ("I_threshold", self.threshold),
                  ("I", None if self.I is None else "%s of shape %s" % (type(self.I), self.I.shape)),
                  ("NROY", size_of(self.NROY)), ("RO", size_of(self.RO))]
        return "History Matching tools created with:\n" + "".join("%s: %s\n" % kv for kv in fields)
