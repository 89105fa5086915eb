"""``MultiOutputGP_GPU`` -- drop-in for the reference class of the same name
(mogp_emulator/MultiOutputGP_GPU.py:27-382): E independent GPs over shared inputs.

One libmogp_b200 handle holds this rank's outputs; every phase of ``fit`` and ``predict`` is one batched launch
over them.  With a communicator (``comm=``) the outputs are block-partitioned over the ranks (one process per
GPU, ``sharding.py``) and ``predict`` ends with a single NCCL all-gather of the packed posteriors, so every rank
returns the full ``(n_emulators, n_predict)`` arrays (and ``(n_emulators, n_predict, D)`` derivatives).  Values
follow the CPU ``MultiOutputGP`` (MultiOutputGP.py:182-319, 331-459).
"""
import numpy as np

from . import libmogp
from .GaussianProcessGPU import PredictResult, GPUUnavailableError, _check_mean
from .hyper import GPParams, GPPriors, make_priors
from .kernels import interpret_kernel
from .meanfunc import design_matrix, design_matrix_inputderiv, MeanFit
from .sharding import shard_bounds, gathered_rows, pack_block, unpack_gathered


class MultiOutputGP_GPU(object):
    def __init__(self, inputs, targets, mean=None, kernel="SquaredExponential", priors=None, nugget="adaptive",
                 inputdict={}, use_patsy=True, batch_size=16000, device=0, comm=None, n_streams=0):
        if not libmogp.HAVE_LIBMOGP:
            raise RuntimeError("Cannot construct MultiOutputGP_GPU: the GPU library (libmogp_b200) could not be "
                               "loaded: " + libmogp.last_error())
        if not libmogp.gpu_usable():
            raise RuntimeError("Cannot construct MultiOutputGP_GPU: a compatible GPU could not be found")
        inputs = np.array(inputs, dtype=np.float64)
        targets = np.array(targets, dtype=np.float64)
        if inputs.ndim == 1:
            inputs = np.reshape(inputs, (-1, 1))
        if targets.ndim == 1:
            targets = np.reshape(targets, (1, -1))
        elif targets.ndim != 2:
            raise ValueError("targets must be either a 1D or 2D array")
        if inputs.ndim != 2:
            raise ValueError("inputs must be either a 1D or 2D array")
        if inputs.shape[0] != targets.shape[1]:
            raise ValueError("the first dimension of inputs must be the same length as the second dimension of "
                             "targets (or first if targets is 1D))")
        if isinstance(nugget, str):
            if nugget == "adaptive":
                nugtype, nugsize = libmogp.nugget_type.adaptive, 0.0
            elif nugget == "fit":
                nugtype, nugsize = libmogp.nugget_type.fit, 0.0
            else:
                raise ValueError("nugget must be a string set to 'adaptive', 'fit', or a float")
        elif isinstance(nugget, (float, int)) and not isinstance(nugget, bool):
            if nugget < 0.0:
                raise ValueError("nugget parameter must be non-negative")
            nugtype, nugsize = libmogp.nugget_type.fixed, float(nugget)
        else:
            raise TypeError("nugget parameter must be a string or a non-negative float")
        self._mean_spec = _check_mean(mean)      # one mean function shared by all emulators: None or a MeanFormula
        self._inputs = np.ascontiguousarray(inputs)
        self._targets = np.ascontiguousarray(targets)
        self._dm = design_matrix(self._mean_spec, self._inputs)
        self._nugget_type = nugtype
        self._nugget_size = nugsize
        self.kernel_type, self.kernel = interpret_kernel(kernel)
        self._comm = comm
        E = self._targets.shape[0]
        rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
        self._lo, self._hi, self._e_pad = shard_bounds(E, rank, world)
        self._handle = None
        if self._hi > self._lo:
            self._handle = libmogp.Handle(self._inputs, self._targets[self._lo:self._hi], self.kernel_type, nugtype,
                                          nugsize, device=device, n_streams=n_streams)
        self._thetas = [GPParams(self.D, self.nugget_type, nugsize if nugtype == libmogp.nugget_type.fixed else None)
                        for _ in range(E)]
        self._logpost_data = [None] * E
        self._fit = [False] * E                # outputs of other ranks: their status as of the last exchange (fit / predict)
        self._meanfit = [None] * E
        if isinstance(priors, (GPPriors, dict)) or priors is None:
            priorslist = E * [priors]
        else:
            priorslist = list(priors)
            assert len(priorslist) == E, "Bad length for list provided for priors to MultiOutputGP"
        # priors only enter logposterior / logpost_deriv; default ones (a root find per input dimension,
        # Priors.py:698-760) are built on first use and shared by all emulators that did not get their own
        self._priors_args = priorslist
        self._priors = None

    # -- properties (MultiOutputGP_GPU.py:146-180) ------------------------------------------------------
    @property
    def inputs(self):
        return self._inputs

    @property
    def targets(self):
        return self._targets

    @property
    def D(self):
        return self._inputs.shape[1]

    @property
    def n(self):
        return self._inputs.shape[0]

    @property
    def nugget_type(self):
        return self._nugget_type.name

    @property
    def nugget(self):
        """Nugget of each emulator: the fixed value, or what the last fit used (0.0 before any fit)."""
        if self._nugget_type == libmogp.nugget_type.fixed:
            return self._nugget_size
        return [0.0 if t.nugget is None else t.nugget for t in self._thetas]

    @property
    def n_emulators(self):
        return self._targets.shape[0]

    @property
    def n_params(self):
        return [t.n_data for t in self._thetas]

    @property
    def n_corr(self):
        return [self.D] * self.n_emulators

    @property
    def priors(self):
        if self._priors is None:
            shared_default = None
            built = []
            for p in self._priors_args:
                if p is None:
                    if shared_default is None:
                        shared_default = make_priors(None, self._inputs, self.D, self.nugget_type)
                    built.append(shared_default)
                else:
                    built.append(make_priors(p, self._inputs, self.D, self.nugget_type))
            self._priors = built
        return self._priors

    @property
    def thetas(self):
        return self._thetas

    @property
    def local_range(self):
        """[lo, hi) -- the outputs held by this rank."""
        return self._lo, self._hi

    def reset_emulator(self, index):
        """Mark one emulator "not fit"."""
        if self._lo <= index < self._hi:
            self._handle.reset(index - self._lo)
        self._fit[index] = False
        self._thetas[index].unset_data()
        self._logpost_data[index] = None
        self._meanfit[index] = None

    def reset_fit_status(self):
        if self._handle is not None:
            self._handle.reset(-1)
        for i in range(self.n_emulators):
            self._fit[i] = False
            self._thetas[i].unset_data()
            self._logpost_data[i] = None
            self._meanfit[i] = None

    # -- fitting (MultiOutputGP_GPU.py:309-326; CPU semantics MultiOutputGP.py:331-360) -----------------
    def _apply_mean(self, indices, quad, logdet):
        """Analytic mean coefficients of freshly fitted (local) emulators: one batched K^-1 H solve per design-matrix
        column, the n_mean x n_mean algebra on the host (meanfunc.MeanFit), then the device's alpha / gradient vectors."""
        if self._dm.shape[1] == 0 or not indices:
            return
        local = [i - self._lo for i in indices]
        M = self._dm.shape[1]
        cols = [self._handle.solve_list(local, np.tile(self._dm[:, q], (len(local), 1))) for q in range(M)]
        good, alphas, Us = [], [], []
        for k, i in enumerate(indices):
            t = self._handle.get(local[k], libmogp.GET_ALPHA)
            W = np.column_stack([cols[q][k] for q in range(M)])
            try:
                mf = MeanFit(self._dm, self._targets[i], t, W, self.n)
            except np.linalg.LinAlgError:
                # H^T K^-1 H numerically not positive definite: this emulator stays "not fit" (host and device), the
                # others of the batch are unaffected (reference: calc_Ainv raises LinAlgError, fitting.py:244-249 skips)
                self._record(i, None, 0.0, 0.0, 0.0, libmogp.ERR_NOT_PD)
                self._handle.reset(local[k])
                continue
            self._meanfit[i] = mf
            self._thetas[i].mean = mf.beta.copy()
            self._logpost_data[i] = mf.data_logpost(float(quad[k]), float(logdet[k]), self.n)
            good.append(local[k])
            alphas.append(mf.alpha_mean)
            Us.append(mf.U.T)
        if good:
            self._handle.set_alpha_list(good, np.array(alphas))
            self._handle.set_mean_vectors_list(good, np.array(Us))

    def _record(self, index, theta, quad, logdet, nug, status):
        self._meanfit[index] = None
        if status == libmogp.OK:
            self._thetas[index].set_data(theta)
            self._thetas[index].nugget = float(nug)
            self._logpost_data[index] = 0.5 * (float(quad) + float(logdet) + self.n * np.log(2.0 * np.pi))
            self._fit[index] = True
        else:
            self._thetas[index].unset_data()
            self._logpost_data[index] = None
            self._fit[index] = False

    def fit(self, thetas):
        """Fit every emulator; ``thetas`` has shape ``(n_emulators, n_params)``.  Emulators whose matrix is not
        positive definite stay "not fit" and are reported by ``get_indices_not_fit`` (fitting.py:180-186)."""
        thetas = np.array(thetas, dtype=np.float64)
        if thetas.ndim == 1:
            thetas = thetas.reshape(1, -1)
        if thetas.shape[0] != self.n_emulators:
            raise RuntimeError("thetas must have first dimension of size n_emulators")
        if thetas.shape[1] != self.n_params[0]:
            raise RuntimeError("bad shape for hyperparameters")
        if self._handle is not None:
            try:
                quad, logdet, nug, status = self._handle.fit(0, thetas[self._lo:self._hi])
            except FloatingPointError:
                # infinite squared distance (calc_r2, Kernel.py:482-483): the call's emulators are left "not fit"
                for i in range(self._lo, self._hi):
                    self._record(i, None, 0.0, 0.0, 0.0, libmogp.ERR_FPE)
                raise
            for k, i in enumerate(range(self._lo, self._hi)):
                self._record(i, thetas[i], quad[k], logdet[k], nug[k], status[k])
            ok = [k for k in range(self._hi - self._lo) if status[k] == libmogp.OK]
            self._apply_mean([self._lo + k for k in ok], quad[ok], logdet[ok])
        self.sync_fit_status()

    def sync_fit_status(self):
        """With a communicator: exchange which outputs are fit (one all-gather of ``e_pad`` words per rank), so that
        ``get_indices_fit`` / ``get_indices_not_fit`` and ``fit_GP_MAP``'s failure check see the outputs of every rank.
        Collective: every rank must call it (``fit`` does)."""
        if self._comm is None:
            return
        block = np.zeros(self._e_pad)
        block[:self._hi - self._lo] = [1.0 if f else 0.0 for f in self._fit[self._lo:self._hi]]
        got = self._comm.allgather(block)
        for r in range(self._comm.world):
            lo, hi, _ = shard_bounds(self.n_emulators, r, self._comm.world)
            if r != self._comm.rank:
                for k in range(hi - lo):
                    self._fit[lo + k] = bool(got[r, k] > 0.5)

    def fit_emulator(self, index, theta):
        theta = np.array(theta, dtype=np.float64).reshape(-1)
        if not 0 <= index < self.n_emulators:
            raise RuntimeError("emulator index out of range")
        if theta.shape[0] != self.n_params[index]:
            raise RuntimeError("bad shape for hyperparameters")
        if self._lo <= index < self._hi:
            try:
                quad, logdet, nug, status = self._handle.fit(index - self._lo, theta)
            except FloatingPointError:
                self._record(index, None, 0.0, 0.0, 0.0, libmogp.ERR_FPE)
                raise
            self._record(index, theta, quad[0], logdet[0], nug[0], status[0])
            if status[0] == libmogp.OK:
                self._apply_mean([index], quad, logdet)

    def logposterior(self, index, theta=None):
        """current_logpost of one (local) emulator, fitting it first when theta is given and differs."""
        if theta is not None:
            theta = np.array(theta, dtype=np.float64).reshape(-1)
            t = self._thetas[index]
            if not t.data_has_been_set() or not np.allclose(theta, t.get_data(), rtol=1.0e-10, atol=1.0e-15):
                self.fit_emulator(index, theta)
                if not self._fit[index]:
                    raise libmogp.NotPositiveDefiniteError("Unable to fit the Gaussian process: matrix not positive definite")
        if not self._fit[index]:
            return None
        return self._logpost_data[index] - self.priors[index].logp(self._thetas[index])

    def logpost_deriv(self, index, theta):
        self.logposterior(index, theta)
        grad = self._handle.logpost_grad(index - self._lo, self.n_params[index])
        return grad - self.priors[index].dlogpdtheta(self._thetas[index])

    def logpost_and_deriv_batch(self, indices, thetas):
        """Negative log-posterior and its gradient of several (local) emulators at once: one batched fit and one
        batched gradient call on the GPU.  Returns ``{index: (value, gradient)}``; an emulator whose matrix is not
        positive definite maps to ``None`` (and is left "not fit")."""
        indices = [int(i) for i in indices]
        thetas = np.array(thetas, dtype=np.float64).reshape(len(indices), -1)
        for i in indices:
            if not self._lo <= i < self._hi:
                raise RuntimeError("emulator %d is not held by this rank" % i)
        local = [i - self._lo for i in indices]
        try:
            quad, logdet, nug, status = self._handle.fit_list(local, thetas)
        except FloatingPointError:
            for i in indices:
                self._record(i, None, 0.0, 0.0, 0.0, libmogp.ERR_FPE)
            raise
        ok = []
        for k, i in enumerate(indices):
            self._record(i, thetas[k], quad[k], logdet[k], nug[k], status[k])
            if status[k] == libmogp.OK:
                ok.append(k)
        out = {i: None for i in indices}
        self._apply_mean([indices[k] for k in ok], quad[ok], logdet[ok])
        ok = [k for k in ok if self._fit[indices[k]]]          # the mean step may have failed for some of them
        if ok:
            grads = self._handle.logpost_grad_list([local[k] for k in ok], self.n_params[indices[ok[0]]])
            for row, k in enumerate(ok):
                i = indices[k]
                pri = self.priors[i]
                out[i] = (self._logpost_data[i] - pri.logp(self._thetas[i]), grads[row] - pri.dlogpdtheta(self._thetas[i]))
        return out

    def get_indices_fit(self):
        return [i for i in range(self.n_emulators) if self._fit[i]]

    def get_indices_not_fit(self):
        """Outputs without hyperparameters.  Sharded emulators report the outputs of other ranks as of the last status
        exchange (every ``fit`` and ``predict`` ends with one; ``sync_fit_status`` after ``fit_emulator`` calls)."""
        return [i for i in range(self.n_emulators) if not self._fit[i]]

    # -- prediction (MultiOutputGP_GPU.py:185-297; CPU semantics MultiOutputGP.py:182-319) ---------------
    def predict(self, testing, unc=True, deriv=True, include_nugget=True, allow_not_fit=False, processes=None,
                full_cov=False):
        """Means / variances ``(n_emulators, m)`` and, with ``deriv=True`` (the reference's default,
        MultiOutputGP_GPU.py:185), mean derivatives ``(n_emulators, m, D)``; NaN rows for emulators that are not fit
        under ``allow_not_fit`` (MultiOutputGP.py:476-546).  Sharded (``comm=``): every rank predicts its own outputs
        and one all-gather leaves the full arrays on every rank; ``full_cov`` is not gathered."""
        testing = np.array(testing, dtype=np.float64)
        if self.D == 1 and testing.ndim == 1:
            testing = np.reshape(testing, (-1, 1))
        elif testing.ndim == 1:
            testing = np.reshape(testing, (1, len(testing)))
        assert testing.ndim == 2, "testing must be a 2D array"
        assert testing.shape[1] == self.D, "second dimension of testing must be the same as the number of input parameters"
        E, m = self.n_emulators, testing.shape[0]
        if self._comm is None:
            if not allow_not_fit and len(self.get_indices_not_fit()) > 0:
                raise ValueError("Hyperparameters have not been fit for this Gaussian Process")
            mean, var, dmean = self._local_predict(testing, unc, deriv, include_nugget, full_cov)
            return PredictResult(mean=mean, unc=var, deriv=dmean)
        if unc and full_cov:
            raise GPUUnavailableError("full predictive covariances are not gathered across ranks: predict them on the "
                                      "rank that holds the output (local_range)")
        world, e_pad, e_loc = self._comm.world, self._e_pad, self._hi - self._lo
        dmean_all = None
        if not deriv and self._dm.shape[1] == 0 and self._handle is not None:
            # the path of the BASELINE metric: device-resident results, ONE ncclAllGather of the packed [e_pad][2][m] blocks
            mean_all, var_all, status_all = self._handle.predict_allgather(self._comm, testing, include_nugget, e_pad)
            # rank r's block starts at row r*e_pad; when the partition is even that is its first global output index and the
            # gathered rows are already in output order (slice, no copy); otherwise pick the rows
            rows = gathered_rows(E, world)
            if rows == list(range(E)):
                mean_all, var_all, status = mean_all[:E], var_all[:E], status_all[:E]
            else:
                mean_all, var_all, status = mean_all[rows], var_all[rows], status_all[rows]
        else:
            # mean function and / or derivatives (or a rank without outputs, which joins the same collective with an all-padding
            # block): each rank finishes its own posteriors on the host, then one all-gather of the packed block (sharding.py)
            D = self.D if deriv else 0
            if e_loc:
                mean, var, dmean = self._local_predict(testing, True, deriv, include_nugget, False)
            else:
                mean, var, dmean = np.empty((0, m)), np.empty((0, m)), (np.empty((0, m, self.D)) if deriv else None)
            block = pack_block(mean, var, self._fit[self._lo:self._hi], e_pad, deriv=dmean if deriv else None)
            got = unpack_gathered(self._comm.allgather(block), E, world, m, d=D)
            mean_all, var_all, status = got[0], got[1], got[2]
            if deriv:
                dmean_all = got[3]
        for i in range(E):
            if not self._lo <= i < self._hi:
                self._fit[i] = bool(status[i] == libmogp.OK)
        if not allow_not_fit and np.any(status != libmogp.OK):
            raise ValueError("Hyperparameters have not been fit for this Gaussian Process")
        return PredictResult(mean=mean_all, unc=var_all if unc else None, deriv=dmean_all if deriv else None)

    def _local_predict(self, testing, unc, deriv, include_nugget, full_cov):
        """Posterior of the outputs this rank holds (local index k <-> output lo + k): mean (e, m), var (e, m) or
        covariances (e, m, m) or None, derivatives (e, m, D) or None; NaN rows for outputs that are not fit."""
        e, m = self._hi - self._lo, testing.shape[0]
        dmean = self._handle.predict_deriv(testing)[0] if deriv else None
        if self._dm.shape[1] > 0:
            return self._local_predict_with_mean(testing, unc, include_nugget, full_cov, dmean)
        if unc and full_cov:
            # (e, m, m) covariances, one output at a time (MultiOutputGP.py:183, 303); NaN for unfit emulators
            mean = np.full((e, m), np.nan)
            cov = np.full((e, m, m), np.nan)
            for k in range(e):
                if self._fit[self._lo + k]:
                    mean[k], cov[k] = self._handle.predict_cov(k, testing, include_nugget=include_nugget)
            return mean, cov, dmean
        mean, var, _ = self._handle.predict(testing, want_var=unc, include_nugget=include_nugget)
        return mean, (var if unc else None), dmean

    def _local_predict_with_mean(self, testing, unc, include_nugget, full_cov, dmean):
        """Posterior with the analytic mean function (GaussianProcess.py:887-920): mean shifted by H* beta, variance plus
        R^T A^-1 R with R = H*^T - H^T K^-1 K* (one fused kernel-matrix pass per design-matrix column), clipped last."""
        e, m, lo = self._hi - self._lo, testing.shape[0], self._lo
        Hs = design_matrix(self._mean_spec, testing)
        M = Hs.shape[1]
        fit = [k for k in range(e) if self._fit[lo + k]]
        if dmean is not None and self._mean_spec.terms:
            dHs = design_matrix_inputderiv(self._mean_spec, testing)          # (m, D, M)
            for k in fit:
                dmean[k] += np.dot(dHs, self._meanfit[lo + k].beta)
        extra = {}
        if unc and fit:
            vecs = np.zeros((e, M, self.n))
            for k in fit:
                vecs[k] = self._meanfit[lo + k].W.T
            dots = self._handle.kstar_dot(testing, vecs)
            extra = {k: self._meanfit[lo + k].variance_term(Hs, dots[k], full_cov=full_cov) for k in fit}
        if unc and full_cov:
            mean = np.full((e, m), np.nan)
            cov = np.full((e, m, m), np.nan)
            for k in fit:
                mean[k], cov[k] = self._handle.predict_cov(k, testing, include_nugget=include_nugget)
                mean[k] += np.dot(Hs, self._meanfit[lo + k].beta)
                cov[k] += extra[k]
            return mean, cov, dmean
        mean, var, _ = self._handle.predict(testing, want_var=2 if unc else 0, include_nugget=include_nugget)
        for k in fit:
            mean[k] += np.dot(Hs, self._meanfit[lo + k].beta)
            if unc:
                var[k] = np.maximum(var[k] + extra[k], 0.0)
        return mean, (var if unc else None), dmean

    def __call__(self, testing, processes=None):
        return self.predict(testing, unc=False, deriv=False, processes=processes)[0]

    def timings(self, reset=False):
        return self._handle.timings(reset) if self._handle is not None else {}

    def close(self):
        """Release the device object now (its buffers go back to the library's cache) instead of at garbage collection."""
        if self._handle is not None:
            self._handle.close()
            self._handle = None

    def __str__(self):
        return ("Multi-Output Gaussian Process with:\n" + str(self.n_emulators) + " emulators\n" +
                str(self.n) + " training examples\n" + str(self.D) + " input variables")
