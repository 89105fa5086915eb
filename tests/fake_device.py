"""TEST DOUBLE for ``libmogp.Handle`` -- numpy arithmetic behind the same method contract.

Purpose: run the *host-side* logic of the front-end classes (GaussianProcessGPU / MultiOutputGP_GPU / fit_GP_MAP /
validation: argument checking, mean-function algebra, bookkeeping of fit status, shapes, exceptions) in the CPU test
suite, where no B200 exists.  It lives under tests/, is installed only by the ``fake_gpu`` fixture through monkeypatching,
and is never importable from the product: ``mogp_emulator_b200`` itself has no CPU path and raises without the CUDA
library and a device.  The arithmetic comes from the oracle (oracle/gp_oracle.py), so the GPU parity tests
(tests/test_gpu_parity.py, ``-m gpu``) remain the only statement about the CUDA kernels.
"""
import numpy as np
import scipy.linalg

import gp_oracle as orc
from mogp_emulator_b200 import libmogp


class FakeHandle(object):
    def __init__(self, inputs, targets, kernel, nug_type, nugget=0.0, device=0, n_streams=0):
        self.X = libmogp.as_f64(inputs)
        self.Y = libmogp.as_f64(targets)
        assert self.X.ndim == 2 and self.Y.ndim == 2 and self.Y.shape[1] == self.X.shape[0]
        self.n, self.d = self.X.shape
        self.n_out = self.Y.shape[0]
        self.kernel = orc.SQEXP if int(kernel) == 0 else orc.MAT52
        self.nug_type = libmogp.nugget_type(int(nug_type)).name
        self.nug_fixed = float(nugget)
        self.state = [None] * self.n_out
        self.closed = False

    # -- fitting -----------------------------------------------------------------------------------
    def _fit_one(self, o, theta):
        theta = np.asarray(theta, dtype=np.float64)
        d = self.d
        K = np.exp(theta[d]) * orc.kernel_f(self.X, self.X, theta[:d], self.kernel)
        nug = {"fixed": self.nug_fixed, "fit": float(np.exp(theta[-1])) if self.nug_type == "fit" else None,
               "adaptive": None}[self.nug_type]
        try:
            L, used = orc.cholesky_factor(K.copy(), nug, self.nug_type)
        except (scipy.linalg.LinAlgError, np.linalg.LinAlgError, FloatingPointError):
            self.state[o] = None
            return 0.0, 0.0, 0.0, libmogp.ERR_NOT_PD
        used = float(nug if self.nug_type != "adaptive" else used)
        alpha = orc.cho_solve(L, self.Y[o])
        self.state[o] = dict(theta=theta.copy(), K=K, L=L, alpha=alpha, nugget=used, U=np.zeros((0, self.n)))
        return float(np.dot(self.Y[o], alpha)), float(orc.logdet(L)), used, libmogp.OK

    def fit(self, first, thetas):
        thetas = np.atleast_2d(libmogp.as_f64(thetas))
        return self.fit_list(range(int(first), int(first) + thetas.shape[0]), thetas)

    def fit_list(self, indices, thetas):
        idx = [int(i) for i in indices]
        thetas = libmogp.as_f64(thetas).reshape(len(idx), -1)
        n_params = self.d + 1 + int(self.nug_type == "fit")
        if thetas.shape[1] != n_params:
            raise RuntimeError("mogp_fit: expected %d hyperparameters" % n_params)
        res = [self._fit_one(o, th) for o, th in zip(idx, thetas)]
        quad, logdet, nug, status = (np.array(c) for c in zip(*res))
        return quad, logdet, nug, status.astype(np.int32)

    def reset(self, idx=-1):
        for o in (range(self.n_out) if idx < 0 else [idx]):
            self.state[o] = None

    def is_fit(self, idx):
        return self.state[idx] is not None

    def _need(self, o):
        if self.state[o] is None:
            raise ValueError("hyperparameters have not been fit for this Gaussian Process")
        return self.state[o]

    # -- analytic-mean primitives ----------------------------------------------------------------------
    def solve_list(self, indices, rhs):
        rhs = libmogp.as_f64(rhs).reshape(len(indices), self.n)
        return np.array([orc.cho_solve(self._need(int(o))["L"], r) for o, r in zip(indices, rhs)])

    def set_alpha_list(self, indices, alpha):
        alpha = libmogp.as_f64(alpha).reshape(len(indices), self.n)
        for o, a in zip(indices, alpha):
            self._need(int(o))["alpha"] = a.copy()

    def set_mean_vectors_list(self, indices, U):
        U = libmogp.as_f64(U)
        if U.shape[1] > 32:
            raise RuntimeError("mogp_set_mean_vectors: at most 32 vectors per output")
        for o, u in zip(indices, U):
            self._need(int(o))["U"] = u.reshape(-1, self.n).copy()

    def _kstar(self, o, testing):
        s = self.state[o]
        th = s["theta"]
        return np.exp(th[self.d]) * orc.kernel_f(self.X, testing, th[:self.d], self.kernel)          # (n, m)

    def kstar_dot(self, testing, vecs):
        testing = libmogp.as_f64(testing)
        vecs = libmogp.as_f64(vecs).reshape(self.n_out, -1, self.n)
        out = np.full((self.n_out, vecs.shape[1], testing.shape[0]), np.nan)
        for o in range(self.n_out):
            if self.state[o] is not None:
                out[o] = np.dot(vecs[o], self._kstar(o, testing))
        return out

    # -- prediction --------------------------------------------------------------------------------------
    def predict(self, testing, want_var=True, include_nugget=True):
        testing = libmogp.as_f64(testing)
        m = testing.shape[0]
        mean = np.full((self.n_out, m), np.nan)
        var = np.full((self.n_out, m), np.nan) if want_var else None
        status = np.full(self.n_out, libmogp.ERR_NOT_FIT, dtype=np.int32)
        for o in range(self.n_out):
            s = self.state[o]
            if s is None:
                continue
            status[o] = libmogp.OK
            Ks = self._kstar(o, testing)
            mean[o] = np.dot(Ks.T, s["alpha"])
            if want_var:
                V = scipy.linalg.solve_triangular(s["L"], Ks, lower=True)
                v = np.exp(s["theta"][self.d]) + (s["nugget"] if include_nugget else 0.0) - np.sum(V * V, axis=0)
                var[o] = v if int(want_var) == 2 else np.maximum(v, 0.0)
        return mean, var, status

    def predict_deriv(self, testing):
        testing = libmogp.as_f64(testing)
        deriv = np.full((self.n_out, testing.shape[0], self.d), np.nan)
        status = np.full(self.n_out, libmogp.ERR_NOT_FIT, dtype=np.int32)
        for o in range(self.n_out):
            s = self.state[o]
            if s is None:
                continue
            status[o] = libmogp.OK
            w = np.exp(s["theta"][:self.d])
            for c, x in enumerate(testing):
                diff = x[np.newaxis, :] - self.X
                r2 = np.sum(w * diff ** 2, axis=1)
                coef = np.exp(s["theta"][self.d]) * orc.calc_dKdr2(r2, self.kernel) * s["alpha"]
                deriv[o, c] = 2.0 * w * np.dot(coef, diff)
        return deriv, status

    def predict_cov(self, idx, testing, include_nugget=True):
        s = self._need(int(idx))
        testing = libmogp.as_f64(testing)
        Ks = self._kstar(int(idx), testing)
        th = s["theta"]
        Kss = np.exp(th[self.d]) * orc.kernel_f(testing, testing, th[:self.d], self.kernel)
        if include_nugget:
            Kss = Kss + s["nugget"] * np.eye(testing.shape[0])
        V = scipy.linalg.solve_triangular(s["L"], Ks, lower=True)
        return np.dot(Ks.T, s["alpha"]), Kss - np.dot(V.T, V)

    def predict_allgather(self, comm, testing, include_nugget, e_pad):
        raise RuntimeError("the test double has no communicator")

    def get(self, idx, which):
        s = self._need(int(idx))
        if which == libmogp.GET_K:
            return s["K"].copy()
        if which == libmogp.GET_L:
            return np.tril(s["L"])
        if which == libmogp.GET_ALPHA:
            return s["alpha"].copy()
        return orc.cho_solve(s["L"], np.eye(self.n))

    # -- gradient of the data part of the negative log-posterior -----------------------------------------------
    def logpost_grad(self, idx, n_params):
        s = self._need(int(idx))
        d, th = self.d, s["theta"]
        G = orc.cho_solve(s["L"], np.eye(self.n)) - np.outer(s["alpha"], s["alpha"]) - np.dot(s["U"].T, s["U"])
        r2 = orc.calc_r2(self.X, self.X, th[:d])
        cov = np.exp(th[d])
        dKdr2 = cov * orc.calc_dKdr2(r2, self.kernel)
        grad = np.zeros(n_params)
        for i in range(d):
            diff2 = (self.X[:, i][:, None] - self.X[:, i][None, :]) ** 2
            grad[i] = 0.5 * np.sum(G * dKdr2 * (np.exp(th[i]) * diff2))
        grad[d] = 0.5 * np.sum(G * cov * orc.calc_K(r2, self.kernel))
        if self.nug_type == "fit":
            grad[d + 1] = 0.5 * s["nugget"] * np.trace(G)
        return grad

    def logpost_grad_list(self, indices, n_params):
        return np.array([self.logpost_grad(int(o), n_params) for o in indices])

    def loo_variance(self, idx):
        s = self._need(int(idx))
        return 1.0 / np.diag(orc.cho_solve(s["L"], np.eye(self.n)))

    def timings(self, reset=False):
        return {}

    def close(self):
        self.closed = True
        self.state = [None] * self.n_out
