"""ctypes binding of libmogp_b200.so -- the counterpart of the reference's ``LibGPGPU.py``
(mogp_emulator/LibGPGPU.py:1-14), which imports the pybind11 module ``libgpgpu``.

The shared library is built in-tree by ``__graft_entry__.build()`` / ``build.sh`` and contains sm_100a
code only.  There is no CPU fallback: every entry point needs a B200.  ``gpu_usable()`` mirrors
``LibGPGPU.gpu_usable`` (LibGPGPU.py:13).
"""
import ctypes
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOGP_LIB: a debug build of the same library (tools/*_timeline.py: -DCHOL_TRACE / -DI8_TRACE), never another implementation
LIB_PATH = os.environ.get("MOGP_LIB") or os.path.join(_HERE, "libmogp_b200.so")

OK, ERR_CUDA, ERR_ARG, ERR_NOT_PD, ERR_NOT_FIT, ERR_NCCL, ERR_NOMEM, ERR_FPE = range(8)
GET_K, GET_L, GET_ALPHA, GET_KINV = range(4)


class nugget_type(enum.IntEnum):
    """mogp_gpu/src/types.hpp:29 -- numeric values are part of the Python contract."""
    adaptive = 0
    fit = 1
    fixed = 2


class kernel_type(enum.IntEnum):
    """mogp_gpu/src/types.hpp:32."""
    SquaredExponential = 0
    Matern52 = 1


_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int_p = ctypes.POINTER(ctypes.c_int32)

#: every symbol declared in include/mogp_b200.h with its ctypes signature
SIGNATURES = {
    "mogp_version": (ctypes.c_int, [_c_int_p, _c_int_p]),
    "mogp_device_count": (ctypes.c_int, [_c_int_p]),
    "mogp_last_error": (ctypes.c_char_p, []),
    "mogp_create": (ctypes.c_int, [_c_double_p, ctypes.c_int64, ctypes.c_int32, _c_double_p, ctypes.c_int32,
                                   ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_int32, ctypes.c_int32,
                                   ctypes.POINTER(ctypes.c_void_p)]),
    "mogp_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "mogp_trim": (ctypes.c_int, []),
    "mogp_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, _c_double_p, ctypes.c_int32,
                                _c_double_p, _c_double_p, _c_double_p, _c_int_p]),
    "mogp_fit_list": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, ctypes.c_int32, _c_double_p, ctypes.c_int32,
                                     _c_double_p, _c_double_p, _c_double_p, _c_int_p]),
    "mogp_logpost_grad_list": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, ctypes.c_int32, _c_double_p, ctypes.c_int32]),
    "mogp_solve_list": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, ctypes.c_int32, _c_double_p, _c_double_p]),
    "mogp_set_alpha_list": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, ctypes.c_int32, _c_double_p]),
    "mogp_set_mean_vectors_list": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, ctypes.c_int32, ctypes.c_int32, _c_double_p]),
    "mogp_kstar_dot": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int64, _c_double_p, ctypes.c_int32, _c_double_p]),
    "mogp_reset": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "mogp_is_fit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _c_int_p]),
    "mogp_predict": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                    _c_double_p, _c_double_p, _c_int_p]),
    "mogp_predict_deriv": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int64, _c_double_p, _c_int_p]),
    "mogp_predict_cov": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _c_double_p, ctypes.c_int64, ctypes.c_int32,
                                        _c_double_p, _c_double_p]),
    "mogp_predict_allgather": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, _c_double_p, ctypes.c_int64,
                                              ctypes.c_int32, ctypes.c_int32, _c_double_p, _c_double_p, _c_int_p]),
    "mogp_get": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, _c_double_p]),
    "mogp_logpost_grad": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _c_double_p, ctypes.c_int32]),
    "mogp_loo_variance": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, _c_double_p]),
    "mogp_timings": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int32, ctypes.c_int32]),
    "mogp_chol_schedule": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _c_int_p]),
    "mogp_comm_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "mogp_comm_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.POINTER(ctypes.c_void_p)]),
    "mogp_comm_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "mogp_comm_allreduce_max": (ctypes.c_int, [ctypes.c_void_p, _c_double_p]),
    "mogp_comm_allgather": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int64, _c_double_p]),
    "mogp_peak_dmma": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, _c_double_p]),
    "mogp_peak_i8": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int32, _c_double_p]),
}

_lib = None
_load_error = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        _load_error = "libmogp_b200.so not found at %s (run `python -c 'import __graft_entry__ as g; g.build()'`)" % LIB_PATH
        raise RuntimeError(_load_error)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


try:
    load()
    HAVE_LIBMOGP = True
except (RuntimeError, OSError, AttributeError) as exc:  # pragma: no cover - depends on the build
    HAVE_LIBMOGP = False
    _load_error = str(exc)


def device_count():
    if not HAVE_LIBMOGP:
        return 0
    c = ctypes.c_int32(0)
    _lib.mogp_device_count(ctypes.byref(c))
    return int(c.value)


def gpu_usable():
    """True when the library is loaded and at least one sm_100 device is visible."""
    return HAVE_LIBMOGP and device_count() > 0


def last_error():
    return _lib.mogp_last_error().decode("utf-8", "replace") if _lib is not None else (_load_error or "")


class NotPositiveDefiniteError(RuntimeError):
    """The kernel matrix (or the mean function's H^T K^-1 H) could not be factorised (LAPACK info > 0).  A RuntimeError like
    the reference GPU class raises (densegp_gpu.hpp:556-570); the MAP search skips the restart on exactly this type, so
    CUDA / argument / NCCL failures are never mistaken for it."""


GRAD_MAX_DIMS = 64     # csrc/grad.cu G_MAXD: input dimensions mogp_logpost_grad_list supports


def check(status, what=""):
    """Map a status code to the exception type the reference's front-end raises
    (std::runtime_error -> RuntimeError, densegp_gpu.hpp:495,556-570; ValueError for predict-before-fit,
    GaussianProcessGPU.py:589)."""
    if status == OK:
        return
    msg = "%s%s" % (what + ": " if what else "", last_error())
    if status == ERR_NOT_FIT:
        raise ValueError("hyperparameters have not been fit for this Gaussian Process")
    if status == ERR_NOMEM:
        raise MemoryError(msg)
    if status == ERR_NOT_PD:
        raise NotPositiveDefiniteError(msg)
    if status == ERR_FPE:
        raise FloatingPointError(last_error())      # calc_r2, Kernel.py:482-483
    raise RuntimeError(msg)


def as_f64(a):
    """C-contiguous float64 view/copy (the reference's ndarray_coerce_type_and_flags,
    GaussianProcessGPU.py:29-54)."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def dptr(a):
    return a.ctypes.data_as(_c_double_p)


def iptr(a):
    return a.ctypes.data_as(_c_int_p)


def chol_schedule(n_block_rows, n_outputs):
    """The tile schedule of the factorisation kernel as a list of (kind, output, i, p, j), kind in ("DIAG", "D", "ROW")."""
    names = ("DIAG", "D", "ROW")
    out = np.zeros(5, dtype=np.int32)
    tiles = []
    for t in range(n_outputs * n_block_rows * (n_block_rows + 2)):
        check(_lib.mogp_chol_schedule(t, int(n_block_rows), int(n_outputs), iptr(out)), "mogp_chol_schedule")
        tiles.append((names[out[0]], int(out[1]), int(out[2]), int(out[3]), int(out[4])))
    return tiles


def trim():
    """Release the library's cache of device / pinned buffers."""
    if _lib is not None:
        _lib.mogp_trim()


def peak_dmma_tflops(device=0, iters=20000):
    """Measured DMMA (FP64 tensor pipe) issue peak of the device, TFLOP/s."""
    v = ctypes.c_double(0.0)
    check(_lib.mogp_peak_dmma(int(device), int(iters), ctypes.byref(v)), "mogp_peak_dmma")
    return float(v.value)


def peak_i8_tops(device=0, iters=20000):
    """Measured int8 tcgen05 issue peaks, TOP/s: (N = 256 MMAs -- the chip's int8 tensor peak, the N = 64 shape of the
    predict kernel)."""
    v = (ctypes.c_double * 2)()
    check(_lib.mogp_peak_i8(int(device), int(iters), v), "mogp_peak_i8")
    return float(v[0]), float(v[1])


class Handle(object):
    """Owner of one mogp_handle (a bank of E outputs over shared inputs on one GPU)."""

    def __init__(self, inputs, targets, kernel, nug_type, nugget=0.0, device=0, n_streams=0):
        if not HAVE_LIBMOGP:
            raise RuntimeError("Cannot construct a GPU Gaussian process: " + (_load_error or "library not loaded"))
        inputs = as_f64(inputs)
        targets = as_f64(targets)
        assert inputs.ndim == 2 and targets.ndim == 2 and targets.shape[1] == inputs.shape[0]
        self.n, self.d = inputs.shape
        self.n_out = targets.shape[0]
        self._h = ctypes.c_void_p()
        check(_lib.mogp_create(dptr(inputs), self.n, self.d, dptr(targets), self.n_out, int(kernel), int(nug_type),
                               float(nugget), int(device), int(n_streams), ctypes.byref(self._h)), "mogp_create")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value and _lib is not None:
            _lib.mogp_destroy(self._h)
            self._h = ctypes.c_void_p()

    __del__ = close

    def fit(self, first, thetas):
        thetas = as_f64(thetas)
        if thetas.ndim == 1:
            thetas = thetas.reshape(1, -1)
        count, n_params = thetas.shape
        quad = np.zeros(count)
        logdet = np.zeros(count)
        nug = np.zeros(count)
        status = np.zeros(count, dtype=np.int32)
        check(_lib.mogp_fit(self._h, int(first), int(count), dptr(thetas), int(n_params), dptr(quad), dptr(logdet),
                            dptr(nug), iptr(status)), "mogp_fit")
        return quad, logdet, nug, status

    def fit_list(self, indices, thetas):
        """Fit the listed (distinct, handle-local) outputs with one batched call; arrays come back in list order."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        thetas = as_f64(thetas).reshape(len(idx), -1)
        count, n_params = thetas.shape
        quad = np.zeros(count)
        logdet = np.zeros(count)
        nug = np.zeros(count)
        status = np.zeros(count, dtype=np.int32)
        check(_lib.mogp_fit_list(self._h, iptr(idx), int(count), dptr(thetas), int(n_params), dptr(quad), dptr(logdet),
                                 dptr(nug), iptr(status)), "mogp_fit_list")
        return quad, logdet, nug, status

    def logpost_grad_list(self, indices, n_params):
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        out = np.zeros((len(idx), n_params))
        check(_lib.mogp_logpost_grad_list(self._h, iptr(idx), len(idx), dptr(out), int(n_params)), "mogp_logpost_grad_list")
        return out

    # -- primitives of the analytic mean function ------------------------------------------------------
    def solve_list(self, indices, rhs):
        """K_i^-1 rhs[i] for the listed fitted outputs; rhs (count, n)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        rhs = as_f64(rhs).reshape(len(idx), self.n)
        out = np.empty_like(rhs)
        check(_lib.mogp_solve_list(self._h, iptr(idx), len(idx), dptr(rhs), dptr(out)), "mogp_solve_list")
        return out

    def set_alpha_list(self, indices, alpha):
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        alpha = as_f64(alpha).reshape(len(idx), self.n)
        check(_lib.mogp_set_alpha_list(self._h, iptr(idx), len(idx), dptr(alpha)), "mogp_set_alpha_list")

    def set_mean_vectors_list(self, indices, U):
        """U: (count, n_vec, n)."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        U = as_f64(U)
        n_vec = U.shape[1]
        check(_lib.mogp_set_mean_vectors_list(self._h, iptr(idx), len(idx), int(n_vec), dptr(U)), "mogp_set_mean_vectors_list")

    def kstar_dot(self, testing, vecs):
        """vecs (n_out, n_vec, n) -> (n_out, n_vec, m): k_o(X*, X) vecs[o][q] for every fitted output."""
        testing = as_f64(testing)
        vecs = as_f64(vecs).reshape(self.n_out, -1, self.n)
        out = np.empty((self.n_out, vecs.shape[1], testing.shape[0]))
        check(_lib.mogp_kstar_dot(self._h, dptr(testing), testing.shape[0], dptr(vecs), int(vecs.shape[1]), dptr(out)),
              "mogp_kstar_dot")
        return out

    def reset(self, idx=-1):
        check(_lib.mogp_reset(self._h, int(idx)))

    def is_fit(self, idx):
        out = ctypes.c_int32(0)
        check(_lib.mogp_is_fit(self._h, int(idx), ctypes.byref(out)))
        return bool(out.value)

    def predict(self, testing, want_var=True, include_nugget=True):
        testing = as_f64(testing)
        m = testing.shape[0]
        mean = np.empty((self.n_out, m))
        var = np.empty((self.n_out, m)) if want_var else None
        status = np.zeros(self.n_out, dtype=np.int32)
        check(_lib.mogp_predict(self._h, dptr(testing), m, int(want_var) if want_var in (0, 1, 2) else 1, int(bool(include_nugget)), dptr(mean),
                                dptr(var) if want_var else None, iptr(status)), "mogp_predict")
        return mean, var, status

    def predict_deriv(self, testing):
        testing = as_f64(testing)
        m = testing.shape[0]
        deriv = np.empty((self.n_out, m, self.d))
        status = np.zeros(self.n_out, dtype=np.int32)
        check(_lib.mogp_predict_deriv(self._h, dptr(testing), m, dptr(deriv), iptr(status)), "mogp_predict_deriv")
        return deriv, status

    def predict_cov(self, idx, testing, include_nugget=True):
        testing = as_f64(testing)
        m = testing.shape[0]
        mean = np.empty(m)
        cov = np.empty((m, m))
        check(_lib.mogp_predict_cov(self._h, int(idx), dptr(testing), m, int(bool(include_nugget)), dptr(mean), dptr(cov)),
              "mogp_predict_cov")
        return mean, cov

    def predict_allgather(self, comm, testing, include_nugget, e_pad):
        testing = as_f64(testing)
        m = testing.shape[0]
        rows = comm.world * e_pad
        mean = np.empty((rows, m))
        var = np.empty((rows, m))
        status = np.zeros(rows, dtype=np.int32)
        check(_lib.mogp_predict_allgather(self._h, comm._c, dptr(testing), m, int(bool(include_nugget)), int(e_pad),
                                          dptr(mean), dptr(var), iptr(status)), "mogp_predict_allgather")
        return mean, var, status

    def get(self, idx, which):
        shape = (self.n,) if which == GET_ALPHA else (self.n, self.n)
        out = np.zeros(shape)
        check(_lib.mogp_get(self._h, int(idx), int(which), dptr(out)), "mogp_get")
        return out

    def logpost_grad(self, idx, n_params):
        out = np.zeros(n_params)
        check(_lib.mogp_logpost_grad(self._h, int(idx), dptr(out), int(n_params)), "mogp_logpost_grad")
        return out

    def loo_variance(self, idx):
        """Leave-one-out predictive variances of all n training points of a fitted output."""
        out = np.zeros(self.n)
        check(_lib.mogp_loo_variance(self._h, int(idx), dptr(out)), "mogp_loo_variance")
        return out

    def timings(self, reset=False):
        out = np.zeros(19)
        check(_lib.mogp_timings(self._h, dptr(out), 19, int(reset)))
        keys = ["kmat_ms", "chol_ms", "solve_ms", "kstar_ms", "trsm_ms", "grad_ms", "n_trsm", "n_launches", "fit_ms",
                "predict_device_wall_ms", "predict_d2h_wall_ms", "i8_prep_ms", "i8_check_ms", "i8_rows_ms", "i8_block_rows",
                "i8_fallbacks", "chol_i8_outputs", "chol_i8_failures_rechecked", "chol_i8_failures_overturned"]
        return dict(zip(keys, out.tolist()))


class Comm(object):
    """One NCCL communicator rank (mogp_comm)."""

    def __init__(self, uid, rank, world, device):
        self.rank, self.world = int(rank), int(world)
        self._c = ctypes.c_void_p()
        check(_lib.mogp_comm_create(uid, self.rank, self.world, int(device), ctypes.byref(self._c)), "mogp_comm_create")

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(128)
        check(_lib.mogp_comm_unique_id(buf), "mogp_comm_unique_id")
        return buf.raw

    def allreduce_max(self, value):
        v = ctypes.c_double(float(value))
        check(_lib.mogp_comm_allreduce_max(self._c, ctypes.byref(v)), "mogp_comm_allreduce_max")
        return float(v.value)

    def allgather(self, block):
        """(count,) float64 per rank -> (world, count): one ncclAllGather between host buffers."""
        block = as_f64(block).reshape(-1)
        out = np.empty((self.world, block.size))
        check(_lib.mogp_comm_allgather(self._c, dptr(block), block.size, dptr(out)), "mogp_comm_allgather")
        return out

    def close(self):
        if getattr(self, "_c", None) is not None and self._c.value and _lib is not None:
            _lib.mogp_comm_destroy(self._c)
            self._c = ctypes.c_void_p()

    __del__ = close
