"""Known-answer vectors that the reference's own test-suite pins for the fit+predict path.

Values restated (not code) from /root/reference/mogp_emulator/tests:
  test_Kernel.py:8-38        scaled squared distances
  test_Kernel.py:721-754     SquaredExponential.kernel_f == exp(-r2/2)
  test_Kernel.py:975-987     Matern-5/2 closed form
  test_linalg.py:103-127     3x3 Cholesky factors (fixed nugget)
  test_linalg.py:129-154     jittered Cholesky: jitter must be 1e-6; non-PD input must fail
  test_GaussianProcess.py:1144-1161  variance stability case
Used by tests/test_oracle.py (oracle pinning) and tests/test_gpu_parity.py (CUDA path).
"""
import numpy as np

# ---- distances: (x1, x2, theta_corr, expected r2) ------------------------------------------------
R2_CASES = [
    (np.array([[1.0], [2.0]]), np.array([[2.0], [3.0]]), np.array([0.0]),
     np.array([[1.0, 4.0], [0.0, 1.0]])),
    (np.array([[1.0, 2.0], [2.0, 3.0]]), np.array([[2.0, 4.0], [3.0, 1.0]]), np.array([0.0, 0.0]),
     np.array([[5.0, 5.0], [1.0, 5.0]])),
    (np.array([[1.0, 2.0], [2.0, 3.0]]), np.array([[2.0, 4.0], [3.0, 1.0]]),
     np.array([np.log(2.0), np.log(4.0)]),
     np.array([[18.0, 12.0], [4.0, 18.0]])),
]

# ---- Cholesky -------------------------------------------------------------------------------------
CHOL_WIKI_A = np.array([[4.0, 12.0, -16.0], [12.0, 37.0, -43.0], [-16.0, -43.0, 98.0]])
CHOL_WIKI_L = np.array([[2.0, 0.0, 0.0], [6.0, 1.0, 0.0], [-8.0, 5.0, 3.0]])

_C = 0.0067379469990855
CHOL_NEAR_SINGULAR_A = np.array([[1.0, 1.0, _C], [1.0, 1.0, _C], [_C, _C, 1.0]])   # nugget/jitter 1e-6 goes on top
CHOL_NEAR_SINGULAR_NUGGET = 1.0e-6
CHOL_NEAR_SINGULAR_L = np.array([
    [1.0000004999998751e+00, 0.0000000000000000e+00, 0.0000000000000000e+00],
    [9.9999950000037496e-01, 1.4142132088085626e-03, 0.0000000000000000e+00],
    [6.7379436301144941e-03, 4.7644444411381860e-06, 9.9997779980004420e-01]])

CHOL_NOT_PD_A = np.array([[1.0e-6, 1.0, 0.0], [1.0, 1.0, 1.0], [0.0, 1.0, 1.0e-10]])

# ---- variance stability ---------------------------------------------------------------------------
VAR_STABILITY = dict(
    x=np.linspace(0.0, 5.0, 21).reshape(-1, 1),
    y=np.linspace(0.0, 5.0, 21) ** 2,
    nugget=1.0e-8,
    theta=np.array([-7.352408190715323, 15.041447753599755]),
    testing=np.linspace(0.0, 5.0, 101).reshape(-1, 1),
    atol=1.0e-3,
)


def matern52_closed_form(r2):
    r = np.sqrt(r2)
    return (1.0 + np.sqrt(5.0) * r + 5.0 / 3.0 * r2) * np.exp(-np.sqrt(5.0) * r)


# ---- validation diagnostics: the mocked predictions of mogp_emulator/tests/test_validation.py:8-15, 67-70, 105-110, 145 ----
VALID_TARGETS1 = np.array([0.5, 2.1, 2.8])
VALID_TARGETS2 = np.array([2.5, 2.9, 3.5])
VALID_MEAN1 = np.array([1.0, 2.0, 3.0])
VALID_MEAN2 = np.array([2.0, 3.0, 4.0])
VALID_COV = np.array([[0.1, 0.05, 0.02], [0.05, 0.2, 0.01], [0.02, 0.01, 0.15]])
VALID_ORDER = np.array([1, 2, 0])          # decreasing variance, and the pivoting order of VALID_COV
