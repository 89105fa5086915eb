"""Exact CPU emulation of the integer arithmetic of csrc/trsm_i8.cu (TEST INFRASTRUCTURE, like the rest of oracle/).

Digits are held in float64 and multiplied with BLAS: every product and partial sum is an integer far below 2^53, so the
result is what tcgen05.mma.kind::i8 with s32 accumulation produces.  Conventions are the kernel's: block 128, signed 7-bit
digits by round-to-nearest, scaled values clamped to +-0.99, row scale of L~ = 2^(frexp exponent of the row maximum + 1),
scale of V = 2^(frexp exponent of sqrt(sigma2 + nugget) + 1), digit pairs (t, u) with t + u <= S + 1, V re-sliced after
every block row, FP64 for blockdiag(L_ii)^-1 L, blockdiag(L_ii)^-1 K* and the final subtraction.
Used to check that the error the GPU path shows against the FP64 path (profiles/r01_i8_check.txt) is the error of the
designed arithmetic, not of its implementation, and as the debugging reference for changes to that kernel.
"""
import numpy as np

BITS = 7
NB = 128


def digits(x, S):
    y = np.clip(x, -0.99, 0.99).copy()
    out = []
    for _ in range(S):
        y = y * 2.0 ** BITS
        d = np.rint(y)
        y = y - d
        out.append(d)
    return out


def sliced_product(A, eA, Vd, S):
    """sum over digit pairs of A_t V_u 2^-7(t+u) with A (rows scaled by 2^-eA) and the digit planes Vd of V."""
    Ad = digits(A * 2.0 ** (-eA)[:, None], S)
    acc = np.zeros((A.shape[0], Vd[0].shape[1]))
    for w in range(2, S + 2):                      # one accumulator per weight, as in TMEM
        part = np.zeros_like(acc)
        for t in range(1, w):
            u = w - t
            if t <= S and u <= S:
                part += Ad[t - 1] @ Vd[u - 1]
        acc += part * 2.0 ** (-BITS * w)
    return acc


def trsm_variance(L, Ks, sigma2, nugget, S, include_nugget=True):
    """L (n, n) lower Cholesky factor of sigma2 k(X, X) + nugget I, Ks (n, m) = sigma2 k(X, X*): predictive variances as
    the int8 path computes them (un-clipped)."""
    n, m = Ks.shape
    n_pad = (n + NB - 1) // NB * NB
    Lp = np.eye(n_pad)
    Lp[:n, :n] = L
    Kp = np.zeros((n_pad, m))
    Kp[:n] = Ks
    eV = int(np.frexp(np.sqrt(sigma2 + nugget))[1]) + 1
    Vd = [np.zeros((n_pad, m)) for _ in range(S)]
    norms = np.zeros(m)
    for i0 in range(0, n_pad, NB):
        blk = slice(i0, i0 + NB)
        Dinv = np.linalg.inv(Lp[blk, blk])
        V = Dinv @ Kp[blk]
        if i0 > 0:
            Lt = Dinv @ Lp[blk, :i0]
            mx = np.abs(Lt).max(axis=1)
            eL = np.where(mx > 0.0, np.frexp(np.where(mx > 0.0, mx, 1.0))[1], 0) + 1
            acc = sliced_product(Lt, eL, [d[:i0] for d in Vd], S)
            V = V - acc * 2.0 ** (eL + eV)[:, None]
        norms += np.sum(V * V, axis=0)
        for t, d in enumerate(digits(V * 2.0 ** -eV, S)):
            Vd[t][blk] = d
    return sigma2 + (nugget if include_nugget else 0.0) - norms
