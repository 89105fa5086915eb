"""Exact CPU emulation of the integer arithmetic of csrc/trsm_i8.cu (TEST INFRASTRUCTURE, like the rest of oracle/).

Digits are held in float64 and multiplied with BLAS: every product and partial sum is an integer far below 2^53, so the
result is what tcgen05.mma.kind::i8 with s32 accumulation produces.  Conventions are the kernel's: block 128, signed 7-bit
digits by round-to-nearest, digit pairs (t, u) with t + u <= S + 1, V re-sliced after every block row.

``trsm_variance`` is the kernel of round 2 (rows-of-L form): the strictly lower blocks of L and every solved block row V_i are
sliced with ONE scale 2^e per output, sqrt(sigma2 + nugget) <= 0.99 2^e; T_i = K*_i - sum_j L_ij V_j from the integer
products; V_i = inv(L_ii) T_i in FP64.  ``trsm_variance_ltilde`` is the kernel of round 1 (L~ = blockdiag(L_ii)^-1 L sliced
with one scale per row, K* pre-multiplied), kept because profiles/r01_i8_check.txt was measured with it.
Used to check that the error the GPU path shows against the FP64 path is the error of the designed arithmetic, not of its
implementation, as the debugging reference for changes to that kernel, and for the study behind the a-posteriori accuracy
check (tools/i8_gate_study.py).
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


def scale_exponent(sigma2, nugget):
    """csrc/trsm_i8.cu i8_scale_exponent: e with sqrt(sigma2 + nugget) <= 0.99 2^e (the leading digit may reach +-127)."""
    f, e = np.frexp(np.sqrt(sigma2 + nugget))
    return int(e) if f <= 0.99 else int(e) + 1


def trsm_variance(L, Ks, sigma2, nugget, S, include_nugget=True):
    """L (n, n) lower Cholesky factor of sigma2 k(X, X) + nugget I, Ks (n, m) = sigma2 k(X, X*): predictive variances as
    i8_trsm_kernel computes them (un-clipped)."""
    n, m = Ks.shape
    n_pad = (n + NB - 1) // NB * NB
    Lp = np.eye(n_pad)
    Lp[:n, :n] = L
    Kp = np.zeros((n_pad, m))
    Kp[:n] = Ks
    e = scale_exponent(sigma2, nugget)
    Vd = [np.zeros((n_pad, m)) for _ in range(S)]
    norms = np.zeros(m)
    for i0 in range(0, n_pad, NB):
        blk = slice(i0, i0 + NB)
        T = Kp[blk].copy()
        if i0 > 0:
            acc = sliced_product(Lp[blk, :i0], np.full(NB, e), [d[:i0] for d in Vd], S)
            T -= acc * 2.0 ** (2 * e)
        V = np.linalg.solve(Lp[blk, blk], T)               # the kernel multiplies by the stored inverse of L_ii (DMMA)
        norms += np.sum(V * V, axis=0)
        for t, d in enumerate(digits(V * 2.0 ** -e, S)):
            Vd[t][blk] = d
    return sigma2 + (nugget if include_nugget else 0.0) - norms


def trsm_variance_ltilde(L, Ks, sigma2, nugget, S, include_nugget=True):
    """The round-1 kernel (L~ form, one scale per row of L~): predictive variances (un-clipped)."""
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


def cholesky_i8(K, sigma2, nugget, S=8):
    """The arithmetic of csrc/chol.cu chol_i8_kernel: left-looking blocked factorisation (block 128) of K = sigma2 k(X, X) +
    nugget I whose history products sum_{k<j} L_ik L_jk^T come from S signed 7-bit digit planes of the blocks of L already
    computed (ONE scale 2^e per output, sqrt(sigma2 + nugget) <= 0.99 2^e; digit pairs t + u <= S + 1; exact integer sums),
    the diagonal blocks factored and the triangular solves done in FP64.  Returns L (n, n)."""
    n = K.shape[0]
    e = scale_exponent(sigma2, nugget)
    L = np.zeros_like(K)
    Ld = [np.zeros_like(K) for _ in range(S)]            # digit planes of the strictly lower blocks
    for j0 in range(0, n, NB):
        j1 = min(n, j0 + NB)
        panel = K[j0:, j0:j1].copy()
        if j0 > 0:
            acc = np.zeros_like(panel)
            for w in range(2, S + 2):
                part = np.zeros_like(panel)
                for t in range(1, w):
                    u = w - t
                    if t <= S and u <= S:
                        part += Ld[u - 1][j0:, :j0] @ Ld[t - 1][j0:j1, :j0].T
                acc += part * 2.0 ** (-BITS * w)
            panel -= acc * 2.0 ** (2 * e)
        Ljj = np.linalg.cholesky(panel[:j1 - j0])
        L[j0:j1, j0:j1] = Ljj
        if j1 < n:
            below = np.linalg.solve(Ljj, panel[j1 - j0:].T).T      # the kernel multiplies by the stored inverse of L_jj (DMMA)
            L[j1:, j0:j1] = below
            for t, dg in enumerate(digits(below * 2.0 ** -e, S)):
                Ld[t][j1:, j0:j1] = dg
    return L
