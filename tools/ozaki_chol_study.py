"""CPU study for a round-2 kernel: the Cholesky's trailing updates on INT8 tensor cores (Ozaki-style slicing).

Left-looking blocked factorisation (block 128) as chol_dataflow_kernel runs it,
    L_jj = chol(A_jj - sum_{k<j} L_jk L_jk^T),     L_ij = (A_ij - sum_{k<j} L_ik L_jk^T) L_jj^-T,
with the sums evaluated from s signed 7-bit digits of the rows of L (one power-of-two scale per row; the rows of a Cholesky
factor are bounded by sqrt(K_rr)), keeping the digit pairs with t + u <= s + 1; the diagonal-block factorisation and the
triangular solve stay FP64.  Digits are held in float64 and multiplied with BLAS: every product and partial sum is an
integer below 2^53, so this is exactly what tcgen05.mma kind::i8 with s32 accumulation computes.  Reports what the
scheme does to the quantities the parity tests check: L, log det K, y^T K^-1 y (log-posterior), the posterior mean and
variance.  Run: python tools/ozaki_chol_study.py [n]
"""
import os
import sys

import numpy as np
import scipy.linalg

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import gp_oracle as orc

BITS = 7


def digits(x, s):
    y = np.clip(x, -0.99, 0.99).copy()
    out = []
    for _ in range(s):
        y = y * 2.0 ** BITS
        d = np.rint(y)
        y = y - d
        out.append(d)
    return out


def sliced_abt(A, B, s):
    """A B^T from s digits per operand row (row scales), digit pairs t + u <= s + 1."""
    ea = np.frexp(np.maximum(np.abs(A).max(axis=1), 1e-300))[1] + 1
    eb = np.frexp(np.maximum(np.abs(B).max(axis=1), 1e-300))[1] + 1
    Ad = digits(A * 2.0 ** (-ea)[:, None], s)
    Bd = digits(B * 2.0 ** (-eb)[:, None], s)
    acc = np.zeros((A.shape[0], B.shape[0]))
    for t in range(s):
        for u in range(s):
            if t + u + 2 <= s + 1:
                acc += (Ad[t] @ Bd[u].T) * 2.0 ** (-BITS * (t + u + 2))
    return acc * 2.0 ** ea[:, None] * 2.0 ** eb[None, :]


def blocked_cholesky(K, s=None, nb=128):
    n = K.shape[0]
    L = np.zeros_like(K)
    for j0 in range(0, n, nb):
        j1 = min(n, j0 + nb)
        if j0 > 0:
            upd = L[j0:, :j0] @ L[j0:j1, :j0].T if s is None else sliced_abt(L[j0:, :j0], L[j0:j1, :j0], s)
        else:
            upd = 0.0
        panel = K[j0:, j0:j1] - upd
        Ljj = np.linalg.cholesky(panel[:j1 - j0])
        L[j0:j1, j0:j1] = Ljj
        if j1 < n:
            L[j1:, j0:j1] = scipy.linalg.solve_triangular(Ljj, panel[j1 - j0:].T, lower=True).T
    return L


def posterior(L, y, Ks, sigma2, nugget):
    alpha = scipy.linalg.cho_solve((L, True), y)
    V = scipy.linalg.solve_triangular(L, Ks, lower=True)
    return alpha, 2.0 * np.sum(np.log(np.diag(L))), float(y @ alpha), Ks.T @ alpha, sigma2 + nugget - np.sum(V * V, axis=0)


def main():
    n, d, m = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 10, 256
    nugget = 1e-6
    for theta_corr, label in ((1.0, "theta_corr=+1 (benchmark setting)"), (-1.0, "theta_corr=-1 (ill-conditioned)")):
        X, Y, Xs = orc.make_workload(n, d, 1, m, seed=2)
        y = Y[0]
        theta = np.append(np.full(d, theta_corr), 0.0)
        K = np.exp(theta[d]) * orc.kernel_f(X, X, theta[:d]) + nugget * np.eye(n)
        Ks = np.exp(theta[d]) * orc.kernel_f(X, Xs, theta[:d])
        Lref = np.linalg.cholesky(K)
        ref = posterior(Lref, y, Ks, 1.0, nugget)
        Lb = blocked_cholesky(K)
        blk = posterior(Lb, y, Ks, 1.0, nugget)
        print("%s: n=%d cond(K)=%.1e;  FP64 blocked vs LAPACK: |dL|/|L| %.1e, logdet %.1e, quad rel %.1e, mean rel %.1e, var abs %.1e"
              % (label, n, np.linalg.cond(K), np.linalg.norm(Lb - Lref) / np.linalg.norm(Lref), abs(blk[1] - ref[1]),
                 abs(blk[2] - ref[2]) / abs(ref[2]), np.abs(blk[3] - ref[3]).max() / np.abs(ref[3]).max(), np.abs(blk[4] - ref[4]).max()))
        for s in (6, 7, 8):
            try:
                Ls = blocked_cholesky(K, s)
            except np.linalg.LinAlgError:
                print("   %d digits: factorisation FAILED (a diagonal block lost positive definiteness)" % s)
                continue
            got = posterior(Ls, y, Ks, 1.0, nugget)
            tol_var = 1e-4 * np.abs(ref[4]) + 1e-4 * nugget
            print("   %d digits (%2d MMAs per K step): |dL|/|L| %.1e, logdet abs %.1e, quad rel %.1e (bar 1e-9 x cond-scaling), "
                  "mean rel %.1e (bar 1e-6), var worst err/tol %.3f"
                  % (s, s * (s + 1) // 2, np.linalg.norm(Ls - Lref) / np.linalg.norm(Lref), abs(got[1] - ref[1]),
                     abs(got[2] - ref[2]) / abs(ref[2]), np.abs(got[3] - ref[3]).max() / np.abs(ref[3]).max(),
                     (np.abs(got[4] - ref[4]) / tol_var).max()))


if __name__ == "__main__":
    main()
