"""CPU study behind csrc/trsm_i8.cu: the predict TRSM's GEMM part on INT8 tensor cores (Ozaki-style slicing).

V_i = inv(L_ii) (K*_i - sum_{j<i} L_ij V_j) with the sum evaluated from s signed 7-bit slices of fixed-point L rows
(one exponent per row of L) and V columns (one global exponent, |V| <= sqrt(sigma2 + nugget) because the predictive
variance is non-negative), keeping the slice pairs with t + u <= s + 1 -- every int8 x int8 product and int32 sum is
exact, so numpy integer matmuls reproduce what tcgen05.mma kind::i8 would compute.  Reports the error this puts on the
predictive variance next to the FP64 result (tolerance of the parity tests: rtol 1e-4, atol 1e-4 * nugget).
"""
import sys, os
import numpy as np
import scipy.linalg
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import gp_oracle as orc

BITS = 7


def slices(x, scale_exp, s):
    """x * 2^-scale_exp in (-1, 1) -> s integer slices d_t with x ~= 2^scale_exp * sum_t d_t 2^(-BITS t), |d_t| <= 64."""
    r = np.ldexp(x, -scale_exp)
    out = []
    for t in range(1, s + 1):
        d = np.rint(np.ldexp(r, BITS * t))          # round to nearest: signed digits
        out.append(d.astype(np.int64))
        r = r - np.ldexp(d, -BITS * t)
    return out


def sliced_gemm(A, B, s, b_exp):
    """A (p, k) row-scaled, B (k, q) with the global exponent b_exp."""
    a_exp = np.frexp(np.abs(A).max(axis=1, keepdims=True) + 1e-300)[1]
    As = slices(A, a_exp, s)
    Bs = slices(B, b_exp, s)
    acc = np.zeros((A.shape[0], B.shape[1]))
    n_mma = 0
    for t in range(s):
        for u in range(s):
            if t + u + 2 <= s + 1:
                acc += np.ldexp((As[t] @ Bs[u]).astype(np.float64), -BITS * (t + u + 2))
                n_mma += 1
    return np.ldexp(acc, a_exp) * 2.0 ** b_exp, n_mma


def trsm_var(L, Ks, sigma2, nugget, s=None, nb=128):
    n, m = Ks.shape
    V = np.zeros_like(Ks)
    b_exp = int(np.frexp(np.sqrt(sigma2 + nugget))[1])
    n_mma = 0
    for i0 in range(0, n, nb):
        i1 = min(n, i0 + nb)
        rhs = Ks[i0:i1].copy()
        if i0 > 0:
            if s is None:
                rhs -= L[i0:i1, :i0] @ V[:i0]
            else:
                upd, n_mma = sliced_gemm(L[i0:i1, :i0], V[:i0], s, b_exp)
                rhs -= upd
        V[i0:i1] = scipy.linalg.solve_triangular(L[i0:i1, i0:i1], rhs, lower=True)
    return sigma2 + nugget - np.sum(V * V, axis=0), n_mma


def trsm_var_ltilde(L, Ks, sigma2, nugget, s, nb=128):
    """The formulation csrc/trsm_i8.cu runs: V_i = inv(L_ii) K*_i - sum_j (inv(L_ii) L_ij) V_j with the rows of
    L~ = blockdiag(L_ii)^-1 L and the solved V sliced (no FP64 diagonal solve inside the integer kernel)."""
    n, m = Ks.shape
    V = np.zeros_like(Ks)
    b_exp = int(np.frexp(np.sqrt(sigma2 + nugget))[1])
    for i0 in range(0, n, nb):
        i1 = min(n, i0 + nb)
        Dinv = np.linalg.inv(L[i0:i1, i0:i1])
        V[i0:i1] = Dinv @ Ks[i0:i1]
        if i0 > 0:
            upd, _ = sliced_gemm(Dinv @ L[i0:i1, :i0], V[:i0], s, b_exp)
            V[i0:i1] -= upd
    return sigma2 + nugget - np.sum(V * V, axis=0)


def main():
    n, d, m = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 10, 256
    for theta_corr, label in ((1.0, "theta_corr=+1 (benchmark setting)"), (-1.0, "theta_corr=-1 (ill-conditioned)")):
        X, Y, Xs = orc.make_workload(n, d, 1, m, seed=2)
        theta = np.append(np.full(d, theta_corr), 0.0)
        gp = orc.OracleGP(X, Y[0], nugget=1e-6, priors="weak").fit(theta)
        Ks = gp.get_cov_matrix(Xs)
        K = gp.get_K_matrix() + 1e-6 * np.eye(n)
        ref, _ = trsm_var(gp.L, Ks, 1.0, 1e-6)
        exact = 1.0 + 1e-6 - np.sum(Ks * np.linalg.solve(K, Ks), axis=0)      # a second FP64 algorithm, for scale
        print("%s: n=%d cond(K)=%.1e  |var| in [%.2e, %.2e]  FP64 blocked vs LU solve: max abs diff %.1e"
              % (label, n, np.linalg.cond(K), ref.min(), ref.max(), np.abs(ref - exact).max()))
        for s in (5, 6, 7, 8):
            v, n_mma = trsm_var(gp.L, Ks, 1.0, 1e-6, s=s)
            err = np.abs(v - ref)
            ok = np.all(err <= 1e-4 * np.abs(ref) + 1e-4 * 1e-6)
            print("   %d slices (%2d int8 MMAs per block): max abs var error %.2e, max rel %.2e -> parity %s"
                  % (s, n_mma, err.max(), (err / np.maximum(np.abs(ref), 1e-300)).max(), "OK" if ok else "BROKEN"))
            err2 = np.abs(trsm_var_ltilde(gp.L, Ks, 1.0, 1e-6, s) - ref)
            tol = 1e-4 * np.abs(ref) + 1e-4 * 1e-6
            print("      L~ formulation (what the kernel runs): max abs %.2e, worst error / tolerance %.3f (rows-of-L form: %.3f)"
                  % (err2.max(), (err2 / tol).max(), (err / tol).max()))


if __name__ == "__main__":
    main()
