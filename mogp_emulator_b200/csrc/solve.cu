// Single right-hand-side triangular solves for libmogp_b200:  z = L^-1 y,  alpha = L^-T z,  quad = z^T z.
// Replaces cusolverDnDpotrs for invQt (reference mogp_gpu/src/densegp_gpu.hpp:585-591) and the cublasDdot
// for y^T alpha (:604-611); CPU semantics: ChoInv.solve = cho_solve((L, True), y), linalg/cholesky.py:22-42,
// used at GaussianProcess.py:666-672.
//
// Blocked substitution over 128-row blocks: the off-diagonal part is a streaming GEMV over L (row-major,
// 16-byte vector loads), the diagonal block is applied through its precomputed inverse (Dinv, produced by
// the Cholesky panel kernel), so there is no scalar dependency chain.  One CTA per right-hand side:
// independent outputs run concurrently on their own streams.
#include "common.cuh"
#include "kernels.h"

namespace mogp {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(512, 1)
solve_alpha_kernel(const double* __restrict__ A, int64_t ld, int T, const double* __restrict__ Dinv,
                   const double* __restrict__ y, double* __restrict__ z_out, double* __restrict__ alpha_out,
                   double* __restrict__ quad, const int* __restrict__ info) {
    extern __shared__ __align__(16) double sv[];  // v[n_pad] | acc[128] | red[4*128]
    if (*info != 0) return;
    const int n_pad = T * NB;
    double* v = sv;
    double* acc = sv + n_pad;
    double* red = acc + NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < n_pad; i += 512) v[i] = y[i];
    __syncthreads();

    // ---- forward substitution ----
    for (int i = 0; i < T; i++) {
        const int c_end = i * NB;
        {
            double s[8];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) s[rr] = 0.0;
            const double* rowp = A + (int64_t)(i * NB + warp * 8) * ld;
            for (int c = lane * 2; c < c_end; c += 64) {
                const double2 vv = *reinterpret_cast<const double2*>(v + c);
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {
                    const double2 l = *reinterpret_cast<const double2*>(rowp + (int64_t)rr * ld + c);
                    s[rr] = fma(l.x, vv.x, s[rr]);
                    s[rr] = fma(l.y, vv.y, s[rr]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double t = warp_sum(s[rr]);
                if (lane == 0) acc[warp * 8 + rr] = v[c_end + warp * 8 + rr] - t;
            }
        }
        __syncthreads();
        {
            const double* Db = Dinv + (int64_t)i * NB * NB;
            const double2 a0 = *reinterpret_cast<const double2*>(acc + lane * 4);
            const double2 a1 = *reinterpret_cast<const double2*>(acc + lane * 4 + 2);
            double res[8];
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double* dr = Db + (int64_t)(warp * 8 + rr) * NB + lane * 4;
                const double2 d0 = *reinterpret_cast<const double2*>(dr);
                const double2 d1 = *reinterpret_cast<const double2*>(dr + 2);
                double t = d0.x * a0.x;
                t = fma(d0.y, a0.y, t);
                t = fma(d1.x, a1.x, t);
                t = fma(d1.y, a1.y, t);
                res[rr] = warp_sum(t);
            }
            __syncthreads();  // all reads of acc done before v (and later acc) change
            if (lane == 0) {
#pragma unroll
                for (int rr = 0; rr < 8; rr++) v[c_end + warp * 8 + rr] = res[rr];
            }
        }
        __syncthreads();
    }
    {
        double q = 0.0;
        for (int i = tid; i < n_pad; i += 512) {
            q = fma(v[i], v[i], q);
            z_out[i] = v[i];
        }
        q = warp_sum(q);
        if (lane == 0) red[warp] = q;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 16; w++) t += red[w];
            *quad = t;
        }
        __syncthreads();
    }

    // ---- backward substitution ----
    const int c = tid & 127, q4 = tid >> 7;
    for (int i = T - 1; i >= 0; i--) {
        {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double* colp = A + (int64_t)i * NB + c;
            int r = (i + 1) * NB + q4;
            for (; r + 12 < n_pad; r += 16) {
                s0 = fma(colp[(int64_t)r * ld], v[r], s0);
                s1 = fma(colp[(int64_t)(r + 4) * ld], v[r + 4], s1);
                s2 = fma(colp[(int64_t)(r + 8) * ld], v[r + 8], s2);
                s3 = fma(colp[(int64_t)(r + 12) * ld], v[r + 12], s3);
            }
            for (; r < n_pad; r += 4) s0 = fma(colp[(int64_t)r * ld], v[r], s0);
            red[q4 * NB + c] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        if (tid < NB) acc[tid] = v[i * NB + tid] - ((red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]));
        __syncthreads();
        {
            const double* Db = Dinv + (int64_t)i * NB * NB;
            double s = 0.0;
            for (int r = q4; r < NB; r += 4)
                if (r >= c) s = fma(Db[r * NB + c], acc[r], s);
            red[q4 * NB + c] = s;
        }
        __syncthreads();
        if (tid < NB) v[i * NB + tid] = (red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]);
        __syncthreads();
    }
    for (int i = tid; i < n_pad; i += 512) alpha_out[i] = v[i];
}

int solve_init() {
    static bool done = false;
    if (done) return 0;
    if (cudaFuncSetAttribute(solve_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return 1;
    done = true;
    return 0;
}

int solve_alpha(const double* A, int64_t n_pad, const double* Dinv, const double* y, double* z, double* alpha,
                double* quad, const int* info, cudaStream_t st) {
    const size_t smem = (size_t)(n_pad + NB + 4 * NB) * sizeof(double);
    if (smem > 200 * 1024) return 2;  // n_pad > ~24k: needs the multi-CTA solver
    solve_alpha_kernel<<<1, 512, smem, st>>>(A, n_pad, (int)(n_pad / NB), Dinv, y, z, alpha, quad, info);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mogp
