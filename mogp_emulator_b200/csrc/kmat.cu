// Kernel-matrix assembly for libmogp_b200 (replaces the scalar one-thread-per-element kernels
// sqexp_cov_batch_kernel / mat52_cov_batch_kernel, reference mogp_gpu/src/kernel.cu:55-65,251-261;
// arithmetic follows the CPU reference StationaryKernel.calc_r2 + calc_K, mogp_emulator/Kernel.py:476-485,
// 787-791, 878-882, and GaussianProcess.get_cov_matrix, GaussianProcess.py:517-558).
//
//   r2[i][j] = sum_d exp(theta_d) (x_i,d - x'_j,d)^2 ;  K = sigma2 * k(r2)
//
// The design matrices are kept transposed in HBM (XT: d x n_pad), so a 128-point tile is a TMA 2-D box
// (128 points x up-to-16 dims) landing in shared memory as [dim][point]: lanes read consecutive points
// (conflict-free) and the row operand is a broadcast.  Each thread owns an 8x8 sub-tile in registers
// and writes 128-byte row segments of K.
//   SYM   : lower 128x128 tiles of K(X,X) + nugget*I into the Cholesky workspace (identity in padding)
//   CROSS : test-major K*(Xs,X) into the predict workspace, with the fused posterior-mean partial
//           dot  sum_r K*[c][r] alpha[r]  (deterministic per-tile partials, no atomics).
#include "common.cuh"
#include "kernels.h"
#include "../../include/mogp_b200.h"

namespace mogp {

constexpr int DCH = 16;  // dims per TMA box / smem chunk (static smem stays under 48 KB)

template <int KT>
__device__ __forceinline__ double kfun(double r2) {
    if (KT == MOGP_KERNEL_SQEXP) {
        return exp(-0.5 * r2);
    } else {
        const double s = sqrt(5.0 * r2);
        return (1.0 + s + (5.0 / 3.0) * r2) * exp(-s);
    }
}

struct KmatParams {
    int64_t n, n_pad;       // training points (real / padded)
    int64_t rows_pad;       // CROSS: padded number of test points (m_pad)
    int d, dbox;            // input dims, dims per TMA box
    int hyper_stride;       // doubles per output in hyper (d + 2)
    const double* hyper;    // [count][d+2]
    double* out;            // SYM: matrix slab base; CROSS: workspace slab base
    int64_t out_stride;     // CROSS: rows per output in the workspace slab
    const double* alpha;    // CROSS: [count][n_pad] or null
    int64_t alpha_stride;
    double* part;           // CROSS: [count][n_tiles][rows_pad] or null
    int outs[MAXG];         // global output index handled by blockIdx.z (hyper/alpha rows)
    int store;              // CROSS: write the matrix (0 when only the mean is wanted)
    int add_nugget;         // SYM: add hyper[d+1] (the nugget) on the diagonal
    int* inf_flag;          // set to 1 when a squared distance is +inf (Kernel.py:482-483 raises FloatingPointError); may be null
};

template <int KT, int CROSS>
__global__ void __launch_bounds__(256, 1)
kmat_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const KmatParams p) {
    __shared__ __align__(128) double Xr[DCH * 128];
    __shared__ __align__(128) double Xc[DCH * 128];
    __shared__ double w_s[256];  // up to 256 dims
    __shared__ double al_s[128];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int I, J, o = 0;
    if (CROSS) {
        I = blockIdx.x;  // test tile
        J = blockIdx.y;  // train tile
        o = blockIdx.z;
    } else {
        const int id = blockIdx.x;
        o = blockIdx.y;
        I = (int)((sqrtf(8.0f * (float)id + 1.0f) - 1.0f) * 0.5f);
        while ((I + 1) * (I + 2) / 2 <= id) I++;
        while (I * (I + 1) / 2 > id) I--;
        J = id - I * (I + 1) / 2;
    }
    const double* hyp = p.hyper + (int64_t)p.outs[o] * p.hyper_stride;
    for (int i = tid; i < 256; i += 256) w_s[i] = (i < p.d) ? hyp[i] : 0.0;
    if (CROSS && tid < 128) al_s[tid] = p.alpha ? p.alpha[(int64_t)p.outs[o] * p.alpha_stride + (int64_t)J * 128 + tid] : 0.0;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    double r2[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) r2[a][b] = 0.0;

    const int nchunk = (p.d + p.dbox - 1) / p.dbox;
    uint32_t phase = 0;
    for (int ch = 0; ch < nchunk; ch++) {
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar, 2u * 128u * (uint32_t)p.dbox * 8u);
            tma_load_2d(Xr, &tmR, I * 128, ch * p.dbox, &bar);
            tma_load_2d(Xc, &tmC, J * 128, ch * p.dbox, &bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1u;
        const int dlim = min(p.dbox, p.d - ch * p.dbox);
        for (int dd = 0; dd < dlim; dd++) {
            const double wv = w_s[ch * p.dbox + dd];
            double xr[8], xc[8];
#pragma unroll
            for (int a = 0; a < 8; a++) xr[a] = Xr[dd * 128 + ty + 16 * a];
#pragma unroll
            for (int b = 0; b < 8; b++) xc[b] = Xc[dd * 128 + tx + 16 * b];
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    const double df = xr[a] - xc[b];
                    r2[a][b] = fma(wv, df * df, r2[a][b]);
                }
        }
        __syncthreads();  // single smem buffer: everyone done before the next box lands
    }

    if (p.inf_flag) {
        // the reference refuses infinite distances (calc_r2, Kernel.py:482-483); integer test of the bit pattern: the kernel
        // is bound by the FP64 pipe
        bool inf = false;
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 8; b++)
                inf |= (__double2hiint(r2[a][b]) == 0x7ff00000) && (__double2loint(r2[a][b]) == 0);
        if (inf) atomicOr(p.inf_flag, 1);
    }
    const double sigma2 = hyp[p.d];
    if (!CROSS) {
        const double nugget = p.add_nugget ? hyp[p.d + 1] : 0.0;
        const int64_t row_base = (int64_t)p.outs[o] * p.out_stride;
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int64_t row = (int64_t)I * 128 + ty + 16 * a;
            double* orow = p.out + (row_base + row) * p.n_pad;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int64_t col = (int64_t)J * 128 + tx + 16 * b;
                double v = sigma2 * kfun<KT>(r2[a][b]);
                if (row == col) v += nugget;
                if (row >= p.n || col >= p.n) v = (row == col) ? 1.0 : 0.0;
                orow[col] = v;
            }
        }
    } else {
#pragma unroll
        for (int a = 0; a < 8; a++) {
            const int64_t row = (int64_t)I * 128 + ty + 16 * a;  // test point
            double* orow = p.out + ((int64_t)o * p.out_stride + row) * p.n_pad;
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const int64_t col = (int64_t)J * 128 + tx + 16 * b;  // training point
                double v = sigma2 * kfun<KT>(r2[a][b]);
                if (col >= p.n) v = 0.0;
                if (p.store) orow[col] = v;
                s = fma(v, al_s[tx + 16 * b], s);
            }
            if (p.part) {
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (tx == 0) p.part[((int64_t)o * gridDim.y + J) * p.rows_pad + row] = s;
            }
        }
    }
}

struct OutList {
    int outs[MAXG];
};

__global__ void mean_reduce_kernel(const double* __restrict__ part, int n_tiles, int64_t m_pad, int64_t m,
                                   double* __restrict__ mean, int64_t mean_stride, const OutList ol) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = blockIdx.y;
    if (c >= m) return;
    double s = 0.0;
    for (int J = 0; J < n_tiles; J++) s += part[((int64_t)o * n_tiles + J) * m_pad + c];
    mean[(int64_t)ol.outs[o] * mean_stride + c] = s;
}



int kmat_dbox(int d) { return d < DCH ? d : DCH; }

// K + nugget*I (lower 128x128 tiles) of `count` outputs in one launch.  Output k uses hyper row outs[k] and is
// written at slab rows slab_idx[k]*n_pad.. of A_slab (slab_idx == nullptr: same as outs).
int kmat_sym(const CUtensorMap& tmXT, int kernel, int64_t n, int64_t n_pad, int d, const double* hyper, const int* outs,
             int count, int add_nugget, double* A_slab, int64_t slab_rows_per_output, cudaStream_t st, int* inf_flag) {
    if (count < 1 || count > MAXG) return 1;
    KmatParams p{};
    p.n = n; p.n_pad = n_pad; p.rows_pad = n_pad; p.d = d; p.dbox = kmat_dbox(d); p.hyper_stride = d + 2;
    p.hyper = hyper; p.out = A_slab; p.out_stride = slab_rows_per_output; p.add_nugget = add_nugget; p.inf_flag = inf_flag;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    const int T = (int)(n_pad / 128);
    const int tiles = T * (T + 1) / 2;
    dim3 grid((unsigned)tiles, (unsigned)count);
    if (kernel == MOGP_KERNEL_SQEXP)
        kmat_kernel<MOGP_KERNEL_SQEXP, 0><<<grid, 256, 0, st>>>(tmXT, tmXT, p);
    else
        kmat_kernel<MOGP_KERNEL_MATERN52, 0><<<grid, 256, 0, st>>>(tmXT, tmXT, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int kmat_cross(const CUtensorMap& tmXsT, const CUtensorMap& tmXT, int kernel, int64_t n, int64_t n_pad, int64_t m_pad,
               int d, const int* outs, int count, const double* hyper, double* W_slab, int64_t w_stride, int store,
               const double* alpha, int64_t alpha_stride, double* part, cudaStream_t st, int* inf_flag) {
    KmatParams p{};
    p.inf_flag = inf_flag;
    p.n = n; p.n_pad = n_pad; p.rows_pad = m_pad; p.d = d; p.dbox = kmat_dbox(d); p.hyper_stride = d + 2;
    p.hyper = hyper; p.out = W_slab; p.out_stride = w_stride; p.alpha = alpha; p.alpha_stride = alpha_stride;
    p.part = part; p.store = store;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    dim3 grid((unsigned)(m_pad / 128), (unsigned)(n_pad / 128), (unsigned)count);
    if (kernel == MOGP_KERNEL_SQEXP)
        kmat_kernel<MOGP_KERNEL_SQEXP, 1><<<grid, 256, 0, st>>>(tmXsT, tmXT, p);
    else
        kmat_kernel<MOGP_KERNEL_MATERN52, 1><<<grid, 256, 0, st>>>(tmXsT, tmXT, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int mean_reduce(const double* part, const int* outs, int count, int n_tiles, int64_t m_pad, int64_t m, double* mean,
                int64_t mean_stride, cudaStream_t st) {
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)count);
    OutList ol{};
    for (int i = 0; i < count; i++) ol.outs[i] = outs[i];
    mean_reduce_kernel<<<grid, 256, 0, st>>>(part, n_tiles, m_pad, m, mean, mean_stride, ol);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// Predictive derivatives d mean / d x*  (replaces cov_deriv_x_batch_gpu + cublasDgemv, reference
// mogp_gpu/src/kernel.cu:69-100, 264-302 and densegp_gpu.hpp:411-448):
//     dmu_c/dx*_q = sum_i alpha_i * sigma2 * k'(r2_ci) * 2 w_q (x*_cq - x_iq),   k' = dk/dr2
// One thread per test point streams the training points (tiles of 32 staged in shared memory, broadcast reads)
// and keeps 16 derivative components in registers (its own coordinates come through L1); more than 16 input
// dimensions take further passes.  No
// (m x n x d) derivative tensor is materialised (the reference allocates exactly that in work_mat_d).
// ------------------------------------------------------------------------------------------
constexpr int KD_TB = 32;    // training points per staged tile
constexpr int KD_Q = 16;     // derivative components per pass
constexpr int KD_MAXD = 256;   // == the library's limit on input dimensions

struct DerivParams {
    const double* XsT;     // [d][xs_stride] test points, transposed
    int64_t xs_stride;
    const double* XT;      // [d][n_pad]
    const double* alpha;   // [E][n_pad]
    const double* hyper;   // [E][d+2]
    double* out;           // [count][m][d]
    int64_t n, n_pad, m;
    int d;
    int outs[MAXG];
};

template <int KT>
__global__ void __launch_bounds__(128) kderiv_kernel(const DerivParams p) {
    extern __shared__ __align__(16) double kd_smem[];
    const int d = p.d;
    double* xt = kd_smem;              // [d][KD_TB]
    double* al = xt + d * KD_TB;       // [KD_TB]
    double* w = al + KD_TB;            // [d]
    const int tid = threadIdx.x;
    const int o = p.outs[blockIdx.y];
    const int64_t c = (int64_t)blockIdx.x * 128 + tid;
    const double* hyp = p.hyper + (int64_t)o * (d + 2);
    for (int i = tid; i < d; i += 128) w[i] = hyp[i];
    // this thread's test point: coalesced, L1-resident reads of XsT (zero-padded up to xs_stride)
    const double* xs = p.XsT + ((c < p.xs_stride) ? c : 0);
    const int64_t xss = p.xs_stride;
    const double sigma2 = hyp[d];
    __syncthreads();
    for (int d0 = 0; d0 < d; d0 += KD_Q) {
        double g[KD_Q];
#pragma unroll
        for (int q = 0; q < KD_Q; q++) g[q] = 0.0;
        for (int64_t i0 = 0; i0 < p.n; i0 += KD_TB) {
            __syncthreads();
            for (int idx = tid; idx < d * KD_TB; idx += 128) {
                const int dd = idx / KD_TB, ii = idx - dd * KD_TB;
                xt[idx] = (i0 + ii < p.n) ? p.XT[(int64_t)dd * p.n_pad + i0 + ii] : 0.0;
            }
            if (tid < KD_TB) al[tid] = (i0 + tid < p.n) ? p.alpha[(int64_t)o * p.n_pad + i0 + tid] : 0.0;
            __syncthreads();
#pragma unroll 2
            for (int ii = 0; ii < KD_TB; ii++) {
                double r2 = 0.0;
                for (int dd = 0; dd < d; dd++) {
                    const double df = __ldg(xs + dd * xss) - xt[dd * KD_TB + ii];
                    r2 = fma(w[dd], df * df, r2);
                }
                double dk;
                if (KT == MOGP_KERNEL_SQEXP) {
                    dk = -0.5 * exp(-0.5 * r2);
                } else {
                    const double sq = sqrt(5.0 * r2);
                    dk = -(5.0 / 6.0) * (1.0 + sq) * exp(-sq);
                }
                const double kp = sigma2 * dk * al[ii];   // al is zero past n
#pragma unroll
                for (int q = 0; q < KD_Q; q++)
                    if (d0 + q < d) g[q] = fma(kp, __ldg(xs + (d0 + q) * xss) - xt[(d0 + q) * KD_TB + ii], g[q]);
            }
        }
        if (c < p.m) {
            double* orow = p.out + ((int64_t)blockIdx.y * p.m + c) * d;
#pragma unroll
            for (int q = 0; q < KD_Q; q++)
                if (d0 + q < d) orow[d0 + q] = 2.0 * w[d0 + q] * g[q];
        }
    }
}

int kderiv_max_dims() { return KD_MAXD; }

int kmat_init() {
    const int maxs = (int)(sizeof(double) * ((size_t)KD_MAXD * KD_TB + KD_TB + KD_MAXD));
    if (cudaFuncSetAttribute(kderiv_kernel<MOGP_KERNEL_SQEXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs) != cudaSuccess ||
        cudaFuncSetAttribute(kderiv_kernel<MOGP_KERNEL_MATERN52>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs) != cudaSuccess)
        return 1;
    return 0;
}

int kmat_deriv(int kernel, const double* XsT, int64_t xs_stride, const double* XT, int64_t n, int64_t n_pad, int64_t m,
               int d, const int* outs, int count, const double* hyper, const double* alpha, double* out, cudaStream_t st) {
    if (count < 1 || count > MAXG || d > KD_MAXD) return 1;
    const size_t smem = sizeof(double) * ((size_t)d * KD_TB + KD_TB + d);
    DerivParams p{};
    p.XsT = XsT; p.xs_stride = xs_stride; p.XT = XT; p.alpha = alpha; p.hyper = hyper; p.out = out;
    p.n = n; p.n_pad = n_pad; p.m = m; p.d = d;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    dim3 grid((unsigned)((m + 127) / 128), (unsigned)count);
    if (kernel == MOGP_KERNEL_SQEXP)
        kderiv_kernel<MOGP_KERNEL_SQEXP><<<grid, 128, smem, st>>>(p);
    else
        kderiv_kernel<MOGP_KERNEL_MATERN52><<<grid, 128, smem, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mogp
