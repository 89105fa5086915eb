// Gradient of the negative log-posterior (data part) for libmogp_b200.
//
// Replaces the reference GPU path: 12 deriv_theta kernels that materialise (D+1) n x n planes plus three
// cublasDgemv per parameter on top of an explicit inverse (mogp_gpu/src/kernel.cu:127-141,334-348;
// densegp_gpu.hpp:663-770).  Formula of the CPU reference for zero mean (GaussianProcess.logpost_deriv,
// GaussianProcess.py:711-782; kernel_deriv Kernel.py:133-173, 487-530, 793-814, 884-906; logdet_deriv
// linalg_utils.py:170-198):
//     dL/dtheta_i = 0.5 * ( tr(K^-1 dK_i) - alpha^T dK_i alpha ) = 0.5 * sum_jk G_jk (dK_i)_jk,   G = K^-1 - alpha alpha^T
//     corr i : (dK_i)_jk = sigma2 * k'(r2_jk) * exp(theta_i) (x_ji - x_ki)^2
//     cov    : dK = sigma2 * k(r2)            nugget (fitted): dK = nugget * I
//
// Two device steps, no derivative planes in memory:
//   1. Wt = (L^-1)^T by the predict TRSM kernel on an identity right-hand side (predict.cu, tri_rhs mode)
//   2. grad_tile_kernel: K^-1 tile = Wt_J . Wt_K^T as a DMMA GEMM (TMA-fed ring, same TN form as the SYRK),
//      and in the epilogue, while the tile is still in registers, every parameter's  G o dK_i  reduction.
//      Per-tile partial sums are written out and reduced in a fixed order (deterministic).
#include "../../include/mogp_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace mogp {

constexpr int G_BM = 128, G_BN = 64, G_NS = 4, G_NCW = 4;
constexpr int G_A_BYTES = G_BM * KC * 8, G_B_BYTES = G_BN * KC * 8, G_STAGE = G_A_BYTES + G_B_BYTES;
constexpr int G_MAXD = 64;  // X tiles of all dims are parked in the (idle) ring during the epilogue
constexpr int G_MAXM = 4;   // mean-function vectors u_q (G -= sum_q u_q u_q^T) staged in shared memory, see mogp_set_mean_vectors
constexpr int G_MAXM_TOTAL = 32;   // vectors per output in all: those beyond G_MAXM are read through L1 in the epilogue
constexpr int G_SMEM = G_NS * G_STAGE + 2 * G_NS * 8 + 128 + (G_BM + G_BN) * 8 * (1 + G_MAXM) + 32 * 8 * G_NCW + G_MAXD * 8;
constexpr int G_DCH = 8;    // parameters reduced per epilogue pass

struct GradParams {
    const double* XT;     // [d][n_pad]
    const double* alpha;  // [n_pad] of this output
    const double* hyper;  // [d+2] of this output
    double* partial;      // [tiles][d+2]
    int64_t n, n_pad;
    int d, kernel;
    int row_base;         // first row of Wt inside its tensor map (0)
    const double* U;      // [n_u][u_stride] mean-function vectors of this output (or null)
    int n_u;
    int64_t u_stride;
    // MODE 1 (full predictive covariance): C (lower 128x64 tiles, row stride ldc) -= W_I . W_J^T over k_chunks chunks
    double* C;
    int64_t ldc;
    int k_chunks;
};

template <int KT>
__device__ __forceinline__ void kval_and_deriv(double r2, double& k, double& dk) {
    if (KT == MOGP_KERNEL_SQEXP) {
        k = exp(-0.5 * r2);
        dk = -0.5 * k;
    } else {
        const double s = sqrt(5.0 * r2);
        const double e = exp(-s);
        k = (1.0 + s + (5.0 / 3.0) * r2) * e;
        dk = -(5.0 / 6.0) * (1.0 + s) * e;
    }
}

template <int KT, int MODE>
__global__ void __launch_bounds__((G_NCW + 4) * 32, 2)
grad_tile_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GradParams p) {
    extern __shared__ __align__(128) unsigned char grad_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(grad_smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + G_NS * G_STAGE);
    uint64_t* empty = full + G_NS;
    double* al_r = reinterpret_cast<double*>(empty + G_NS);  // [128]
    double* al_c = al_r + G_BM;                               // [64]
    double* u_rc = al_c + G_BN;                               // [G_MAXM][128 + 64]
    double* red = u_rc + G_MAXM * (G_BM + G_BN);              // [G_NCW][32]
    double* w_s = red + 32 * G_NCW;                           // [G_MAXD] exp(theta_i)
    double* Xs = reinterpret_cast<double*>(base);            // epilogue: [d][192] (rows then cols), aliases the ring

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int id = blockIdx.x;
    int I = (int)((sqrtf(4.0f * (float)id + 1.0f) - 1.0f) * 0.5f);
    while ((I + 1) * (I + 2) <= id) I++;
    while (I * (I + 1) > id) I--;
    const int J2 = id - I * (I + 1);
    const int row0 = I * G_BM, col0 = J2 * G_BN;
    const int T = (int)(p.n_pad / NB);
    // MODE 0: contraction over r >= 128*I (Wt is upper triangular); MODE 1: over the whole row of W
    const int nchunk = (MODE == 0) ? (T - I) * (NB / KC) : p.k_chunks;
    const int kout0 = (MODE == 0) ? I * (NB / 8) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], G_NCW);
        }
        fence_mbar_init();
    }
    if (MODE == 0) {
        for (int i = threadIdx.x; i < G_BM + G_BN; i += blockDim.x)
            al_r[i] = (i < G_BM) ? p.alpha[row0 + i] : p.alpha[col0 + i - G_BM];
        const int n_u_sm = p.n_u < G_MAXM ? p.n_u : G_MAXM;
        for (int i = threadIdx.x; i < n_u_sm * (G_BM + G_BN); i += blockDim.x) {
            const int q = i / (G_BM + G_BN), k = i - q * (G_BM + G_BN);
            u_rc[i] = p.U[(int64_t)q * p.u_stride + ((k < G_BM) ? row0 + k : col0 + k - G_BM)];
        }
        for (int i = threadIdx.x; i < G_MAXD; i += blockDim.x) w_s[i] = (i < p.d) ? p.hyper[i] : 0.0;
    }
    __syncthreads();

    if (warp >= G_NCW) {
        reg_dealloc<24>();
        if (warp == G_NCW && lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmB);
            PipeState<G_NS> ps;
            for (int c = 0; c < nchunk; c++) {
                mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                unsigned char* st = base + ps.stage * G_STAGE;
                mbar_arrive_expect_tx(&full[ps.stage], G_STAGE);
                tma_load_3d(st, &tmA, 0, p.row_base + row0, kout0 + c * (KC / 8), &full[ps.stage]);
                tma_load_3d(st + G_A_BYTES, &tmB, 0, p.row_base + col0, kout0 + c * (KC / 8), &full[ps.stage]);
                ps.advance();
            }
        }
        return;
    }
    reg_alloc<232>();
    const int g = lane >> 2, t = lane & 3;
    const int arow0 = warp * 32;
    double acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.0;
    PipeState<G_NS> ps;
    for (int c = 0; c < nchunk; c++) {
        mbar_wait(&full[ps.stage], ps.phase);
        const double* As = reinterpret_cast<const double*>(base + ps.stage * G_STAGE);
        const double* Bs = reinterpret_cast<const double*>(base + ps.stage * G_STAGE + G_A_BYTES);
        mma_stage<2, 8, KC>(acc, As, G_BM, arow0, Bs, G_BN, 0, g, t);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[ps.stage]);
        ps.advance();
    }

    if (MODE == 1) {
        // full predictive covariance: C_tile -= W_I . W_J^T   (GaussianProcess.py:899-911)
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                double* rowp = p.C + (int64_t)(row0 + arow0 + mt * 16 + g + h * 8) * p.ldc + col0 + 2 * t;
#pragma unroll
                for (int nt = 0; nt < 8; nt++) {
                    double2 v = *reinterpret_cast<const double2*>(rowp + nt * 8);
                    v.x -= acc[mt][nt][2 * h];
                    v.y -= acc[mt][nt][2 * h + 1];
                    *reinterpret_cast<double2*>(rowp + nt * 8) = v;
                }
            }
        return;
    }
    // ---- epilogue: all TMA traffic has landed and been consumed; park the X tiles in the ring ----
    named_bar_sync(1, G_NCW * 32);
    const int d = p.d;
    for (int i = threadIdx.x; i < d * (G_BM + G_BN); i += G_NCW * 32) {
        const int dd = i / (G_BM + G_BN), q = i - dd * (G_BM + G_BN);
        const int64_t pt = (q < G_BM) ? (row0 + q) : (col0 + q - G_BM);
        Xs[i] = p.XT[(int64_t)dd * p.n_pad + pt];
    }
    named_bar_sync(1, G_NCW * 32);

    const double sigma2 = p.hyper[d];
    double s_cov = 0.0, s_tr = 0.0;
    // pass A: G_jk and the kernel terms; acc becomes c_jk = weight * G_jk * sigma2 * k'(r2_jk)
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int rl = arow0 + mt * 16 + g + ((e >> 1) << 3);
                const int cl = nt * 8 + 2 * t + (e & 1);
                const int64_t j = row0 + rl, k = col0 + cl;
                double w = (k < j) ? 2.0 : ((k == j) ? 1.0 : 0.0);
                if (j >= p.n || k >= p.n) w = 0.0;
                double r2 = 0.0;
                for (int dd = 0; dd < d; dd++) {
                    const double df = Xs[dd * (G_BM + G_BN) + rl] - Xs[dd * (G_BM + G_BN) + G_BM + cl];
                    r2 = fma(w_s[dd], df * df, r2);
                }
                double kv, dk;
                kval_and_deriv<KT>(r2, kv, dk);
                const double kinv = acc[mt][nt][e];
                double G = kinv - al_r[rl] * al_c[cl];
                for (int q = 0; q < p.n_u && q < G_MAXM; q++)    // analytic mean function: K^-1 -> K^-1 - K^-1 H A^-1 H^T K^-1
                    G = fma(-u_rc[q * (G_BM + G_BN) + rl], u_rc[q * (G_BM + G_BN) + G_BM + cl], G);
                for (int q = G_MAXM; q < p.n_u; q++)             // wide design matrices: the rest straight from global memory
                    G = fma(-__ldg(p.U + (int64_t)q * p.u_stride + j), __ldg(p.U + (int64_t)q * p.u_stride + k), G);
                s_cov = fma(w * G, sigma2 * kv, s_cov);
                if (k == j && j < p.n) s_tr += kinv;
                acc[mt][nt][e] = w * G * sigma2 * dk;
            }
    // block reduction helper: warp shuffle, then across the 4 consumer warps through smem
    auto block_sum_store = [&](double v, int slot) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp * 32 + slot] = v;
    };
    block_sum_store(s_cov, 0);
    block_sum_store(s_tr, 1);
    named_bar_sync(1, G_NCW * 32);
    if (threadIdx.x < 2) {
        const double v = (red[threadIdx.x] + red[32 + threadIdx.x]) + (red[64 + threadIdx.x] + red[96 + threadIdx.x]);
        p.partial[(int64_t)id * (d + 2) + d + threadIdx.x] = v;   // [d] = cov term, [d+1] = trace
    }
    named_bar_sync(1, G_NCW * 32);
    // pass B: correlation-length parameters, G_DCH at a time
    for (int d0 = 0; d0 < d; d0 += G_DCH) {
        double s[G_DCH];
#pragma unroll
        for (int q = 0; q < G_DCH; q++) s[q] = 0.0;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int rl = arow0 + mt * 16 + g + ((e >> 1) << 3);
                    const int cl = nt * 8 + 2 * t + (e & 1);
                    const double c = acc[mt][nt][e];
#pragma unroll
                    for (int q = 0; q < G_DCH; q++) {
                        if (d0 + q < d) {
                            const double df = Xs[(d0 + q) * (G_BM + G_BN) + rl] - Xs[(d0 + q) * (G_BM + G_BN) + G_BM + cl];
                            s[q] = fma(c, df * df, s[q]);
                        }
                    }
                }
#pragma unroll
        for (int q = 0; q < G_DCH; q++) block_sum_store(s[q], q);
        named_bar_sync(1, G_NCW * 32);
        if (threadIdx.x < G_DCH && d0 + threadIdx.x < d) {
            const int q = threadIdx.x;
            const double v = (red[q] + red[32 + q]) + (red[64 + q] + red[96 + q]);
            p.partial[(int64_t)id * (d + 2) + d0 + q] = v * w_s[d0 + q];   // exp(theta_i) factor of dr2/dtheta_i
        }
        named_bar_sync(1, G_NCW * 32);
    }
}

// grad[i] = 0.5 * sum_tiles partial[tile][i]; fitted nugget: 0.5 * nugget * (tr K^-1 - alpha^T alpha - sum_q u_q^T u_q)
__global__ void grad_reduce_kernel(const double* __restrict__ partial, int tiles, int d, const double* __restrict__ alpha,
                                   int64_t n, const double* __restrict__ hyper, int fit_nugget, double* __restrict__ grad,
                                   const double* __restrict__ U, int n_u, int64_t u_stride) {
    __shared__ double sh[256];
    const int i = blockIdx.x;  // 0..d+1
    double s = 0.0;
    for (int tI = threadIdx.x; tI < tiles; tI += 256) s += partial[(int64_t)tI * (d + 2) + i];
    double aa = 0.0;
    if (i == d + 1)
        for (int64_t r = threadIdx.x; r < n; r += 256) {
            aa = fma(alpha[r], alpha[r], aa);
            for (int q = 0; q < n_u; q++) aa = fma(U[q * u_stride + r], U[q * u_stride + r], aa);
        }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    const double tot = sh[0];
    __syncthreads();
    sh[threadIdx.x] = aa;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (i <= d) grad[i] = 0.5 * tot;
        else if (fit_nugget) grad[i] = 0.5 * hyper[d + 1] * (tot - sh[0]);
    }
}

__global__ void set_identity_kernel(double* __restrict__ W, int64_t n_pad) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n_pad * n_pad) W[idx] = ((idx / n_pad) == (idx % n_pad)) ? 1.0 : 0.0;
}

// out[c] = 1 / sum_r W[c][r]^2 for the first n rows of a row-major n_pad x n_pad matrix: with W = (L^-1)^T (row c = column c of
// L^-1) that is 1 / (K^-1)_cc, the leave-one-out predictive variance at training point c
__global__ void row_inv_sumsq_kernel(const double* __restrict__ W, int64_t n_pad, int64_t n, double* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= n) return;
    const double* row = W + c * n_pad;
    double s = 0.0;
    for (int64_t r = lane; r < n_pad; r += 32) s = fma(row[r], row[r], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[c] = 1.0 / s;
}

int grad_row_inv_sumsq(const double* W, int64_t n_pad, int64_t n, double* out, cudaStream_t st) {
    row_inv_sumsq_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(W, n_pad, n, out);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int grad_init() {
    if (cudaFuncSetAttribute(grad_tile_kernel<MOGP_KERNEL_SQEXP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(grad_tile_kernel<MOGP_KERNEL_MATERN52, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(grad_tile_kernel<MOGP_KERNEL_SQEXP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM) != cudaSuccess)
        return 1;
    return 0;
}

int grad_max_dims() { return G_MAXD; }
int grad_max_mean() { return G_MAXM_TOTAL; }

int grad_set_identity(double* W, int64_t n_pad, cudaStream_t st) {
    const int64_t total = n_pad * n_pad;
    set_identity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(W, n_pad);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// Wt: (L^-1)^T, row-major n_pad x n_pad; partial: [T(T+1)][d+2] scratch; grad: device [d+2]
int grad_reduce_tiles(const CUtensorMap& tmW128, const CUtensorMap& tmW64, int kernel, const double* XT, int64_t n,
                      int64_t n_pad, int d, const double* alpha, const double* hyper, int fit_nugget, double* partial,
                      double* grad, const double* U, int n_u, int64_t u_stride, cudaStream_t st) {
    if (n_u > G_MAXM_TOTAL) return 1;
    GradParams p{};
    p.U = U; p.n_u = n_u; p.u_stride = u_stride;
    p.XT = XT; p.alpha = alpha; p.hyper = hyper; p.partial = partial; p.n = n; p.n_pad = n_pad; p.d = d; p.kernel = kernel;
    p.row_base = 0;
    const int T = (int)(n_pad / NB);
    const int tiles = T * (T + 1);
    if (kernel == MOGP_KERNEL_SQEXP)
        grad_tile_kernel<MOGP_KERNEL_SQEXP, 0><<<tiles, (G_NCW + 4) * 32, G_SMEM, st>>>(tmW128, tmW64, p);
    else
        grad_tile_kernel<MOGP_KERNEL_MATERN52, 0><<<tiles, (G_NCW + 4) * 32, G_SMEM, st>>>(tmW128, tmW64, p);
    grad_reduce_kernel<<<d + 2, 256, 0, st>>>(partial, tiles, d, alpha, n, hyper, fit_nugget, grad, U, n_u, u_stride);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// C (m_pad x m_pad, lower 128x64 tiles) -= W W^T with W = the rows [row_base, row_base + m_pad) of a K-blocked tensor
// map (box 128 / box 64), contraction over k_len columns: the SYRK of predict(full_cov=True).
int cov_syrk_sub(const CUtensorMap& tmW128, const CUtensorMap& tmW64, int row_base, int64_t m_pad, int64_t k_len, double* C,
                 cudaStream_t st) {
    GradParams p{};
    p.row_base = row_base; p.n_pad = m_pad; p.C = C; p.ldc = m_pad; p.k_chunks = (int)(k_len / KC);
    const int T = (int)(m_pad / NB);
    const int tiles = T * (T + 1);
    grad_tile_kernel<MOGP_KERNEL_SQEXP, 1><<<tiles, (G_NCW + 4) * 32, G_SMEM, st>>>(tmW128, tmW64, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mogp
