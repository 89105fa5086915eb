// Blocked right-looking FP64 Cholesky for libmogp_b200 (replaces cusolverDnDpotrf, reference
// mogp_gpu/src/densegp_gpu.hpp:451-475; semantics of LAPACK dpotrf as used by the CPU reference,
// mogp_emulator/linalg/cholesky.py:225-281).
//
// Storage: row-major lower triangle of an (n_pad x n_pad) matrix, ld = n_pad, n_pad % 128 == 0
// (padding rows/cols carry an identity block).  Per 128-wide block column k:
//   potf2_inv_kernel : warp-cooperative factorisation of the 128x128 diagonal block in shared memory
//                      + its triangular inverse (kept in Dinv for the panel solve, the fit solves and
//                      the predict TRSM) + running log-determinant + LAPACK-style info.
//   tile kernel TRSM : L_ik = A_ik * inv(L_kk)^T  as a DMMA GEMM (64x128 tiles)
//   tile kernel SYRK : A_ij -= L_ik * L_jk^T      as a DMMA GEMM (128x64 tiles, lower tiles only)
// The two GEMM kernels are one template: TMA (cp.async.bulk.tensor.3d) producer warp -> mbarrier
// full/empty ring of K-blocked stages -> 4 consumer warps issuing mma.sync.m16n8k8.f64 (setmaxnreg moves the producer warpgroup's
// registers to them), two CTAs per SM.
#include "common.cuh"
#include "kernels.h"

namespace mogp {

// ------------------------------------------------------------------------------------------
// diagonal block: factor + invert
// ------------------------------------------------------------------------------------------
constexpr int PS = NB + 1;  // shared-memory row stride (doubles): odd => conflict-free column walks
constexpr int SB = 16;      // sub-block width inside the diagonal block

struct Potf2Smem {
    double S[NB * PS];
    double X[NB * (SB + 1)];
    double Lr[SB * (SB + 1)];
    double rd[NB];
    double red[32];
    double ri;
    int fail;
};

__global__ void __launch_bounds__(512, 1)
potf2_inv_kernel(double* __restrict__ A, int64_t ld, int kblk, double* __restrict__ Dinv, int* __restrict__ info,
                 double* __restrict__ logdet) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Potf2Smem& sm = *reinterpret_cast<Potf2Smem*>(smem_raw);
    if (*info != 0) return;
    const int tid = threadIdx.x;
    const int64_t k0 = (int64_t)kblk * NB;
    double* Ablk = A + k0 * ld + k0;

    for (int idx = tid; idx < NB * NB; idx += 512) {
        const int r = idx >> 7, c = idx & 127;
        sm.S[r * PS + c] = (c <= r) ? Ablk[(int64_t)r * ld + c] : 0.0;
    }
    if (tid == 0) sm.fail = 0;
    __syncthreads();

    // ---- factorisation: 8 sub-blocks of 16 columns -------------------------------------------
    for (int s = 0; s < NB / SB; s++) {
        const int j0 = s * SB;
        if (tid < NB) {
            // one thread per row; rows below j0 hold their 16 panel entries in registers
            const int r = tid;
            const bool active = r >= j0;
            double a[SB];
#pragma unroll
            for (int jj = 0; jj < SB; jj++) a[jj] = active ? sm.S[r * PS + j0 + jj] : 0.0;
#pragma unroll
            for (int j = 0; j < SB; j++) {
                if (r == j0 + j) {
                    const double d = a[j];
                    if (!(d > 0.0)) sm.fail = j0 + j + 1;  // also catches NaN (LAPACK: ajj <= 0 or isnan)
                    const double p = sqrt(d);
                    const double ri = 1.0 / p;
                    a[j] = p;
                    sm.ri = ri;
                    sm.rd[r] = ri;
                }
                named_bar_sync(1, NB);
                if (sm.fail) break;
                const double ri = sm.ri;
                if (r > j0 + j) a[j] *= ri;  // LAPACK scales the column by the reciprocal pivot
                if (r >= j0 + j && r < j0 + SB) sm.Lr[(r - j0) * (SB + 1) + j] = a[j];
                named_bar_sync(1, NB);
                if (r > j0 + j) {
#pragma unroll
                    for (int c = j + 1; c < SB; c++) {
                        const double l = sm.Lr[c * (SB + 1) + j];
                        // the entry that becomes a pivot is updated as a - round(l*l) (two roundings, the
                        // dot-then-subtract form of LAPACK's unblocked kernel) so that exactly duplicated
                        // rows fail the un-jittered factorisation the same way the CPU reference does.
                        if (r == j0 + c) a[c] = __dsub_rn(a[c], __dmul_rn(a[j], l));
                        else a[c] = fma(-a[j], l, a[c]);
                    }
                }
            }
            if (active && !sm.fail) {
#pragma unroll
                for (int jj = 0; jj < SB; jj++) sm.S[r * PS + j0 + jj] = (j0 + jj <= r) ? a[jj] : 0.0;
            }
        }
        __syncthreads();
        if (sm.fail) break;
        // trailing update inside the diagonal block: S[r][c] -= sum_kk S[r][j0+kk] * S[c][j0+kk]
        {
            const int r = tid & 127, q = tid >> 7;
            if (r >= j0 + SB) {
                double a[SB];
#pragma unroll
                for (int kk = 0; kk < SB; kk++) a[kk] = sm.S[r * PS + j0 + kk];
                for (int c = j0 + SB + q; c <= r; c += 4) {
                    double dot = 0.0;
#pragma unroll
                    for (int kk = 0; kk < SB; kk++) dot = fma(a[kk], sm.S[c * PS + j0 + kk], dot);
                    sm.S[r * PS + c] -= dot;
                }
            }
        }
        __syncthreads();
    }
    if (sm.fail) {
        if (tid == 0) *info = (int)k0 + sm.fail;
        return;
    }

    // ---- write L_kk back (upper part of the block zeroed) and accumulate log det -------------
    for (int idx = tid; idx < NB * NB; idx += 512) {
        const int r = idx >> 7, c = idx & 127;
        Ablk[(int64_t)r * ld + c] = sm.S[r * PS + c];
    }
    {
        double v = (tid < NB) ? log(sm.S[tid * PS + tid]) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) sm.red[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) *logdet += 2.0 * (sm.red[0] + sm.red[1] + sm.red[2] + sm.red[3]);
    }

    // ---- in-place inverse of the lower-triangular block --------------------------------------
    // (i) the eight 16x16 diagonal sub-blocks: thread (b, c) solves L_bb x = e_c
    {
        double x[SB];
        const int b = tid >> 4, c = tid & 15;
        if (tid < NB) {
            const double* Lb = sm.S + (b * SB) * PS + b * SB;
#pragma unroll
            for (int i = 0; i < SB; i++) {
                double sacc = 0.0;
#pragma unroll
                for (int kk = 0; kk < i; kk++) sacc = fma(Lb[i * PS + kk], x[kk], sacc);
                const double rdi = sm.rd[b * SB + i];
                x[i] = (i == c) ? rdi : ((i > c) ? -sacc * rdi : 0.0);
            }
        }
        __syncthreads();
        if (tid < NB) {
            double* Lb = sm.S + (b * SB) * PS + b * SB;
#pragma unroll
            for (int i = 0; i < SB; i++)
                if (i >= c) Lb[i * PS + c] = x[i];
        }
        __syncthreads();
    }
    // (ii) block columns right to left:  Inv21 = -Inv22 * L21 * Inv11
    for (int J = NB / SB - 2; J >= 0; J--) {
        const int R0 = SB * (J + 1), C0 = SB * J;
        const int r = tid & 127, cq = tid >> 7;
        if (r >= R0) {
            double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
            for (int q = R0; q <= r; q++) {
                const double a = sm.S[r * PS + q];
                const double* bq = sm.S + q * PS + C0 + 4 * cq;
                x0 = fma(a, bq[0], x0);
                x1 = fma(a, bq[1], x1);
                x2 = fma(a, bq[2], x2);
                x3 = fma(a, bq[3], x3);
            }
            double* xr = sm.X + r * (SB + 1) + 4 * cq;
            xr[0] = x0; xr[1] = x1; xr[2] = x2; xr[3] = x3;
        }
        __syncthreads();
        if (r >= R0) {
            const double* xr = sm.X + r * (SB + 1);
#pragma unroll
            for (int cc = 0; cc < 4; cc++) {
                const int c = 4 * cq + cc;
                double y = 0.0;
                for (int kk = c; kk < SB; kk++) y = fma(xr[kk], sm.S[(C0 + kk) * PS + C0 + c], y);
                sm.S[r * PS + C0 + c] = -y;
            }
        }
        __syncthreads();
    }
    double* Dblk = Dinv + k0 * NB;
    for (int idx = tid; idx < NB * NB; idx += 512) {
        const int r = idx >> 7, c = idx & 127;
        Dblk[idx] = (c <= r) ? sm.S[r * PS + c] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------
// DMMA tile kernel (TRSM panel solve and SYRK trailing update)
// ------------------------------------------------------------------------------------------
enum { OP_SYRK = 0, OP_TRSM = 1 };

template <int WGM, int WGN, int NT, int NS>
struct TileCfg {
    static constexpr int BM = 32 * WGM;
    static constexpr int BN = 8 * NT * WGN;
    static constexpr int NCW = WGM * WGN;
    static constexpr int THREADS = (NCW + 4) * 32;  // consumer warpgroup + producer warpgroup (one active lane)
    static constexpr int A_BYTES = BM * KC * 8;
    static constexpr int B_BYTES = BN * KC * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = NS * STAGE_BYTES + 2 * NS * 8 + 128;
};

template <int OP, int WGM, int WGN, int NT, int NS>
__global__ void __launch_bounds__(TileCfg<WGM, WGN, NT, NS>::THREADS, 2)
chol_tile_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 double* __restrict__ A, int64_t ld, int row_base, int kblk, const int* __restrict__ info) {
    // A: slab base; row_base: first slab row of this output's matrix (also its first Dinv slab row)
    using Cfg = TileCfg<WGM, WGN, NT, NS>;
    extern __shared__ __align__(128) unsigned char tile_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tile_smem_raw) + 127) & ~uintptr_t(127));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + NS * Cfg::STAGE_BYTES);
    uint64_t* empty = full + NS;

    if (*info != 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int arow, brow, a_kout0, b_kout0;
    if (OP == OP_SYRK) {
        const int id = blockIdx.x;
        int I = (int)((sqrtf(4.0f * (float)id + 1.0f) - 1.0f) * 0.5f);
        while ((I + 1) * (I + 2) <= id) I++;
        while (I * (I + 1) > id) I--;
        const int J2 = id - I * (I + 1);
        arow = row_base + (kblk + 1 + I) * NB;
        brow = row_base + (kblk + 1) * NB + J2 * Cfg::BN;
        a_kout0 = b_kout0 = kblk * (NB / 8);
    } else {
        arow = row_base + (kblk + 1) * NB + blockIdx.x * Cfg::BM;
        brow = row_base + kblk * NB;  // rows of the Dinv slab
        a_kout0 = kblk * (NB / 8);
        b_kout0 = 0;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::NCW);
        }
        fence_mbar_init();
    }
    __syncthreads();

    constexpr int NCHUNK = NB / KC;
    if (warp >= Cfg::NCW) {
        // ---- TMA producer warpgroup (hands its registers to the consumers) ----
        reg_dealloc<24>();
        if (warp == Cfg::NCW && lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmB);
            PipeState<NS> ps;
            for (int c = 0; c < NCHUNK; c++) {
                mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(&full[ps.stage], Cfg::STAGE_BYTES);
                tma_load_3d(st, &tmA, 0, arow, a_kout0 + c * (KC / 8), &full[ps.stage]);
                tma_load_3d(st + Cfg::A_BYTES, &tmB, 0, brow, b_kout0 + c * (KC / 8), &full[ps.stage]);
                ps.advance();
            }
        }
        return;
    }

    // ---- DMMA consumers ----
    reg_alloc<232>();
    const int wm = warp / WGN, wn = warp % WGN;
    const int g = lane >> 2, t = lane & 3;
    double acc[2][NT][4];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.0;

    PipeState<NS> ps;
    for (int c = 0; c < NCHUNK; c++) {
        mbar_wait(&full[ps.stage], ps.phase);
        const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
        const double* Bs = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES + Cfg::A_BYTES);
        mma_stage<2, NT, KC>(acc, As, Cfg::BM, wm * 32, Bs, Cfg::BN, wn * 8 * NT, g, t);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[ps.stage]);
        ps.advance();
    }

    // ---- epilogue ----
    const int64_t crow0 = arow + wm * 32 + g;
    const int64_t ccol0 = (OP == OP_SYRK ? brow - row_base : kblk * NB) + wn * 8 * NT + 2 * t;
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            double* rowp = A + (crow0 + mt * 16 + h * 8) * ld + ccol0;
            if (OP == OP_SYRK) {
                double2 v[NT];
#pragma unroll
                for (int nt = 0; nt < NT; nt++) v[nt] = *reinterpret_cast<const double2*>(rowp + nt * 8);
#pragma unroll
                for (int nt = 0; nt < NT; nt++) {
                    v[nt].x -= acc[mt][nt][2 * h];
                    v[nt].y -= acc[mt][nt][2 * h + 1];
                    *reinterpret_cast<double2*>(rowp + nt * 8) = v[nt];
                }
            } else {
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    *reinterpret_cast<double2*>(rowp + nt * 8) = make_double2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
            }
        }
    }
}

using SyrkCfg = TileCfg<4, 1, 8, 4>;  // 128 x 64 tiles
using TrsmCfg = TileCfg<2, 2, 8, 4>;  // 64 x 128 tiles

int chol_init() {
    static bool done = false;
    if (done) return 0;
    cudaError_t e;
    e = cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Potf2Smem));
    if (e != cudaSuccess) return 1;
    e = cudaFuncSetAttribute(chol_tile_kernel<OP_SYRK, 4, 1, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SyrkCfg::SMEM_BYTES);
    if (e != cudaSuccess) return 1;
    e = cudaFuncSetAttribute(chol_tile_kernel<OP_TRSM, 2, 2, 8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             TrsmCfg::SMEM_BYTES);
    if (e != cudaSuccess) return 1;
    done = true;
    return 0;
}

int chol_make_maps(CholMaps* maps, double* A_slab, double* Dinv_slab, int64_t total_rows, int64_t n_pad) {
    if (make_kblocked_tmap(&maps->a128, A_slab, total_rows, n_pad, 128)) return 1;
    if (make_kblocked_tmap(&maps->a64, A_slab, total_rows, n_pad, 64)) return 1;
    if (make_kblocked_tmap(&maps->d128, Dinv_slab, total_rows, NB, 128)) return 1;
    return 0;
}

// Enqueue the whole factorisation of output `o` of the slab on `st`.  info/logdet must have been zeroed.
// Returns the number of kernels launched (negative on launch error).
int chol_factor(const CholMaps& maps, double* A_slab, double* Dinv_slab, int o, int64_t n_pad, int* info,
                double* logdet, cudaStream_t st) {
    const int T = (int)(n_pad / NB);
    const int row_base = (int)(o * n_pad);
    double* A = A_slab + (int64_t)row_base * n_pad;
    double* Dinv = Dinv_slab + (int64_t)row_base * NB;
    int launches = 0;
    for (int k = 0; k < T; k++) {
        potf2_inv_kernel<<<1, 512, sizeof(Potf2Smem), st>>>(A, n_pad, k, Dinv, info, logdet);
        launches++;
        const int Tt = T - k - 1;
        if (Tt > 0) {
            chol_tile_kernel<OP_TRSM, 2, 2, 8, 4>
                <<<Tt * (NB / TrsmCfg::BM), TrsmCfg::THREADS, TrsmCfg::SMEM_BYTES, st>>>(maps.a64, maps.d128, A_slab, n_pad, row_base, k, info);
            chol_tile_kernel<OP_SYRK, 4, 1, 8, 4>
                <<<Tt * (Tt + 1), SyrkCfg::THREADS, SyrkCfg::SMEM_BYTES, st>>>(maps.a128, maps.a64, A_slab, n_pad, row_base, k, info);
            launches += 2;
        }
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace mogp
