// Blocked FP64 Cholesky for libmogp_b200 (replaces cusolverDnDpotrf, reference
// mogp_gpu/src/densegp_gpu.hpp:451-475; semantics of LAPACK dpotrf as used by the CPU reference,
// mogp_emulator/linalg/cholesky.py:225-281).
//
// Storage: row-major lower triangle of an (n_pad x n_pad) matrix per output, ld = n_pad, n_pad % 128 == 0
// (padding rows/cols carry an identity block).  The factorisation of ALL outputs of a call is ONE persistent
// dataflow kernel (one CTA per SM) that draws tiles from a global ticket counter in block-column-major order:
//
//   ROW  tile (i, p, j), i > j : the 64 rows p of block row i, block column j (left-looking)
//            L_ij = (A_ij - sum_{k<j} L_ik L_jk^T) inv(L_jj)^T
//   DIAG tile (j, p, j)        : R_jj = A_jj - sum_{k<j} L_jk L_jk^T   (64 rows p of the diagonal block)
//   D    tile (j)              : L_jj = chol(R_jj), its triangular inverse (Dinv, used by every later solve),
//                                log-determinant, LAPACK-style info
//
// Both products of a ROW tile are FP64 tensor-pipe GEMMs (mma.sync.m16n8k8.f64) in TN form, computed transposed
// (block-column rows x panel rows) so the 64-row panel is the N side: the L_jk / inv(L_jj) tiles (A operand) and
// the panel's own solved tiles L_(i,p),k (B operand) stream through a TMA -> mbarrier ring, A_ij lands in a
// resident staging buffer that becomes the B operand of the diagonal product -- the same tile body as the
// predict TRSM (predict.cu): a Cholesky panel row IS a forward substitution against the rows above it.
// Every dependency of a tile has a smaller ticket, hence is held by a running CTA: progress counters in global
// memory (red.release / ld.acquire + proxy fences for the TMA readers) order the tiles, nothing deadlocks, and
// there is no per-step launch, tail or wave quantisation; outputs interleave in the ticket order, so many small
// factorisations fill the machine as well as one large one.
//
// Two kernels share the tiles, the ticket order, the progress counters and the D tile (chol_d_tile):
//   chol_dataflow_kernel   the products above on the FP64 tensor pipe -- launches of a few small matrices, which are bound by
//                          the chain D(j) -> ROW(j+1, ., j) -> DIAG(j+1, .) -> D(j+1);
//   chol_i8_kernel         the history products sum_k L_ik L_jk^T as exact integer GEMM on 8 signed 7-bit planes per operand
//                          (tcgen05.mma kind::i8, s32 accumulators in TMEM), FP64 epilogue -- launches bound by their O(n^3)
//                          arithmetic (see the comment in front of it); it also leaves the planes of L the predict TRSM reads.
#include <type_traits>

#include "i8_common.cuh"
#include "kernels.h"

namespace mogp {

// -DCHOL_TRACE: globaltimer stamps of the diagonal (D) tiles for tools/chol_dtile_timeline.py; compiled out of the product
#ifdef CHOL_TRACE
constexpr int CH_TRACE_EV = 8, CH_TRACE_N = 4096;
__device__ unsigned long long chol_trace_buf[CH_TRACE_N * CH_TRACE_EV];
__device__ int chol_trace_count;
#define CH_STAMP(slot, ev) do { if ((slot) >= 0 && (slot) < CH_TRACE_N) chol_trace_buf[(slot) * CH_TRACE_EV + (ev)] = globaltimer_ns(); } while (0)
#else
#define CH_STAMP(slot, ev) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// diagonal block: factor + invert, cooperatively by NTHR threads of one CTA (device function)
// ------------------------------------------------------------------------------------------
constexpr int PS = NB + 1;  // shared-memory row stride (doubles): odd => conflict-free column walks
constexpr int SB = 16;      // sub-block width inside the diagonal block

constexpr int NSB = NB / SB;                  // 8 sub-blocks per side

struct Potf2Smem {
    double S[NB * PS];                         // lower: L; upper blocks (J, b): the off-diagonal blocks X_bJ of inv(L)
    double InvD[NSB * SB * (SB + 1)];          // inverses of the eight 16x16 diagonal sub-blocks of L (zero above the diagonal)
    double Tmp[(NSB - 1) * SB * (SB + 1)];     // per (b, J) pair: sum_b' L_bb' X_b'J
    double Lr[2][SB * (SB + 1)];               // L11 of the current panel (parity of the panel: the chain warp is one panel ahead)
    double rd[NB];
    double red[32];
    double ri;
    int fail;
#ifdef CHOL_TRACE
    unsigned long long stamps[32];
#endif
};

// pv = sqrt(d) and ri = 1 / sqrt(d) for a pivot d > 0.  A dependent FP64 instruction costs ~45-55 cycles on this chip and the
// pivots of a factorisation are one dependency chain (128 per diagonal tile, on the critical path of every block column;
// sqrt() followed by 1.0 / pv is ~17 dependent operations).  Only the reciprocal is on that chain (the next pivot needs l = a * ri), so it is taken straight from the refined reciprocal square
// root -- hardware seed (MUFU.RSQ64H, what sqrt() starts from too) + one third-order step: 4 dependent operations, error
// below an ulp -- and the square root itself (FMA-corrected d * y, correctly rounded) is computed off the chain.  LAPACK
// scales by fl(1 / fl(sqrt(d))); this reciprocal differs from that by at most an ulp or two.
__device__ __forceinline__ void sqrt_and_reciprocal(double d, double& pv, double& ri) {
    double y0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double t = y0 * y0;
    const double e = fma(-d, t, 1.0);                 // 1 - d y0^2
    const double y1 = fma(fma(e, 0.375, 0.5), y0 * e, y0);      // y0 (1 + e/2 + 3 e^2/8)
    ri = y1;
    const double s0 = d * y1;
    pv = fma(fma(-s0, s0, d), 0.5 * y1, s0);
}

// (a1) of potf2_inv_block: Cholesky of the 16x16 diagonal sub-block at (j0, j0) by ONE warp.  Lane r (mod 16) owns
// row j0+r in registers; pivots and multipliers travel by shuffle, so the 16-step dependency chain has no barrier.
// Writes L11 back to sm.S and sm.Lr, the reciprocal pivots to sm.rd, or sets sm.fail (1-based failing column).
//
// This is ONE warp's instruction stream on the critical path of the whole factorisation: what counts is its length.
//   * Every lane updates every column c > j with one FMA; entries above the diagonal (c > r) turn into garbage that nothing
//     reads (shuffles read a[j] of the lanes c > j only, lanes 16..31 hold zeros).  Selects that kept them clean, and the
//     two-rounding form computed beside the FMA for every entry, were 2/3 of ~230 instructions per pivot.
//   * The entry that becomes the lane's own pivot lives in a separate register and is updated as d - round(l*l) (two
//     roundings, the dot-then-subtract form of LAPACK's unblocked kernel) so that exactly duplicated rows fail the un-jittered
//     factorisation the same way the CPU reference does.  Its multiplier is the lane's own a[j]: no shuffle on that chain.
//   * The warp must arrive converged: the shuffles take a software path (BRA.DIV) otherwise, ~10 x slower -- a lane-0 trace
//     stamp in front of the call did exactly that to the timeline tool (a "34 us first pivot block" that the product never had).
//   (Rolling the pivot loop around a shifting register window -- 90 instructions instead of 16 x 85 -- was measured: 5.5 us per
//   block against 3.5 us, every pivot then updates all 15 columns; instruction fetch is not what bounds this function.)
__device__ __noinline__ void potf2_diag16(Potf2Smem& sm, int j0, int lane) {
    __syncwarp();
    const int r = lane & 15;
    const bool act = lane < SB;                        // lanes 16..31 idle along
    double a[SB];
#pragma unroll
    for (int jj = 0; jj < SB; jj++) a[jj] = (jj <= r && act) ? sm.S[(j0 + r) * PS + j0 + jj] : 0.0;
    double diag = act ? sm.S[(j0 + r) * PS + j0 + r] : 1.0;
    volatile double* Lr = sm.Lr[(j0 / SB) & 1];        // column j of L11 is exchanged through sm.Lr (where it has to end up anyway)
    volatile double* dd = &sm.ri;                      // the next pivot
    if (lane == 0) *dd = diag;
    __syncwarp();
    int fail = 0;
#pragma unroll
    for (int j = 0; j < SB; j++) {
        const double d = *dd;
        // also catches NaN (LAPACK: ajj <= 0 or isnan); warp-uniform.  No early exit: the loop stays fully
        // unrolled (register-resident a[]); what follows a failed pivot is never used.
        if (fail == 0 && !(d > 0.0)) fail = j0 + j + 1;
        double pv, ri;
        sqrt_and_reciprocal(d, pv, ri);
        a[j] = (r == j) ? pv : a[j] * ri;   // LAPACK scales the column by the reciprocal pivot
        if (lane == 0) sm.rd[j0 + j] = ri;
        diag = __dsub_rn(diag, __dmul_rn(a[j], a[j]));
        __syncwarp();                        // every lane has read the pivot
        if (act) Lr[r * (SB + 1) + j] = (r >= j) ? a[j] : 0.0;
        if (lane == j + 1) *dd = diag;
        __syncwarp();
#pragma unroll
        for (int c = j + 1; c < SB; c++) a[c] = fma(-a[j], Lr[c * (SB + 1) + j], a[c]);
    }
    if (fail) {
        if (lane == 0) sm.fail = fail;
    } else if (act) {
#pragma unroll
        for (int jj = 0; jj < SB; jj++)
            if (jj <= r) sm.S[(j0 + r) * PS + j0 + jj] = a[jj];
    }
}

// On entry sm.S holds the lower triangle of the block (upper part zero).  On success (returns 0) the diagonal 16x16
// blocks of sm.InvD hold inv(L)'s diagonal blocks and the upper blocks (J, b) of sm.S its off-diagonal blocks X_bJ, the lower
// triangle of sm.S holds L (the caller writes both out), and *logdet_add is 2*sum(log L_ii) on thread 0.  On failure returns the 1-based index of the failing pivot (LAPACK info).
// BAR_ALL: named barrier id for the NTHR participating threads.
#ifdef CHOL_TRACE
__device__ int chol_trace_cur[256];     // per SM: slot of the D tile in flight (trace build only)
#define CH_STAMP_IN(ev) do { if (tid == 0) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); CH_STAMP(chol_trace_cur[sm_ & 255], ev); } } while (0)
__device__ unsigned long long chol_trace_buf2[CH_TRACE_N * 32];
// (stamps go to shared memory and are flushed after the tile: a global store per stamp perturbed what it measured)
#define CH_STAMP2(cond, ev) do { if (cond) sm.stamps[ev] = globaltimer_ns(); } while (0)
#define CH_STAMP2_FLUSH() do { if (tid < 32) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); const int sl_ = chol_trace_cur[sm_ & 255]; if (sl_ >= 0 && sl_ < CH_TRACE_N) chol_trace_buf2[sl_ * 32 + tid] = sm.stamps[tid]; } } while (0)
#else
#define CH_STAMP_IN(ev) do { } while (0)
#define CH_STAMP2(cond, ev) do { } while (0)
#define CH_STAMP2_FLUSH() do { } while (0)
#endif

template <int NTHR, int BAR_ALL>
__device__ __forceinline__ int potf2_inv_block(Potf2Smem& sm, int tid, double* logdet_add) {
    constexpr int NG = NTHR / NB;   // thread groups of 128
    static_assert(NTHR % NB == 0 && NG >= 1 && SB % NG == 0, "thread count");
    if (tid == 0) sm.fail = 0;
    named_bar_sync(BAR_ALL, NTHR);

    // ---- factorisation: 8 panels of 16 columns ------------------------------------------------
    // Per panel s (columns j0 = 16 s ..):  (a1) Cholesky of the 16x16 diagonal sub-block by warp 0 (potf2_diag16),
    // (a2) the rows below it, one thread per row, (b) the trailing update of the block.  The 16-step dependency
    // chain of (a1) is the long pole, so (b) is split: (b1) first the 16 columns the NEXT panel needs (all threads),
    // then warp 0 runs (a1) of panel s+1 while warps 1.. run (b2), the rest of the trailing update.
    auto rows_below = [&](int j0, int r) {   // (a2) of row r: needs only L11 (sm.Lr) and the reciprocal pivots (sm.rd) of the panel
        const double* Lr = sm.Lr[(j0 / SB) & 1];
        double a[SB];
#pragma unroll
        for (int jj = 0; jj < SB; jj++) a[jj] = sm.S[r * PS + j0 + jj];
#pragma unroll
        for (int j = 0; j < SB; j++) {
            a[j] *= sm.rd[j0 + j];
#pragma unroll
            for (int c = j + 1; c < SB; c++) a[c] = fma(-a[j], Lr[c * (SB + 1) + j], a[c]);
        }
#pragma unroll
        for (int jj = 0; jj < SB; jj++) sm.S[r * PS + j0 + jj] = a[jj];
    };
    // ---- inverse of the block, one 16-row block row at a time -------------------------------------------------------
    // inv_diag16(b): Inv_bb by one warp, lane c (< 16) solves L_bb x = e_c with the stored reciprocal pivots.
    // inv_offdiag(b, J): X_bJ = -Inv_bb * sum_{b'=J}^{b-1} L_bb' X_b'J (X_JJ = Inv_JJ) by one warp, every lane a 2x4 micro-tile
    // (rows ti, ti+8; columns 4tc..4tc+3) with four independent accumulator sets over k % 4 (a dependent DFMA costs ~50
    // cycles); X_bJ is parked in the unused UPPER block (J, b) of sm.S (element [i][c] at S[J*16+i][b*16+c]).
    // Block row b needs only the finished panels <= b of L and the block rows < b of the inverse, so it is computed by the
    // warps 1..7 in the shadow of warp 0's pivot block b+1 (the long pole of the factorisation) instead of after it.
    auto inv_diag16 = [&](int b, int lane) {
        if (lane < SB) {
            const int c = lane;
            const double* Lb = sm.S + (b * SB) * PS + b * SB;
            double* Ib = sm.InvD + b * SB * (SB + 1);
            double x[SB];
#pragma unroll
            for (int i = 0; i < SB; i++) {
                double sacc = 0.0;
#pragma unroll
                for (int kk = 0; kk < i; kk++) sacc = fma(Lb[i * PS + kk], x[kk], sacc);
                const double rdi = sm.rd[b * SB + i];
                x[i] = (i == c) ? rdi : ((i > c) ? -sacc * rdi : 0.0);
            }
#pragma unroll
            for (int i = 0; i < SB; i++) Ib[i * (SB + 1) + c] = x[i];
        }
    };
    // FP64 tensor pipe: both products are 16 x 16 x 16 (two column tiles of 8, two K steps each); T passes through sm.Tmp to turn
    // from accumulator layout into a B operand.  (Scalar 2 x 4 micro-tiles took 1.1 .. 5.3 us per block row -- longer than the
    // chain warp's panel -- and 4.8 us for the last one, which is not in anybody's shadow.)
    auto inv_offdiag = [&](int b, int J, int lane) {
        if (J >= b) return;
        const int g = lane >> 2, t4 = lane & 3;
        double acc[2][4];
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[n][e] = 0.0;
        for (int bp = J; bp < b; bp++) {
            // X_b'J[k][c]: the inverted diagonal block (b' == J) or the parked block (J, b')
            const double* xs = (bp == J) ? (sm.InvD + J * SB * (SB + 1)) : (sm.S + (J * SB) * PS + bp * SB);
            const int xstride = (bp == J) ? (SB + 1) : PS;
            const double* la = sm.S + (b * SB + g) * PS + bp * SB + t4;
#pragma unroll
            for (int kk = 0; kk < SB; kk += 8) {
                const double a0 = la[kk], a1 = la[8 * PS + kk], a2 = la[kk + 4], a3 = la[8 * PS + kk + 4];
#pragma unroll
                for (int n = 0; n < 2; n++)
                    dmma_16x8x8(acc[n], a0, a1, a2, a3, xs[(kk + t4) * xstride + 8 * n + g], xs[(kk + t4 + 4) * xstride + 8 * n + g]);
            }
        }
        double* tp = sm.Tmp + J * SB * (SB + 1);
#pragma unroll
        for (int n = 0; n < 2; n++) {
            tp[g * (SB + 1) + 8 * n + 2 * t4] = acc[n][0];
            tp[g * (SB + 1) + 8 * n + 2 * t4 + 1] = acc[n][1];
            tp[(g + 8) * (SB + 1) + 8 * n + 2 * t4] = acc[n][2];
            tp[(g + 8) * (SB + 1) + 8 * n + 2 * t4 + 1] = acc[n][3];
        }
        __syncwarp();
        const double* ia = sm.InvD + b * SB * (SB + 1) + g * (SB + 1) + t4;      // Inv_bb (zero above the diagonal)
        double y[2][4];
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
            for (int e = 0; e < 4; e++) y[n][e] = 0.0;
#pragma unroll
        for (int kk = 0; kk < SB; kk += 8) {
            const double a0 = ia[kk], a1 = ia[8 * (SB + 1) + kk], a2 = ia[kk + 4], a3 = ia[8 * (SB + 1) + kk + 4];
#pragma unroll
            for (int n = 0; n < 2; n++)
                dmma_16x8x8(y[n], a0, a1, a2, a3, tp[(kk + t4) * (SB + 1) + 8 * n + g], tp[(kk + t4 + 4) * (SB + 1) + 8 * n + g]);
        }
        double* xo = sm.S + (J * SB + g) * PS + b * SB + 2 * t4;
#pragma unroll
        for (int n = 0; n < 2; n++) {
            xo[8 * n] = -y[n][0];
            xo[8 * n + 1] = -y[n][1];
            xo[8 * PS + 8 * n] = -y[n][2];
            xo[8 * PS + 8 * n + 1] = -y[n][3];
        }
    };
    // C[R0 + i][C0 + c] -= sum_k P[R0 + i][k] P[C0 + c][k] for a 16 x (8 NT) tile, P = S[:, j0 .. j0+15]: DMMA into a zero
    // accumulator, then ONE subtraction per entry (the dot-then-subtract form of the scalar update_row)
    auto update_tile = [&](int j0, int R0, int C0, int lane, auto nt_tag) {
        constexpr int NT = decltype(nt_tag)::value;
        const int g = lane >> 2, t4 = lane & 3;
        const double* pa = sm.S + (R0 + g) * PS + j0 + t4;
        double acc[NT][4];
#pragma unroll
        for (int n = 0; n < NT; n++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[n][e] = 0.0;
#pragma unroll
        for (int kk = 0; kk < SB; kk += 8) {
            const double a0 = pa[kk], a1 = pa[8 * PS + kk], a2 = pa[kk + 4], a3 = pa[8 * PS + kk + 4];
#pragma unroll
            for (int n = 0; n < NT; n++) {
                const double* pb = sm.S + (C0 + 8 * n + g) * PS + j0 + t4 + kk;
                dmma_16x8x8(acc[n], a0, a1, a2, a3, pb[0], pb[4]);
            }
        }
        double* pc = sm.S + (R0 + g) * PS + C0 + 2 * t4;
#pragma unroll
        for (int n = 0; n < NT; n++) {
            pc[8 * n] -= acc[n][0];
            pc[8 * n + 1] -= acc[n][1];
            pc[8 * PS + 8 * n] -= acc[n][2];
            pc[8 * PS + 8 * n + 1] -= acc[n][3];
        }
    };

    // The chain of the tile is a1(0) -> [a2 of the next diagonal block's 16 rows -> their 16x16 update -> a1(s+1)] x 7, and ONE warp
    // walks it without waiting for anybody: warp 0 solves the 16 rows below its pivot block itself (a2n), applies them to the next
    // diagonal block (b1n, 4 DMMAs) and goes straight on to that block's pivots.  The other seven warps follow one panel behind: the
    // rows further down (a2r), block row s of the inverse, then -- once warp 0's 16 rows are in shared memory (named barrier
    // BAR_ALL + 2: warp 0 only arrives) -- the trailing update of everything below the next diagonal block (b1r + b2).  One barrier
    // of all warps per panel.  (Before: a1, a2 and b1 each ended in a barrier of all eight warps, 5 us of every 8.4 us panel.)
    // Measured per panel (tools/chol_dtile_timeline.py): chain warp 4.4 - 5.5 us, the others 3.3 - 4.4 us.
    CH_STAMP2(tid == 0, 0);
    if (tid < 32) potf2_diag16(sm, 0, tid);
    CH_STAMP2(tid == 0, 1);
    named_bar_sync(BAR_ALL, NTHR);
    if (sm.fail) return sm.fail;
    for (int s = 0; s + 1 < NB / SB; s++) {
        const int j0 = s * SB;
        const int nrow2 = NB - j0 - 2 * SB;   // rows below the next diagonal block
        if (tid < 32) {
            if (tid < SB) rows_below(j0, j0 + SB + tid);                                            // a2n
            __syncwarp();
            update_tile(j0, j0 + SB, j0 + SB, tid, std::integral_constant<int, 2>());                // b1n (entries above the diagonal: unused)
            __threadfence_block();
            asm volatile("bar.arrive %0, %1;" ::"r"(BAR_ALL + 2), "r"(NTHR) : "memory");
            __syncwarp();
            potf2_diag16(sm, j0 + SB, tid);                                                         // a1 of panel s+1
            CH_STAMP2(tid == 0, 2 + 3 * s);       // chain warp done with panel s+1's pivots
        } else {
            static_assert(NTHR - 32 >= NB - 2 * SB, "one thread per row below the next diagonal block");
            const int tt = tid - 32;
            if (tt < nrow2) rows_below(j0, j0 + 2 * SB + tt);                                        // a2r (warps 1..3 at most)
            // block row s of the inverse (panel s of L is final): Inv_ss by warp 7 (it has no row of a2r), then X_sJ for J < s by the
            // warps 1..7 -- before the trailing update, which has to wait for the chain warp's 16 rows anyway
            if (tid >= NTHR - 32) inv_diag16(s, tid & 31);
            named_bar_sync(BAR_ALL + 1, NTHR - 32);
            inv_offdiag(s, (tid >> 5) - 1, tid & 31);
            CH_STAMP2(tid == 32, 3 + 3 * s);      // block row s of the inverse done (warp 1)
            named_bar_sync(BAR_ALL + 2, NTHR);
            // (b1r + b2) the trailing update below the next diagonal block on the FP64 tensor pipe:
            //     S[r][c] -= sum_k P[r][k] P[c][k],   P = S[:, j0 .. j0+15],   rows r >= j0+32, columns j0+16 <= c <= r
            // in tiles of 16 rows x 16 columns (K = 16: DMMAs into zero accumulators, then ONE subtraction per entry -- the
            // dot-then-subtract form of the scalar update it replaces), dealt round-robin to the seven warps.  Tiles that cross the
            // diagonal also write entries above it: nothing reads those (potf2_diag16, rows_below and this update read the lower
            // triangle, the parked inverse blocks of these rows are written later, the write-back masks).  Scalar FMAs from shared
            // memory (one 8-byte load per FMA) took 8.3 us for panel 0 -- twice the chain warp's panel; this runs at the SM's FP64
            // rate (216 DMMAs for panel 0: 1.8 us).
            {
                const int wid = (tid >> 5) - 1, lane = tid & 31;
                const int mt = nrow2 / SB;                       // 16-row tiles; tile row mi has 2 + mi PAIRS of column tiles
                const int total = mt * (mt + 3) / 2;
                for (int tile = wid; tile < total; tile += NTHR / 32 - 1) {
                    int mi = 0;
                    while ((mi + 1) * (mi + 4) / 2 <= tile) mi++;
                    const int ni = tile - mi * (mi + 3) / 2;
                    update_tile(j0, j0 + 2 * SB + SB * mi, j0 + SB + 16 * ni, lane, std::integral_constant<int, 2>());
                }
            }
            CH_STAMP2(tid == 32, 4 + 3 * s);      // trailing update done (warp 1)
        }
        named_bar_sync(BAR_ALL, NTHR);
        if (sm.fail) return sm.fail;
    }
    if (sm.fail) return sm.fail;
    CH_STAMP_IN(2);
    CH_STAMP2_FLUSH();

    // ---- log det (L_kk itself is written back by the caller, after inv(L_kk) has been published) -------------
    {
        double v = (tid < NB) ? log(sm.S[tid * PS + tid]) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (tid < NB && (tid & 31) == 0) sm.red[tid >> 5] = v;
        named_bar_sync(BAR_ALL, NTHR);
        if (tid == 0) *logdet_add = 2.0 * ((sm.red[0] + sm.red[1]) + (sm.red[2] + sm.red[3]));
    }

    CH_STAMP_IN(3);
    // ---- last block row of the inverse (the others were computed in the shadow of the pivot blocks, see the loop above) ----
    if (tid >= 32 && tid < 64) inv_diag16(NSB - 1, tid & 31);
    named_bar_sync(BAR_ALL, NTHR);
    CH_STAMP_IN(4);
    if (tid >= 32) inv_offdiag(NSB - 1, (tid >> 5) - 1, tid & 31);
    named_bar_sync(BAR_ALL, NTHR);
    return 0;
}

// The D tile: R_jj (written by the two DIAG tiles, possibly on other SMs) -> shared memory, factor + invert (potf2_inv_block),
// inv(L_jj) to the Dinv slab, log-determinant, LAPACK info, PUBLISH (dprog: the ROW tiles of column j wait for inv(L_jj) only),
// then L_jj back to the matrix (nothing inside the factorisation reads the diagonal block of L again).  NTHR consumer threads
// (ctid), `base`: the borrowed shared memory (>= sizeof(Potf2Smem)); named barriers 1 (all NTHR threads), 2, 3 and 4.  Ends with
// barrier 1: every thread is done with the shared memory.
template <int NTHR>
__device__ __forceinline__ void chol_d_tile(unsigned char* base, int ctid, double* A, double* Dinv, int64_t n_pad, int64_t rb,
                                            int j, int o, int* info, double* scal, int* dprog, bool skip, int trs) {
    (void)trs;
    Potf2Smem& sm = *reinterpret_cast<Potf2Smem*>(base);
    double* Ablk = A + (rb + (int64_t)j * NB) * n_pad + (int64_t)j * NB;
    int fail = 1;
    if (!skip) {
        // 16-byte loads, eight in flight per thread (a single 8-byte load per iteration left the 128 KB block waiting on 64
        // sequential L2 round trips: 16 us)
        {
            static_assert(NTHR == 256, "load mapping");
            const int cp = ctid & 63, r0 = ctid >> 6;          // column pair, first row; rows r0, r0 + 4, ...
#pragma unroll
            for (int it = 0; it < 32; it += 8) {
                double2 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int r = r0 + 4 * (it + u);
                    v[u] = (2 * cp <= r) ? __ldcg(reinterpret_cast<const double2*>(Ablk + (int64_t)r * n_pad) + cp)
                                         : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int r = r0 + 4 * (it + u);
                    sm.S[r * PS + 2 * cp] = (2 * cp <= r) ? v[u].x : 0.0;
                    sm.S[r * PS + 2 * cp + 1] = (2 * cp + 1 <= r) ? v[u].y : 0.0;
                }
            }
        }
        double ld_add = 0.0;
#ifdef CHOL_TRACE
        named_bar_sync(1, NTHR);
        if (ctid == 0) CH_STAMP(trs, 1);
#endif
        fail = potf2_inv_block<NTHR, 2>(sm, ctid, &ld_add);
#ifdef CHOL_TRACE
        if (ctid == 0) CH_STAMP(trs, 5);
#endif
        if (fail) {
            if (ctid == 0) info[o] = j * NB + fail;
        } else {
            // inv(L_jj): two columns per thread and store (a pair never straddles a 16-column block)
            double* Dblk = Dinv + (rb + (int64_t)j * NB) * NB;
            for (int idx = ctid; idx < NB * NB / 2; idx += NTHR) {
                const int r = idx >> 6, c = (idx & 63) * 2;
                const int b = r >> 4, J = c >> 4;
                double2 v = make_double2(0.0, 0.0);
                if (b == J) {                                                            // (zero above the diagonal)
                    const double* q = sm.InvD + (b * SB + (r & 15)) * (SB + 1) + (c & 15);
                    v = make_double2(q[0], q[1]);
                } else if (b > J) {                                                      // parked in the upper block (J, b)
                    const double* q = sm.S + (J * SB + (r & 15)) * PS + b * SB + (c & 15);
                    v = make_double2(q[0], q[1]);
                }
                *reinterpret_cast<double2*>(Dblk + 2 * idx) = v;
            }
            // the D tiles of one output run strictly in order: plain read-modify-write is race-free (and it must precede the
            // publication: the next D tile of this output can start as soon as the tiles that wait for this one have run)
            if (ctid == 0) scal[2 * o] = (j > 0 ? __ldcg(scal + 2 * o) : 0.0) + ld_add;
        }
    }
    __threadfence();
    fence_proxy_async();
    named_bar_sync(1, NTHR);
#ifdef CHOL_TRACE
    if (ctid == 0) CH_STAMP(trs, 6);
#endif
    if (ctid == 0) red_release_gpu_add(dprog, 1);
    if (!fail) {
        for (int idx = ctid; idx < NB * NB / 2; idx += NTHR) {
            const int r = idx >> 6, c = (idx & 63) * 2;
            const double* q = sm.S + r * PS + c;
            // (above the diagonal: parked inverse blocks and spill of the trailing updates)
            *reinterpret_cast<double2*>(Ablk + (int64_t)r * n_pad + c) = make_double2(c <= r ? q[0] : 0.0, c + 1 <= r ? q[1] : 0.0);
        }
        __threadfence();
    }
    named_bar_sync(1, NTHR);   // everyone is done with the borrowed shared memory
}

// ------------------------------------------------------------------------------------------
// the dataflow kernel
// ------------------------------------------------------------------------------------------
struct CholCfg {
    static constexpr int BN = 64;                  // panel rows per tile (the N side of the transposed products)
    static constexpr int NT = BN / 16;
    static constexpr int NCW = 8;                  // DMMA consumer warps
    static constexpr int THREADS = (NCW + 4) * 32; // + producer warpgroup (one active lane)
    static constexpr int NS = 4;
    static constexpr int A_BYTES = NB * KC * 8;
    static constexpr int B_BYTES = BN * KC * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int VS_BYTES = NB * BN * 8;
    static constexpr int BAR_BYTES = (2 * NS + 2 + 4 + 1) * 8 + 16;
    // the diagonal-block factorisation borrows the ring + staging buffer (and some more behind them)
    static constexpr int TILE_BYTES = NS * STAGE_BYTES + VS_BYTES;
    static constexpr int REGION_BYTES = ((int)sizeof(Potf2Smem) > TILE_BYTES ? ((int)sizeof(Potf2Smem) + 127) / 128 * 128 : TILE_BYTES);
    static constexpr int SMEM_BYTES = REGION_BYTES + BAR_BYTES + 128;
};
static_assert(CholCfg::SMEM_BYTES <= 232448, "shared memory per CTA");

constexpr int CH_HDR = 32;   // ints in front of the per-output progress blocks (word 0 = ticket counter)

struct CholParams {
    double* A;          // matrix slab [E*n_pad][n_pad]
    double* Dinv;       // [E*n_pad][128]
    int64_t n_pad;
    int T;              // n_pad / 128
    int count;          // outputs factorised by this launch
    int outs[MAXG];     // their slab indices
    int* info;          // [E]   LAPACK info per output (0 = ok); must be zero on entry
    double* scal;       // [E][2] scal[2*o] receives log det
    int* sync;          // [CH_HDR + count*(2T+8)]: ticket, then per output prog[T][2] and dprog (zeroed per launch)
};

enum { TK_DIAG = 0, TK_D = 1, TK_ROW = 2 };

struct TileId {
    int kind, lo, i, p, j;
};

// ticket -> tile, with look-ahead on the critical chain.  The first 3*count tickets are [2*count DIAG(0)][count D(0)]; then
// block j = 0 .. T-2 holds count*(2T+1-2j) tickets:
//     [2*count ROW(j+1, p, j)] [2*count DIAG(j+1, p)] [count D(j+1)] [count * 2*(T-2-j) ROW(i, p, j), i = j+2 .. T-1]
// i.e. the tiles the NEXT column's diagonal block waits for, that diagonal block and its factorisation are drawn before the
// bulk of column j's row tiles.  (In plain column order DIAG(j+1) sat behind every ROW tile of column j: with a few matrices
// per launch -- C3 on 8 GPUs: 4 -- that is more tiles than SMs, and the chain DIAG -> D -> ROW stalled for a wave of row
// tiles in every column.)  Still a topological order: DIAG(j+1, p) needs ROW(j+1, p, k <= j), all drawn before it.
__host__ __device__ __forceinline__ TileId decode_ticket(int t, int T, int count) {
    TileId id;
    if (t < 3 * count) {
        id.j = 0; id.i = 0;
        if (t < 2 * count) {
            id.kind = TK_DIAG; id.lo = t >> 1; id.p = t & 1;
        } else {
            id.kind = TK_D; id.lo = t - 2 * count; id.p = 0;
        }
        return id;
    }
    t -= 3 * count;
    const int x = t / count;
    int j = (int)((double)(T + 1) - sqrt((double)(T + 1) * (double)(T + 1) - (double)x));
    if (j < 0) j = 0;
    if (j > T - 2) j = T - 2;
    while (j > 0 && j * (2 * T + 2 - j) > x) j--;
    while (j < T - 2 && (j + 1) * (2 * T + 2 - (j + 1)) <= x) j++;
    int u = t - count * j * (2 * T + 2 - j);
    if (u < 2 * count) {
        id.kind = TK_ROW; id.lo = u >> 1; id.p = u & 1; id.i = j + 1; id.j = j;
    } else if (u < 4 * count) {
        u -= 2 * count;
        id.kind = TK_DIAG; id.lo = u >> 1; id.p = u & 1; id.i = j + 1; id.j = j + 1;
    } else if (u < 5 * count) {
        id.kind = TK_D; id.lo = u - 4 * count; id.p = 0; id.i = j + 1; id.j = j + 1;
    } else {
        u -= 5 * count;
        const int per = 2 * (T - 2 - j);
        id.kind = TK_ROW; id.lo = u / per;
        const int w = u - id.lo * per;
        id.i = j + 2 + (w >> 1); id.p = w & 1; id.j = j;
    }
    return id;
}

__global__ void __launch_bounds__(CholCfg::THREADS, 1)
chol_dataflow_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ CUtensorMap tmD, const CholParams p) {
    using Cfg = CholCfg;
    constexpr int NS = Cfg::NS, BN = Cfg::BN, NT = Cfg::NT;
    extern __shared__ __align__(128) unsigned char chol_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(chol_smem_raw) + 127) & ~uintptr_t(127));
    double* VS = reinterpret_cast<double*>(base + NS * Cfg::STAGE_BYTES);  // [128/8][BN][8]
    uint64_t* full = reinterpret_cast<uint64_t*>(base + Cfg::REGION_BYTES);
    uint64_t* empty = full + NS;
    uint64_t* ks_full = empty + NS;
    uint64_t* vs_free = ks_full + 1;
    uint64_t* tq_full = vs_free + 1;   // [2]
    uint64_t* tq_empty = tq_full + 2;  // [2]
    uint64_t* d_done = tq_empty + 2;   // consumers are done with the borrowed smem of a D tile
    int* tq = reinterpret_cast<int*>(d_done + 1);  // [2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int total = p.count * T * (T + 2);
    const int blk = 2 * T + 8;   // ints per output in the progress area

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::NCW);
        }
        mbar_init(ks_full, 1);
        mbar_init(vs_free, Cfg::NCW);
        for (int s = 0; s < 2; s++) {
            mbar_init(&tq_full[s], 1);
            mbar_init(&tq_empty[s], Cfg::NCW);
        }
        mbar_init(d_done, Cfg::NCW);
        fence_mbar_init();
    }
    __syncthreads();

    constexpr int NCH = NB / KC;  // chunks per 128-wide K block
    constexpr int SKIP = 1 << 30;
    if (warp >= Cfg::NCW) {
        // =========================== ticket + TMA producer ===========================
        reg_dealloc<56>();
        if (warp == Cfg::NCW && elect_one_sync()) {
            prefetch_tmap(&tmL);
            prefetch_tmap(&tmW);
            prefetch_tmap(&tmD);
            PipeState<NS> ps;
            int seq = 0, dseq = 0;
            for (int nq = 0;; nq++) {
                const int slot = nq & 1;
                mbar_wait(&tq_empty[slot], (uint32_t)(((nq >> 1) & 1) ^ 1));
                const int t = atomicAdd(p.sync, 1);
                if (t >= total) {
                    tq[slot] = -1;
                    mbar_arrive(&tq_full[slot]);
                    break;
                }
                const TileId id = decode_ticket(t, T, p.count);
                const int o = p.outs[id.lo];
                int* prog = p.sync + CH_HDR + id.lo * blk;   // prog[i*2 + p]
                int* dprog = prog + 2 * T;
                const int j = id.j;
                if (id.kind == TK_D) {
                    // both halves of R_jj are in place; consumers factor it with plain (L2) loads
                    wait_counter(prog + 2 * j, Cfg::NCW * (j + 1));
                    wait_counter(prog + 2 * j + 1, Cfg::NCW * (j + 1));
                    const bool skip = ld_acquire_gpu(p.info + o) != 0;
                    tq[slot] = t | (skip ? SKIP : 0);
                    mbar_arrive(&tq_full[slot]);
                    // the D tile borrows the ring and the staging buffer: nothing may be in flight into them
                    mbar_wait(d_done, (uint32_t)(dseq & 1));
                    dseq++;
                    continue;
                }
                const bool skip = ld_acquire_gpu(p.info + o) != 0;   // a failed output only bumps its counters
                tq[slot] = t | (skip ? SKIP : 0);
                mbar_arrive(&tq_full[slot]);
                if (skip) continue;
                const int rb = (int)(o * p.n_pad);
                const int wrow = rb + id.i * NB + id.p * BN;   // the tile's 64 rows
                const int lrow = rb + j * NB;                  // block row j (L_jk tiles, Dinv_jj)
                int* prog_own = prog + 2 * id.i + id.p;
                bool vs_loaded = false;
                int issued = 0;
                auto load_vs = [&]() {
                    if (seq > 0) mbar_wait(vs_free, (uint32_t)((seq - 1) & 1));
                    mbar_arrive_expect_tx(ks_full, Cfg::VS_BYTES);
                    for (int ch = 0; ch < NCH; ch++)
                        tma_load_3d(VS + ch * (KC / 8) * BN * 8, &tmW, 0, wrow, j * (NB / 8) + ch * (KC / 8), ks_full);
                    vs_loaded = true;
                };
                // operands of block k: L_jk (both row halves of block row j) and the tile's own L_(i,p),k
                bool all_ready = (j == 0) || (ld_acquire_gpu(prog_own) >= Cfg::NCW * j &&
                                              ld_acquire_gpu(prog + 2 * j) >= Cfg::NCW * j &&
                                              ld_acquire_gpu(prog + 2 * j + 1) >= Cfg::NCW * j);
                if (all_ready) fence_proxy_async();
                for (int k = 0; k < j; k++) {
                    if (!all_ready) {
                        wait_counter(prog_own, Cfg::NCW * (k + 1));
                        wait_counter(prog + 2 * j, Cfg::NCW * (k + 1));
                        wait_counter(prog + 2 * j + 1, Cfg::NCW * (k + 1));
                        fence_proxy_async();
                    }
                    for (int ch = 0; ch < NCH; ch++) {
                        if (!vs_loaded && issued == NS - 1) load_vs();
                        mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                        unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ps.stage], Cfg::STAGE_BYTES);
                        const int kout = k * (NB / 8) + ch * (KC / 8);
                        tma_load_3d(st, &tmL, 0, lrow, kout, &full[ps.stage]);
                        tma_load_3d(st + Cfg::A_BYTES, &tmW, 0, wrow, kout, &full[ps.stage]);
                        ps.advance();
                        issued++;
                    }
                }
                if (!vs_loaded) load_vs();
                if (id.kind == TK_ROW) {
                    // inv(L_jj): published by the D tile of column j
                    wait_counter(dprog, j + 1);
                    fence_proxy_async();
                    for (int ch = 0; ch < NCH; ch++) {
                        mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                        unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ps.stage], Cfg::A_BYTES);
                        tma_load_3d(st, &tmD, 0, lrow, ch * (KC / 8), &full[ps.stage]);
                        ps.advance();
                    }
                }
                seq++;
            }
        }
        return;
    }

    // =========================== DMMA consumers ===========================
    reg_alloc<224>();
    const int ctid = threadIdx.x;   // 0..255
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, t4 = lane & 3;
    const int arow0 = wm * 32, bcol0 = wn * 8 * NT;

    PipeState<NS> ps;
    int seq = 0;
    for (int nq = 0;; nq++) {
        const int slot = nq & 1;
        mbar_wait(&tq_full[slot], (uint32_t)((nq >> 1) & 1));
        const int tword = tq[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&tq_empty[slot]);
        if (tword < 0) break;
        const bool skip = (tword & SKIP) != 0;
        const TileId id = decode_ticket(tword & (SKIP - 1), T, p.count);
        const int o = p.outs[id.lo];
        int* prog = p.sync + CH_HDR + id.lo * blk;
        int* dprog = prog + 2 * T;
        const int j = id.j;
        const int64_t rb = (int64_t)o * p.n_pad;

        if (id.kind == TK_D) {
            // the factorisation borrows the ring and the staging buffer: every consumer warp must be done reading them
            // for the previous tile (its last MMA stage, a DIAG tile's write-back) before the first store lands there
            named_bar_sync(1, Cfg::NCW * 32);
#ifndef CHOL_TRACE
            const int trs = -1;
#else
            __shared__ int tr_slot;
            if (ctid == 0) tr_slot = atomicAdd(&chol_trace_count, 1);
            named_bar_sync(1, Cfg::NCW * 32);
            const int trs = tr_slot;
            if (ctid == 0) { CH_STAMP(trs, 0); if (trs < CH_TRACE_N) chol_trace_buf[trs * CH_TRACE_EV + 7] = (unsigned long long)j;
                             unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); chol_trace_cur[sm_ & 255] = trs; }
            named_bar_sync(1, Cfg::NCW * 32);
#endif
            chol_d_tile<Cfg::NCW * 32>(base, ctid, p.A, p.Dinv, p.n_pad, rb, j, o, p.info, p.scal, dprog, skip, trs);
            if (lane == 0) mbar_arrive(d_done);
            continue;
        }
        if (skip) {
            __syncwarp();
            if (lane == 0) red_release_gpu_add(prog + 2 * id.i + id.p, 1);
            continue;
        }

        const int64_t wrow = rb + (int64_t)id.i * NB + id.p * BN;
        double acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.0;

        // acc[c][r] = sum_{k<j} L_jk[c,:] . L_(i,p),k[r,:]      (c: column inside block j, r: panel row)
        for (int c = 0; c < j * NCH; c++) {
            mbar_wait(&full[ps.stage], ps.phase);
            const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
            const double* Bs = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES + Cfg::A_BYTES);
            mma_stage<2, NT, KC>(acc, As, NB, arow0, Bs, BN, bcol0, g, t4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance();
        }

        // R = A_tile - acc, in place in VS (element (k = column c of block j, n = panel row r) at VS[c/8][r][c%8])
        mbar_wait(ks_full, (uint32_t)(seq & 1));
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int r = arow0 + mt * 16 + g + ((e >> 1) << 3);
                    const int c = bcol0 + nt * 8 + 2 * t4 + (e & 1);
                    double* qv = VS + ((size_t)((r >> 3) * BN + c) * 8 + (r & 7));
                    *qv = *qv - acc[mt][nt][e];
                    acc[mt][nt][e] = 0.0;
                }
        named_bar_sync(1, Cfg::NCW * 32);

        if (id.kind == TK_DIAG) {
            // write R back in place: 64 B segments (8 columns of one row) per thread
            for (int idx = ctid; idx < (NB / 8) * BN; idx += Cfg::NCW * 32) {
                const int k8 = idx / BN, r = idx - k8 * BN;
                const double2* src = reinterpret_cast<const double2*>(VS + (size_t)idx * 8);
                double2* dst = reinterpret_cast<double2*>(p.A + (wrow + r) * p.n_pad + (int64_t)j * NB + k8 * 8);
                dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
            }
            fence_proxy_async();       // the warp's generic-proxy writes into the staging tile -> the TMA load that overwrites it
            __syncwarp();
            if (lane == 0) mbar_arrive(vs_free);
            seq++;
        } else {
            // L_ij^T = inv(L_jj) * R^T
            for (int ch = 0; ch < NCH; ch++) {
                mbar_wait(&full[ps.stage], ps.phase);
                const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
                mma_stage<2, NT, KC>(acc, As, NB, arow0, VS + ch * (KC / 8) * BN * 8, BN, bcol0, g, t4);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[ps.stage]);
                ps.advance();
            }
            fence_proxy_async();       // the warp's generic-proxy writes into the staging tile -> the TMA load that overwrites it
            __syncwarp();
            if (lane == 0) mbar_arrive(vs_free);
            seq++;
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e1 = 0; e1 < 2; e1++) {
                    const int c = bcol0 + nt * 8 + 2 * t4 + e1;   // panel row
                    double* wr = p.A + (wrow + c) * p.n_pad + (int64_t)j * NB + arow0 + g;
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        wr[mt * 16] = acc[mt][nt][e1];
                        wr[mt * 16 + 8] = acc[mt][nt][2 + e1];
                    }
                }
        }
        // publish: generic-proxy global writes -> gpu scope -> async-proxy (TMA) readers of other CTAs
        __threadfence();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) red_release_gpu_add(prog + 2 * id.i + id.p, 1);
    }
}

// ------------------------------------------------------------------------------------------
// the same factorisation with the trailing updates on the int8 tensor cores (tcgen05)
// ------------------------------------------------------------------------------------------
// Same tiles, same ticket order, same progress counters and the same D tile as chol_dataflow_kernel; what changes is where the
// O(n^3) part -- the history products sum_{k<j} L_jk L_(i,p),k^T of the ROW and DIAG tiles -- is evaluated: as exact integer
// GEMM on S = 8 signed 7-bit planes per operand (tcgen05.mma kind::i8, s32 accumulators in TMEM; 36 plane pairs t + u <= 9,
// products resolved to 2^-63 of the squared scale), the error-free splitting of the predict TRSM (trsm_i8.cu) with one more
// digit (profiles/r01_ozaki_chol_study.txt: 8 digits are indistinguishable from the FP64 blocked factorisation on the
// ill-conditioned cases, 7 are not).  A ROW tile is exactly a tile of that TRSM whose right-hand side is the panel of L itself:
//
//     T = A_(i,p),j^T - 2^(2e) sum_w 2^-7w acc_w          (TMEM -> registers -> the K-blocked buffer A_(i,p),j was TMA-loaded into)
//     L_(i,p),j^T = inv(L_jj) T                           (FP64 DMMA; inv(L_jj) streamed in 8-column slabs, upper triangle skipped)
//
// and its result leaves the SM twice: as FP64 into the matrix (every other kernel of the library reads L in FP64) and as 8
// digit planes into Lq -- the operands of the later columns of this factorisation AND the planes of L the predict TRSM needs
// (no slicing pass after the fit).  |L_rc| <= sqrt(K_rr) = sqrt(sigma2 + nugget) gives the one power-of-two scale per output.
// A DIAG tile stops after T (= its 64 rows of R_jj, written back in FP64); the D tile factors R_jj in FP64 as before.
// Per CTA (384 threads): warps 0-7 consumers (TMEM drain, FP64 epilogue, D tiles), warp 8 MMA issuer, warp 9 tickets + plane
// loader (progress-counter waits, cp.async.bulk into a 3-stage ring), warp 10 loader of the A tile and of inv(L_jj).
constexpr int CHOL_I8_PLANES = 8;

template <int S>
struct CholI8Cfg {
    static constexpr int NS = 3;
    static constexpr int ASTAGE = S * I8_APLANE;            // planes of L_jk, one K = 32 step (128 rows)
    static constexpr int BSTAGE = S * I8_BPLANE;            // planes of the tile's own 64 rows
    static constexpr int STAGE = ASTAGE + BSTAGE;
    static constexpr int OFF_T = NS * STAGE;                // T (FP64, K-blocked [16][64][8]); later the plane image [4][S][2048]
    static constexpr int DSLAB = NB * 8 * 8;                // one 8-column K slab of inv(L_jj): 8 KB
    static constexpr int OFF_D = OFF_T + NB * I8_BN * 8;
    static constexpr int OFF_BAR = OFF_D + 2 * DSLAB;
    static constexpr int SMEM = OFF_BAR + 256 + 128;
    static constexpr int THREADS = (I8_NCW + 4) * 32;       // three role warps + one idle warp: setmaxnreg works on warpgroups
    static_assert(S <= I8_LP && S * I8_BN <= 512, "planes / TMEM columns");
    static_assert(4 * BSTAGE <= NB * I8_BN * 8, "the plane image reuses the T buffer");
    static_assert((int)sizeof(Potf2Smem) <= OFF_D, "the D tile borrows the ring and the T buffer");
    static_assert(SMEM <= 232448, "shared memory per CTA");
};

struct CholI8Params {
    CholParams c;
    int8_t* Lq;             // [E][lq_stride] planes of the strictly lower blocks of L
    int64_t lq_stride;
    int eS[MAXG];           // scale exponent per listed output: |L| 2^-eS <= 0.99
};

template <int S>
__global__ void __launch_bounds__(CholI8Cfg<S>::THREADS, 1)
chol_i8_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmD8, const CholI8Params pp) {
    using Cfg = CholI8Cfg<S>;
    constexpr int NS = Cfg::NS, STAGE = Cfg::STAGE, ASTAGE = Cfg::ASTAGE, BSTAGE = Cfg::BSTAGE, DSLAB = Cfg::DSLAB;
    constexpr int SKIP = 1 << 30;
    const CholParams& p = pp.c;
    extern __shared__ __align__(128) unsigned char chol_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(chol_smem_raw) + 127) & ~uintptr_t(127));
    double* Ts = reinterpret_cast<double*>(base + Cfg::OFF_T);                       // [16 slabs][64 rows of the panel][8]
    unsigned char* img = base + Cfg::OFF_T;                                          // [4 K steps][S planes][2048]
    unsigned char* dring = base + Cfg::OFF_D;
    uint64_t* full = reinterpret_cast<uint64_t*>(base + Cfg::OFF_BAR);               // [NS]
    uint64_t* empty = full + NS;                                                     // [NS]
    uint64_t* acc_full = empty + NS;
    uint64_t* acc_empty = acc_full + 1;
    uint64_t* d_full = acc_empty + 1;                                                // [2]
    uint64_t* d_empty = d_full + 2;                                                  // [2]
    uint64_t* ts_full = d_empty + 2;                                                 // the A tile has landed in the T buffer
    uint64_t* ts_free = ts_full + 1;                                                 // the T buffer (R written back / plane image read) is free
    uint64_t* tq_full = ts_free + 1;                                                 // [QN]
    uint64_t* tq_empty = tq_full + I8_QN;                                            // [QN]
    uint64_t* d_done = tq_empty + I8_QN;                                             // [2]: a D tile is done with the borrowed memory (one per loader warp)
    int* tq = reinterpret_cast<int*>(d_done + 2);                                    // [QN]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tq + I8_QN);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int T = p.T;
    const int total = p.count * T * (T + 2);
    const int blk = 2 * T + 8;   // ints per output in the progress area: prog[T][2] (blocks finished per row half), dprog

    if (warp == I8_NCW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, I8_NCW);
        for (int s = 0; s < 2; s++) {
            mbar_init(&d_full[s], 1);
            mbar_init(&d_empty[s], I8_NCW);
            mbar_init(&d_done[s], I8_NCW);
        }
        mbar_init(ts_full, 1);
        mbar_init(ts_free, 1);
        for (int s = 0; s < I8_QN; s++) {
            mbar_init(&tq_full[s], 1);
            mbar_init(&tq_empty[s], I8_NCW + 2);     // consumer warps + MMA issuer + A-tile / inv(L_jj) loader
        }
        fence_mbar_init();
    }
    i8_fence_before();
    __syncthreads();
    i8_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp >= I8_NCW) {
        reg_dealloc<72>();
        if (warp == I8_NCW + 1) {
            // ================================ tickets + plane loader ================================
            if (i8_elect_one()) {
                int it = 0, dseq = 0;
                for (int nq = 0;; nq++) {
                    const int slot = nq % I8_QN;
                    i8_wait(&tq_empty[slot], (uint32_t)(((nq / I8_QN) & 1) ^ 1));
                    const int t = atomicAdd(p.sync, 1);
                    if (t >= total) {
                        tq[slot] = -1;
                        mbar_arrive(&tq_full[slot]);
                        break;
                    }
                    const TileId id = decode_ticket(t, T, p.count);
                    const int o = p.outs[id.lo];
                    int* prog = p.sync + CH_HDR + id.lo * blk;
                    const int j = id.j;
                    if (id.kind == TK_D) {
                        // both halves of R_jj are in place (ROW tiles of the block row + its DIAG tile: j + 1 per half)
                        wait_counter(prog + 2 * j, j + 1);
                        wait_counter(prog + 2 * j + 1, j + 1);
                        const bool skip = ld_acquire_gpu(p.info + o) != 0;
                        tq[slot] = t | (skip ? SKIP : 0);
                        mbar_arrive(&tq_full[slot]);
                        // the D tile borrows the ring and the T buffer: nothing may be in flight into them
                        i8_wait(&d_done[0], (uint32_t)(dseq & 1));
                        dseq++;
                        continue;
                    }
                    const bool skip = ld_acquire_gpu(p.info + o) != 0;   // a failed output only bumps its counters
                    tq[slot] = t | (skip ? SKIP : 0);
                    mbar_arrive(&tq_full[slot]);
                    if (skip || j == 0) continue;
                    const int8_t* lq = pp.Lq + (size_t)o * pp.lq_stride;
                    const int8_t* a_src = lq + (size_t)(j * (j - 1) / 2) * I8_LBLOCK;                           // blocks (j, k)
                    const int8_t* b_src = lq + (size_t)(id.i * (id.i - 1) / 2) * I8_LBLOCK + id.p * I8_BPLANE;  // blocks (i, k), rows of half p
                    int* prog_own = prog + 2 * id.i + id.p;
                    const bool all_ready = ld_acquire_gpu(prog_own) >= j && ld_acquire_gpu(prog + 2 * j) >= j &&
                                           ld_acquire_gpu(prog + 2 * j + 1) >= j;
                    if (all_ready) fence_proxy_async();
                    for (int k = 0; k < j; k++) {
                        if (!all_ready) {
                            wait_counter(prog_own, k + 1);
                            wait_counter(prog + 2 * j, k + 1);
                            wait_counter(prog + 2 * j + 1, k + 1);
                            fence_proxy_async();
                        }
                        for (int s = 0; s < 4; s++, it++) {
                            const int rs = it % NS;
                            if (it >= NS) i8_wait(&empty[rs], (uint32_t)(((it / NS) - 1) & 1));
                            unsigned char* dst = base + rs * STAGE;
                            const size_t off = (size_t)(4 * k + s) * I8_LSTAGE;
                            mbar_arrive_expect_tx(&full[rs], STAGE);
                            i8_bulk_load(dst, a_src + off, ASTAGE, &full[rs]);
#pragma unroll
                            for (int tt = 0; tt < S; tt++)
                                i8_bulk_load(dst + ASTAGE + tt * I8_BPLANE, b_src + off + (size_t)tt * I8_APLANE, I8_BPLANE, &full[rs]);
                        }
                    }
                }
            }
        } else if (warp == I8_NCW + 2) {
            // ================================ A tile and inv(L_jj) loader ================================
            if (i8_elect_one()) {
                prefetch_tmap(&tmD8);
                prefetch_tmap(&tmW);
                int tseq = 0, dc = 0, dseq = 0;
                for (int nq = 0;; nq++) {
                    const int slot = nq % I8_QN;
                    i8_wait(&tq_full[slot], (uint32_t)((nq / I8_QN) & 1));
                    const int tw = tq[slot];
                    mbar_arrive(&tq_empty[slot]);
                    if (tw < 0) break;
                    const TileId id = decode_ticket(tw & (SKIP - 1), T, p.count);
                    if (id.kind == TK_D) {
                        i8_wait(&d_done[1], (uint32_t)(dseq & 1));
                        dseq++;
                        continue;
                    }
                    const int j = id.j;
                    if ((tw & SKIP) || (id.kind == TK_DIAG && j == 0)) continue;
                    const int o = p.outs[id.lo];
                    const int rb = (int)(o * p.n_pad);
                    const int wrow = rb + id.i * NB + id.p * I8_BN;
                    if (tseq > 0) i8_wait(ts_free, (uint32_t)((tseq - 1) & 1));
                    tseq++;
                    mbar_arrive_expect_tx(ts_full, NB * I8_BN * 8);
                    for (int c2 = 0; c2 < NB / KC; c2++)
                        tma_load_3d(Ts + c2 * (KC / 8) * I8_BN * 8, &tmW, 0, wrow, j * (NB / 8) + c2 * (KC / 8), ts_full);
                    if (id.kind == TK_ROW) {
                        const int* dprog = p.sync + CH_HDR + id.lo * blk + 2 * T;
                        wait_counter(dprog, j + 1);          // inv(L_jj): published by the D tile of column j
                        fence_proxy_async();
                        for (int ch = 0; ch < NB / 8; ch++, dc++) {
                            const int ds = dc & 1;
                            if (dc >= 2) i8_wait(&d_empty[ds], (uint32_t)(((dc >> 1) - 1) & 1));
                            mbar_arrive_expect_tx(&d_full[ds], DSLAB);
                            tma_load_3d(dring + ds * DSLAB, &tmD8, 0, rb + j * NB, ch, &d_full[ds]);
                        }
                    }
                }
            }
        } else if (warp == I8_NCW) {
            // ================================ MMA issuer (one elected thread) ================================
            if (i8_elect_one()) {
                int it = 0, k = 0;
                for (int nq = 0;; nq++) {
                    const int slot = nq % I8_QN;
                    i8_wait(&tq_full[slot], (uint32_t)((nq / I8_QN) & 1));
                    const int tw = tq[slot];
                    mbar_arrive(&tq_empty[slot]);
                    if (tw < 0) break;
                    if (tw & SKIP) continue;
                    const TileId id = decode_ticket(tw, T, p.count);
                    if (id.kind == TK_D || id.j == 0) continue;
                    if (k > 0) {          // the consumers must have drained the accumulators of the previous tile
                        i8_wait(acc_empty, (uint32_t)((k - 1) & 1));
                        i8_fence_after();
                    }
                    k++;
                    for (int st = 0; st < 4 * id.j; st++, it++) {
                        const int rs = it % NS;
                        i8_wait(&full[rs], (uint32_t)((it / NS) & 1));
                        i8_fence_after();
                        const uint32_t a0 = smem_u32(base + rs * STAGE), b0 = a0 + ASTAGE;
                        const uint64_t bd0 = i8_desc(b0), bd1 = i8_desc(b0 + 4 * I8_BPLANE);
#pragma unroll
                        for (int tt = 1; tt <= S; tt++) {
                            const int ncols = I8_BN * (S + 1 - tt);
                            const uint32_t accum = (st == 0 && tt == 1) ? 0u : 1u;
                            const uint32_t d0 = tmem + (uint32_t)(tt - 1) * I8_BN;
                            const uint64_t ad = i8_desc(a0 + (tt - 1) * I8_APLANE);
                            const int n1 = ncols > 256 ? 256 : ncols;
                            i8_mma(d0, ad, bd0, accum, i8_idesc(n1));
                            if (ncols > 256) i8_mma(d0 + 256, ad, bd1, accum, i8_idesc(ncols - 256));
                        }
                        i8_commit(&empty[rs]);
                    }
                    i8_commit(acc_full);
                }
            }
        }
    } else {
        // ================================ consumers ================================
        reg_alloc<216>();
        const int q4 = warp & 3, h = warp >> 2;
        const int r = q4 * 32 + lane;             // column inside block j = TMEM lane
        const int g = lane >> 2, t4 = lane & 3;   // DMMA fragment coordinates
        int k = 0, dc = 0, tseq = 0;
        for (int nq = 0;; nq++) {
            const int slot = nq % I8_QN;
            i8_wait(&tq_full[slot], (uint32_t)((nq / I8_QN) & 1));
            const int tw = tq[slot];
            __syncwarp();
            if (lane == 0) mbar_arrive(&tq_empty[slot]);
            if (tw < 0) break;
            const bool skip = (tw & SKIP) != 0;
            const TileId id = decode_ticket(tw & (SKIP - 1), T, p.count);
            const int o = p.outs[id.lo];
            int* prog = p.sync + CH_HDR + id.lo * blk;
            int* dprog = prog + 2 * T;
            const int j = id.j;
            const int64_t rb = (int64_t)o * p.n_pad;

            if (id.kind == TK_D) {
                // every consumer warp is done with the previous tile (the barrier that ends every tile), its MMAs have completed
                // (acc_full) and both loader warps are parked on d_done: the ring and the T buffer are free to borrow
                named_bar_sync(1, I8_NCW * 32);
                chol_d_tile<I8_NCW * 32>(base, tid, p.A, p.Dinv, p.n_pad, rb, j, o, p.info, p.scal, dprog, skip, -1);
                if (lane == 0) {
                    mbar_arrive(&d_done[0]);
                    mbar_arrive(&d_done[1]);
                }
                continue;
            }
            int* prog_own = prog + 2 * id.i + id.p;
            if (skip || (id.kind == TK_DIAG && j == 0)) {      // (R_00 = A_00 is in place)
                if (tid == 0) red_release_gpu_add(prog_own, 1);
                continue;
            }
            const int es = pp.eS[id.lo];
            const int64_t wrow = rb + (int64_t)id.i * NB + id.p * I8_BN;

            // ---- T = A_(i,p),j^T - 2^(2 es) sum_w 2^-7w acc_w, in place in the K-blocked buffer the A tile was loaded into ----
            if (j > 0) {
                long long hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 32; c++) hi[c] = lo[c] = 0;
                i8_wait(acc_full, (uint32_t)(k & 1));
                k++;
                i8_fence_after();
#pragma unroll
                for (int w = 0; w < S; w++) {
                    uint32_t v[32];
                    i8_tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(w * I8_BN + h * 32), v);
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        if (w < 4) hi[c] += (long long)(int32_t)v[c] << (I8_BITS * (3 - w));
                        else lo[c] += (long long)(int32_t)v[c] << (I8_BITS * (S - 1 - w));
                    }
                }
                i8_fence_before();
                double acc[32];
                {
                    const double whi = __longlong_as_double((long long)(1023 - I8_BITS * 5) << 52);         // accumulator 3: 2^-35
                    const double wlo = __longlong_as_double((long long)(1023 - I8_BITS * (S + 1)) << 52);   // accumulator S-1
#pragma unroll
                    for (int c = 0; c < 32; c++) acc[c] = fma((double)lo[c], wlo, (double)hi[c] * whi);
                }
                i8_wait(ts_full, (uint32_t)(tseq & 1));
                const double nscale = -ldexp(1.0, 2 * es);
                const int odd = (lane >> 3) & 1;
                double* trow = Ts + (size_t)(r >> 3) * I8_BN * 8 + (size_t)h * 32 * 8 + (r & 7);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    const double va = odd ? acc[c + 1] : acc[c], vb = odd ? acc[c] : acc[c + 1];
                    double* pa = trow + (c + odd) * 8;
                    double* pb = trow + (c + 1 - odd) * 8;
                    *pa = fma(va, nscale, *pa);
                    *pb = fma(vb, nscale, *pb);
                }
            } else {
                i8_wait(ts_full, (uint32_t)(tseq & 1));
            }
            tseq++;
            named_bar_sync(1, I8_NCW * 32);

            if (id.kind == TK_DIAG) {
                if (lane == 0) mbar_arrive(acc_empty);          // (j > 0 here)
                // R (64 rows of the diagonal block) back in place: 64-byte segments (8 columns of one row) per thread
                for (int idx = tid; idx < (NB / 8) * I8_BN; idx += I8_NCW * 32) {
                    const int k8 = idx / I8_BN, rr = idx - k8 * I8_BN;
                    const double2* src = reinterpret_cast<const double2*>(Ts + (size_t)idx * 8);
                    double2* dst = reinterpret_cast<double2*>(p.A + (wrow + rr) * p.n_pad + (int64_t)j * NB + k8 * 8);
                    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
                }
                __threadfence();
                fence_proxy_async();       // generic-proxy writes into the T buffer -> the TMA load of the next tile that overwrites it
                named_bar_sync(1, I8_NCW * 32);
                if (tid == 0) {
                    mbar_arrive(ts_free);
                    red_release_gpu_add(prog_own, 1);
                }
                continue;
            }

            // ---- L_(i,p),j^T = inv(L_jj) T : warp w owns the panel rows 8 w .. 8 w + 7 and all 128 columns of block j ----
            double vf[8][4];
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int e = 0; e < 4; e++) vf[mt][e] = 0.0;
#pragma unroll
            for (int ch = 0; ch < NB / 8; ch++, dc++) {
                const int ds = dc & 1;
                i8_wait(&d_full[ds], (uint32_t)((dc >> 1) & 1));
                const double* As = reinterpret_cast<const double*>(dring + ds * DSLAB);
                const double2 bf = *reinterpret_cast<const double2*>(Ts + ((size_t)ch * I8_BN + warp * 8 + g) * 8 + 2 * t4);
#pragma unroll
                for (int mt = ch >> 1; mt < 8; mt++) {         // inv(L_jj) is lower triangular: rows 16 mt .. need columns <= 16 mt + 15
                    const double* ap = As + ((size_t)(mt * 16 + g) * 8 + 2 * t4);
                    const double2 a0 = *reinterpret_cast<const double2*>(ap);
                    const double2 a1 = *reinterpret_cast<const double2*>(ap + 64);
                    dmma_16x8x8(vf[mt], a0.x, a1.x, a0.y, a1.y, bf.x, bf.y);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[ds]);
            }
            named_bar_sync(1, I8_NCW * 32);       // every warp is done reading T: the buffer becomes the plane image
            // FP64 and int8 MMAs share one datapath (tools/probe_concurrency.cu): the accumulators go back to the MMA warp only
            // after the FP64 part, as in the predict TRSM
            if (j > 0 && lane == 0) mbar_arrive(acc_empty);

            // ---- FP64 copy of the tile: element (column c of block j, panel row rr) -> A[wrow + rr][j NB + c] ----
            // vf[mt][e]: c = 16 mt + g (+ 8 for e >= 2), rr = 8 warp + 2 t4 (+ 1 for odd e)
#pragma unroll
            for (int e1 = 0; e1 < 2; e1++) {
                double* wr = p.A + (wrow + warp * 8 + 2 * t4 + e1) * p.n_pad + (int64_t)j * NB + g;
#pragma unroll
                for (int mt = 0; mt < 8; mt++) {
                    wr[mt * 16] = vf[mt][e1];
                    wr[mt * 16 + 8] = vf[mt][2 + e1];
                }
            }
            // ---- its S digit planes: byte (column c, panel row rr, plane tt) of the image at K step c / 32, then plane, then
            //      (rr / 8) 256 + (c / 16 % 2) 128 + (rr % 8) 16 + c % 16 (4 x 4 byte transposes by shuffle, see trsm_i8.cu) ----
            {
                const int jj = g & 3;
                unsigned char* ib = img + warp * 256 + (2 * t4 + (jj & 1)) * 16 + (g >> 2) * 4 + (jj >> 1) * 8;
#pragma unroll
                for (int mt = 0; mt < 8; mt++) {
                    int8_t dig[4][S];
#pragma unroll
                    for (int e = 0; e < 4; e++) i8_digits_int<S>(vf[mt][e], es, dig[e]);
                    unsigned char* dst = ib + (size_t)(mt >> 1) * BSTAGE + (mt & 1) * 128;
#pragma unroll
                    for (int tt = 0; tt < S; tt++) {
                        uint32_t x = (uint32_t)(uint8_t)dig[0][tt] | ((uint32_t)(uint8_t)dig[1][tt] << 8) |
                                     ((uint32_t)(uint8_t)dig[2][tt] << 16) | ((uint32_t)(uint8_t)dig[3][tt] << 24);
                        uint32_t y = __shfl_xor_sync(0xffffffffu, x, 4);
                        x = __byte_perm(x, y, (jj & 1) ? 0x3715 : 0x6240);
                        y = __shfl_xor_sync(0xffffffffu, x, 8);
                        x = __byte_perm(x, y, (jj & 2) ? 0x3276 : 0x5410);
                        *reinterpret_cast<uint32_t*>(dst + tt * I8_BPLANE) = x;
                    }
                }
                fence_proxy_async();              // generic-proxy writes of the image -> the bulk stores' async-proxy reads
            }
            __threadfence();                      // the FP64 copy
            named_bar_sync(1, I8_NCW * 32);
            if (tid == 0) {
                // the 64 rows of half p of every (K step, plane) of block (i, j): 2 KB pieces of the 4 KB planes
                int8_t* dstb = pp.Lq + (size_t)o * pp.lq_stride + (size_t)(id.i * (id.i - 1) / 2 + j) * I8_LBLOCK + id.p * I8_BPLANE;
#pragma unroll 1
                for (int s4 = 0; s4 < 4; s4++)
#pragma unroll 1
                    for (int tt = 0; tt < S; tt++)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(dstb + (size_t)s4 * I8_LSTAGE + (size_t)tt * I8_APLANE),
                                       "r"(smem_u32(img + (s4 * S + tt) * I8_BPLANE)), "r"(I8_BPLANE) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the image has been read: the next A tile may land
                mbar_arrive(ts_free);
                i8_bulk_store_wait();                                             // the planes are in global memory
                fence_proxy_async();
                __threadfence();
                red_release_gpu_add(prog_own, 1);
            }
        }
    }
    i8_fence_before();
    __syncthreads();
    if (warp == I8_NCW) {
        i8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// the schedule, for the host (tests check that it is a permutation of the tiles in a topological order of their dependencies:
// that order is what makes the spin-waits of the persistent kernel deadlock-free)
void chol_ticket(int t, int T, int count, int out[5]) {
    const TileId id = decode_ticket(t, T, count);
    out[0] = id.kind; out[1] = id.lo; out[2] = id.i; out[3] = id.p; out[4] = id.j;
}

// kernel attributes are per device: called once per device by mogp_create (api.cu keeps the per-device flag)
int chol_init() {
    if (cudaFuncSetAttribute(chol_dataflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CholCfg::SMEM_BYTES) !=
            cudaSuccess ||
        cudaFuncSetAttribute(chol_i8_kernel<CHOL_I8_PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             CholI8Cfg<CHOL_I8_PLANES>::SMEM) != cudaSuccess)
        return 1;
    return 0;
}

int chol_make_maps(CholMaps* maps, double* A_slab, double* Dinv_slab, int64_t total_rows, int64_t n_pad) {
    if (make_kblocked_tmap(&maps->a128, A_slab, total_rows, n_pad, 128)) return 1;
    if (make_kblocked_tmap(&maps->a64, A_slab, total_rows, n_pad, 64)) return 1;
    if (make_kblocked_tmap(&maps->d128, Dinv_slab, total_rows, NB, 128)) return 1;
    if (make_kblocked_tmap(&maps->d8, Dinv_slab, total_rows, NB, 128, 1)) return 1;
    return 0;
}

size_t chol_sync_bytes(int count, int T) { return sizeof(int) * ((size_t)CH_HDR + (size_t)count * (2 * T + 8)); }

// Enqueue the factorisation of `count` outputs (slab indices outs[]) on `st` as one launch.  info[o] and
// scal[2*o] of those outputs must have been zeroed.  Returns the number of kernels launched (negative on error).
int chol_factor_batch(const CholMaps& maps, double* A_slab, double* Dinv_slab, const int* outs, int count,
                      int64_t n_pad, int* info, double* scal, int* sync, int n_sms, cudaStream_t st) {
    if (count < 1 || count > MAXG) return -1;
    CholParams p{};
    p.A = A_slab; p.Dinv = Dinv_slab; p.n_pad = n_pad; p.T = (int)(n_pad / NB); p.count = count;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    p.info = info; p.scal = scal; p.sync = sync;
    if (cudaMemsetAsync(sync, 0, chol_sync_bytes(count, p.T), st) != cudaSuccess) return -1;
    const int64_t tiles = (int64_t)count * p.T * (p.T + 2);
    const unsigned grid = (unsigned)(tiles < n_sms ? tiles : n_sms);
    chol_dataflow_kernel<<<grid, CholCfg::THREADS, CholCfg::SMEM_BYTES, st>>>(maps.a128, maps.a64, maps.d128, p);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 1;
}

// The same factorisation with the history products on the int8 tensor cores; also leaves the planes of the strictly lower
// blocks of L in Lq (exps[k]: i8_scale_exponent of outs[k] with the nugget of this attempt).
int chol_i8_factor_batch(const CholMaps& maps, double* A_slab, double* Dinv_slab, const int* outs, const int* exps, int count,
                         int64_t n_pad, int8_t* Lq, int64_t lq_stride, int* info, double* scal, int* sync, int n_sms,
                         cudaStream_t st) {
    if (count < 1 || count > MAXG) return -1;
    CholI8Params pp{};
    CholParams& p = pp.c;
    p.A = A_slab; p.Dinv = Dinv_slab; p.n_pad = n_pad; p.T = (int)(n_pad / NB); p.count = count;
    for (int i = 0; i < count; i++) {
        p.outs[i] = outs[i];
        pp.eS[i] = exps[i];
    }
    p.info = info; p.scal = scal; p.sync = sync;
    pp.Lq = Lq; pp.lq_stride = lq_stride;
    if (cudaMemsetAsync(sync, 0, chol_sync_bytes(count, p.T), st) != cudaSuccess) return -1;
    const int64_t tiles = (int64_t)count * p.T * (p.T + 2);
    const unsigned grid = (unsigned)(tiles < n_sms ? tiles : n_sms);
    chol_i8_kernel<CHOL_I8_PLANES><<<grid, CholI8Cfg<CHOL_I8_PLANES>::THREADS, CholI8Cfg<CHOL_I8_PLANES>::SMEM, st>>>(maps.a64, maps.d8, pp);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 1;
}
int chol_i8_planes() { return CHOL_I8_PLANES; }

}  // namespace mogp

#ifdef CHOL_TRACE
extern "C" int mogp_debug_chol_trace(unsigned long long* out, int n_words, int reset) {
    int cnt = 0;
    cudaMemcpyFromSymbol(&cnt, mogp::chol_trace_count, sizeof(int));
    const int words = mogp::CH_TRACE_N * mogp::CH_TRACE_EV;
    if (out) cudaMemcpyFromSymbol(out, mogp::chol_trace_buf, sizeof(unsigned long long) * (size_t)(n_words < words ? n_words : words));
    if (reset) { int z = 0; cudaMemcpyToSymbol(mogp::chol_trace_count, &z, sizeof(int)); }
    return cnt;
}
extern "C" int mogp_debug_chol_trace2(unsigned long long* out, int n_words) {
    return cudaMemcpyFromSymbol(out, mogp::chol_trace_buf2, sizeof(unsigned long long) * (size_t)n_words) == cudaSuccess ? 0 : 1;
}
#endif
