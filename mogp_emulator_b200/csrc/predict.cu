// Predictive variance for libmogp_b200:  V = L^-1 K*  (in place, test-major) and
// var_c = max(sigma2 [+ nugget] - ||V_c||^2, 0).
//
// Replaces the reference's explicit inverse + GEMM + batched dot (cusolverDnDpotrs with n right-hand
// sides, cublasDgemm, cublasDgemmStridedBatched; mogp_gpu/src/densegp_gpu.hpp:576-582, 374-396).  Values
// follow the CPU reference GaussianProcess.predict (GaussianProcess.py:896-920): Kinv.solve(Ktest) followed by
// sum(Ktest * Kinv_Ktest) equals ||L^-1 k*||^2, obtained here with one triangular solve (n^2 m flops instead
// of 2 n^2 m) and clipped at zero.
//
// One CTA owns a panel of BN test points and walks the block rows of L (left-looking blocked forward
// substitution):   V_i = inv(L_ii) * (K*_i - sum_{j<i} L_ij V_j).
// Both products are FP64 tensor-pipe GEMMs (mma.sync.m16n8k8.f64) in TN form: the L / inv(L_ii) tiles and the
// already-solved V_j tiles are streamed by a TMA producer warp through an mbarrier ring of K-blocked
// stages; the right-hand side tile K*_i is TMA-loaded into a resident staging buffer that doubles as the
// B operand of the diagonal product.  Column norms accumulate in registers across the whole walk.
#include "common.cuh"
#include "kernels.h"

namespace mogp {

template <int NT>
struct PredCfg {
    static constexpr int BM = NB;
    static constexpr int BN = 16 * NT;
    static constexpr int NCW = 8;
    static constexpr int THREADS = (NCW + 4) * 32;  // two consumer warpgroups + producer warpgroup
    static constexpr int NS = 4;
    static_assert(NT == 2 || NT == 4, "panel width 32 or 64 (128 does not fit next to the staging buffer)");
    static constexpr int A_BYTES = BM * KC * 8;
    static constexpr int B_BYTES = BN * KC * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int VS_BYTES = NB * BN * 8;
    static constexpr int BAR_BYTES = (2 * NS + 2 + 4) * 8 + 16;   // ring, ks_full, vs_free, ticket queue (+ 2 ticket words)
    static constexpr int SMEM_BYTES = NS * STAGE_BYTES + VS_BYTES + BAR_BYTES + 4 * BN * 8 + 128;
};

constexpr int SYNC_HDR = 32;   // ints in front of the flag array (word 0 = ticket counter)

struct PredParams {
    double* W;              // workspace slab [count][w_stride][n_pad]
    int64_t w_stride;       // rows per output in the slab
    int64_t n_pad;
    int64_t m;              // real number of test points
    int T;                  // n_pad / 128
    int panels;             // panels of BN test points per output
    int count;              // outputs in this launch
    int outs[MAXG];         // global output index of local output k
    const double* hyper;    // [E][hyper_stride]
    int hyper_stride;       // d + 2
    int d;
    int include_nugget;
    int tri_rhs;            // right-hand side is the identity: panel c0 starts its walk at block row c0/128
    int no_clip;            // write sigma2 [+ nugget] - ||V_c||^2 without the max(., 0) (the caller adds the mean-function term first)
    int keep_v;             // also store the last block row of V (full predictive covariance needs all of V)
    int diag_only;          // empty history: W_i <- inv(L_ii) W_i for every block row; no norms, no variance.  (The first K~* pass of
                            // trsm_i8.cu, 16.1 ms at C3; superseded by i8_ktilde_kernel at 10.1 ms and kept for that comparison.)
    double* var;            // result rows: var of output o at var + o*var_stride
    int64_t var_stride;
    int* sync;              // [SYNC_HDR + count*panels*T]: ticket counter, then one ready-flag per tile (zeroed per launch)
    double* normacc;        // [count][w_stride] running ||V_c||^2 over the block rows done so far
};

// Dataflow (tile-ticket) blocked forward substitution.
//
// Unit of work = tile (block row i, panel of BN test points, output):   V_i = inv(L_ii) (K*_i - sum_{j<i} L_ij V_j).
// A persistent grid (one CTA per SM) draws tiles from a global ticket counter in block-row-major order, so every
// tile a CTA may have to wait for (same panel, smaller block row) has a smaller ticket and is already held by a
// running CTA: the ready-flag spin below cannot deadlock, and the machine stays full for any (m, outputs) --
// few right-hand sides (C4: 8 panels) are parallel over block rows, many (C3) over panels, with no wave
// quantisation.  The products are FP64 tensor-pipe GEMMs in TN form: L / inv(L_ii) tiles and the solved V_j tiles
// (written by other CTAs, published with fence + red.release, read back by TMA after ld.acquire + proxy fence)
// stream through an mbarrier ring; K*_i lands in a resident staging buffer that is the B operand of the diagonal
// product.  Column norms ride along in a per-test-point accumulator that the chain of tiles of one panel updates
// strictly in order (deterministic); the last block row writes the clipped variance.
template <int NT>
__global__ void __launch_bounds__(PredCfg<NT>::THREADS, 1)
predict_trsm_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmD,
                    const __grid_constant__ CUtensorMap tmW, const PredParams p) {
    using Cfg = PredCfg<NT>;
    constexpr int NS = Cfg::NS;
    constexpr int BN = Cfg::BN;
    extern __shared__ __align__(128) unsigned char pred_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pred_smem_raw) + 127) & ~uintptr_t(127));
    double* VS = reinterpret_cast<double*>(base + NS * Cfg::STAGE_BYTES);  // [128/8][BN][8]
    uint64_t* full = reinterpret_cast<uint64_t*>(base + NS * Cfg::STAGE_BYTES + Cfg::VS_BYTES);
    uint64_t* empty = full + NS;
    uint64_t* ks_full = empty + NS;
    uint64_t* vs_free = ks_full + 1;
    uint64_t* tq_full = vs_free + 1;   // [2]
    uint64_t* tq_empty = tq_full + 2;  // [2]
    int* tq = reinterpret_cast<int*>(tq_empty + 2);  // [2] ticket handed from the producer to the consumers
    double* nred = reinterpret_cast<double*>(tq + 4);  // [4][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int per_row = p.count * p.panels;
    const int total = T * per_row;
    int* flags = p.sync + SYNC_HDR;

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
        fence_mbar_init();
    }
    __syncthreads();

    constexpr int NCH = NB / KC;  // chunks per 128-wide K block
    if (warp >= Cfg::NCW) {
        // =========================== ticket + TMA producer ===========================
        reg_dealloc<56>();
        if (warp == Cfg::NCW && elect_one_sync()) {
            prefetch_tmap(&tmL);
            prefetch_tmap(&tmD);
            prefetch_tmap(&tmW);
            PipeState<NS> ps;
            int seq = 0;
            for (int nq = 0;; nq++) {
                const int slot = nq & 1;
                mbar_wait(&tq_empty[slot], (uint32_t)(((nq >> 1) & 1) ^ 1));
                int t, i = 0, q = 0, pnl = 0, i0 = 0;
                for (;;) {
                    t = atomicAdd(p.sync, 1);
                    if (t >= total) break;
                    i = t / per_row;
                    q = t - i * per_row;
                    pnl = q % p.panels;
                    i0 = p.diag_only ? i : (p.tri_rhs ? (pnl * BN) / NB : 0);
                    if (i >= i0) break;   // identity right-hand side: block rows above the panel's own are exactly zero
                }
                tq[slot] = (t < total) ? t : -1;
                mbar_arrive(&tq_full[slot]);
                if (t >= total) break;
                const int o_local = q / p.panels;
                const int wrow = (int)(o_local * p.w_stride) + pnl * BN;   // row of the panel inside the W slab
                const int lrow = (int)(p.outs[o_local] * p.n_pad);         // first row of this output's L / Dinv
                const int* fl = flags + (size_t)q * T;
                bool vs_loaded = false;
                int issued = 0;
                auto load_vs = [&]() {
                    // right-hand side tile K*_i -> VS (the previous tile's diagonal product must be done with VS)
                    if (seq > 0) mbar_wait(vs_free, (uint32_t)((seq - 1) & 1));
                    mbar_arrive_expect_tx(ks_full, Cfg::VS_BYTES);
                    for (int ch = 0; ch < NCH; ch++)
                        tma_load_3d(VS + ch * (KC / 8) * BN * 8, &tmW, 0, wrow, i * (NB / 8) + ch * (KC / 8), ks_full);
                    vs_loaded = true;
                };
                // V_j of this panel become ready in order of j: if the last one is, all are
                bool all_ready = (i == i0) || (ld_acquire_gpu(fl + i - 1) >= Cfg::NCW);
                if (all_ready) fence_proxy_async();
                // sum_{j<i} L_ij V_j
                for (int j = i0; j < i; j++) {
                    if (!all_ready) {
                        const unsigned long long t0 = globaltimer_ns();
                        while (ld_acquire_gpu(fl + j) < Cfg::NCW) {
                            __nanosleep(200);
                            if (globaltimer_ns() - t0 > 20000000000ull) __trap();   // 20 s: never in a correct run
                        }
                        fence_proxy_async();   // V_j was written through the generic proxy, TMA reads it through the async proxy
                    }
                    for (int ch = 0; ch < NCH; ch++) {
                        if (!vs_loaded && issued == NS - 1) load_vs();
                        mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                        unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ps.stage], Cfg::STAGE_BYTES);
                        const int kout = j * (NB / 8) + ch * (KC / 8);
                        tma_load_3d(st, &tmL, 0, lrow + i * NB, kout, &full[ps.stage]);
                        tma_load_3d(st + Cfg::A_BYTES, &tmW, 0, wrow, kout, &full[ps.stage]);
                        ps.advance();
                        issued++;
                    }
                }
                if (!vs_loaded) load_vs();
                // inv(L_ii)
                for (int ch = 0; ch < NCH; ch++) {
                    mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                    unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ps.stage], Cfg::A_BYTES);
                    tma_load_3d(st, &tmD, 0, lrow + i * NB, ch * (KC / 8), &full[ps.stage]);
                    ps.advance();
                }
                seq++;
            }
        }
        return;
    }

    // =========================== DMMA consumers ===========================
    reg_alloc<224>();
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, t4 = lane & 3;
    const int arow0 = wm * 32, bcol0 = wn * 8 * NT;

    PipeState<NS> ps;
    int seq = 0;
    for (int nq = 0;; nq++) {
        const int slot = nq & 1;
        mbar_wait(&tq_full[slot], (uint32_t)((nq >> 1) & 1));
        const int t = tq[slot];
        __syncwarp();
        if (lane == 0) mbar_arrive(&tq_empty[slot]);
        if (t < 0) break;
        const int i = t / per_row;
        const int q = t - i * per_row;
        const int pnl = q % p.panels;
        const int o_local = q / p.panels;
        const int c0 = pnl * BN;
        const int i0 = p.diag_only ? i : (p.tri_rhs ? c0 / NB : 0);
        const int wrow = (int)(o_local * p.w_stride) + c0;

        double acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.0;

        for (int c = 0; c < (i - i0) * NCH; c++) {
            mbar_wait(&full[ps.stage], ps.phase);
            const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
            const double* Bs = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES + Cfg::A_BYTES);
            mma_stage<2, NT, KC>(acc, As, NB, arow0, Bs, BN, bcol0, g, t4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance();
        }

        // rhs = K*_i - acc, in place in VS (element (k = L row r, n = test point c) at VS[r/8][c][r%8])
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

        // V_i = inv(L_ii) * rhs
        for (int ch = 0; ch < NCH; ch++) {
            mbar_wait(&full[ps.stage], ps.phase);
            const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
            mma_stage<2, NT, KC>(acc, As, NB, arow0, VS + ch * (KC / 8) * BN * 8, BN, bcol0, g, t4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance();
        }
        fence_proxy_async();           // the warp's generic-proxy writes into the staging tile -> the TMA load that overwrites it
        __syncwarp();
        if (lane == 0) mbar_arrive(vs_free);
        seq++;

        const bool last = (i + 1 == T);
        if (!p.tri_rhs && !p.diag_only) {
            // column norms of this tile -> running accumulator of the panel (the chain of tiles of one panel is
            // strictly ordered by the ready flags, so this read-modify-write is race-free and deterministic)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e1 = 0; e1 < 2; e1++) {
                    double v = 0.0;
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        v = fma(acc[mt][nt][e1], acc[mt][nt][e1], v);
                        v = fma(acc[mt][nt][2 + e1], acc[mt][nt][2 + e1], v);
                    }
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    if (g == 0) nred[wm * BN + bcol0 + nt * 8 + 2 * t4 + e1] = v;
                }
            named_bar_sync(1, Cfg::NCW * 32);
            for (int c = threadIdx.x; c < BN; c += Cfg::NCW * 32) {
                const int64_t cg = (int64_t)c0 + c;
                double nrm = (nred[c] + nred[BN + c]) + (nred[2 * BN + c] + nred[3 * BN + c]);
                double* na = p.normacc + (int64_t)o_local * p.w_stride + cg;
                if (i > 0) nrm += __ldcg(na);
                if (!last) {
                    *na = nrm;
                } else if (cg < p.m) {
                    const int o = p.outs[o_local];
                    const double* hyp = p.hyper + (int64_t)o * p.hyper_stride;
                    const double top = hyp[p.d] + (p.include_nugget ? hyp[p.d + 1] : 0.0);
                    p.var[(int64_t)o * p.var_stride + cg] = p.no_clip ? (top - nrm) : fmax(top - nrm, 0.0);
                }
            }
        }
        if (!last || p.tri_rhs || p.keep_v || p.diag_only) {   // the last block row is only needed when V itself is the result
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e1 = 0; e1 < 2; e1++) {
                    const int c = bcol0 + nt * 8 + 2 * t4 + e1;
                    double* wr = p.W + ((int64_t)wrow + c) * p.n_pad + (int64_t)i * NB + arow0 + g;
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        wr[mt * 16] = acc[mt][nt][e1];
                        wr[mt * 16 + 8] = acc[mt][nt][2 + e1];
                    }
                }
        }
        if (!last) {
            // publish V_i (+ the norm accumulator): generic-proxy global writes -> gpu scope -> async-proxy (TMA) readers
            __threadfence();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) red_release_gpu_add(p.sync + SYNC_HDR + (size_t)q * T + i, 1);
        }
    }
}

template <int NT>
static int launch_pred(const TrsmPlan& plan, int count, const CUtensorMap& tmL, const CUtensorMap& tmD,
                       const CUtensorMap& tmW, const PredParams& p, int n_sms, cudaStream_t st) {
    using Cfg = PredCfg<NT>;
    const int64_t tiles = (int64_t)p.T * plan.panels * count;
    const unsigned grid = (unsigned)(tiles < n_sms ? tiles : n_sms);
    if (cudaMemsetAsync(p.sync, 0, predict_sync_bytes(plan, count, p.T), st) != cudaSuccess) return 1;
    predict_trsm_kernel<NT><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmL, tmD, tmW, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int predict_init() {
    if (cudaFuncSetAttribute(predict_trsm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PredCfg<2>::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(predict_trsm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PredCfg<4>::SMEM_BYTES) != cudaSuccess)
        return 1;
    return 0;
}

size_t predict_sync_bytes(const TrsmPlan& plan, int count, int T) {
    return sizeof(int) * ((size_t)SYNC_HDR + (size_t)count * plan.panels * T);
}

// Panel width: the widest of {64, 32} whose dependency chain (T block rows x 2 tile products) stays well
// below the time the whole solve needs on n_sms SMs; narrower panels shorten the chain when there are few
// right-hand sides.
static int pick_width(double flops, int T, int n_sms, int64_t cols) {
    const double per_sm = 37.0e12 / 148.0;   // FP64 tensor rate of one SM (flop/s)
    const double work = flops / (per_sm * n_sms);
    int best = 32;
    const int widths[2] = {64, 32};
    for (int bn : widths) {
        const double crit = (double)T * 2.0 * (2.0 * NB * NB * bn) / per_sm;
        if (crit <= 0.5 * work) {
            best = bn;
            break;
        }
    }
    while (best > 32 && best / 2 >= cols) best /= 2;
    return best;
}

TrsmPlan predict_plan(int64_t m, int n_outputs, int n_pad, int n_sms) {
    const int bn = pick_width((double)n_outputs * n_pad * (double)n_pad * m, n_pad / NB, n_sms, m);
    return TrsmPlan{bn, (int)((m + bn - 1) / bn)};
}

// identity right-hand side (n_pad columns): widths divide 128 so a panel never straddles a block row
TrsmPlan predict_plan_square(int64_t n_pad, int n_sms) {
    const int bn = pick_width((double)n_pad * n_pad * n_pad / 3.0, (int)(n_pad / NB), n_sms, n_pad);
    return TrsmPlan{bn, (int)(n_pad / bn)};
}

int predict_trsm(const TrsmPlan& plan, const int* outs, int count, const CUtensorMap& tmL, const CUtensorMap& tmD,
                 const CUtensorMap& tmW, double* W, int64_t w_stride, const double* hyper, int d, int include_nugget,
                 int64_t n_pad, int64_t m, double* var, int64_t var_stride, int tri_rhs, int* sync, double* normacc,
                 int n_sms, cudaStream_t st, int keep_v, int no_clip, int diag_only) {
    PredParams p{};
    p.diag_only = diag_only;
    p.keep_v = keep_v;
    p.no_clip = no_clip;
    p.W = W; p.w_stride = w_stride; p.n_pad = n_pad; p.m = m; p.T = (int)(n_pad / NB);
    p.panels = plan.panels; p.count = count;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    p.hyper = hyper; p.hyper_stride = d + 2; p.d = d; p.include_nugget = include_nugget; p.var = var;
    p.var_stride = var_stride;
    p.tri_rhs = tri_rhs;
    p.sync = sync;
    p.normacc = normacc;
    switch (plan.nw / 16) {
        case 2: return launch_pred<2>(plan, count, tmL, tmD, tmW, p, n_sms, st);
        case 4: return launch_pred<4>(plan, count, tmL, tmD, tmW, p, n_sms, st);
    }
    return 2;
}

}  // namespace mogp
