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
    static constexpr int NS = (NT == 8) ? 3 : 4;
    static constexpr int A_BYTES = BM * KC * 8;
    static constexpr int B_BYTES = BN * KC * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int VS_BYTES = NB * BN * 8;
    static constexpr int BAR_BYTES = (2 * NS + 3) * 8;
    static constexpr int SMEM_BYTES = NS * STAGE_BYTES + VS_BYTES + BAR_BYTES + 4 * BN * 8 + 128;
};

struct PredParams {
    double* W;              // workspace slab [count][w_stride][n_pad]
    int64_t w_stride;       // rows per output in the slab
    int64_t n_pad;
    int64_t m;              // real number of test points
    int T;                  // n_pad / 128
    int outs[MAXG];         // global output index handled by blockIdx.y
    const double* hyper;    // [E][hyper_stride]
    int hyper_stride;       // d + 2
    int d;
    int include_nugget;
    int tri_rhs;            // right-hand side is the identity: panel c0 starts its walk at block row c0/128
    double* var;            // result rows: var of output o at var + o*var_stride
    int64_t var_stride;
};

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
    uint64_t* step_done = vs_free + 1;
    double* nred = reinterpret_cast<double*>(step_done + 1);  // [4][BN]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o_local = blockIdx.y;
    const int o = p.outs[o_local];
    const int c0 = blockIdx.x * BN;                             // first test point of the panel
    const int wrow = (int)(o_local * p.w_stride) + c0;          // row of the panel inside the W slab
    const int lrow = (int)(o * p.n_pad);                        // first row of this output's L / Dinv
    const int T = p.T;
    const int i0 = p.tri_rhs ? (c0 / NB) : 0;   // V rows above block row i0 are exactly zero

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::NCW);
        }
        mbar_init(ks_full, 1);
        mbar_init(vs_free, Cfg::NCW);
        mbar_init(step_done, Cfg::NCW);
        fence_mbar_init();
    }
    __syncthreads();

    constexpr int NCH = NB / KC;  // chunks per 128-wide K block
    if (warp >= Cfg::NCW) {
        // =========================== TMA producer ===========================
        reg_dealloc<40>();
        if (warp == Cfg::NCW && lane == 0) {
            prefetch_tmap(&tmL);
            prefetch_tmap(&tmD);
            prefetch_tmap(&tmW);
            PipeState<NS> ps;
            for (int i = i0; i < T; i++) {
                const int step = i - i0;
                // right-hand side tile K*_i -> VS (needs the previous step's diagonal product done with VS)
                if (step > 0) mbar_wait(vs_free, (uint32_t)((step - 1) & 1));
                mbar_arrive_expect_tx(ks_full, Cfg::VS_BYTES);
                for (int ch = 0; ch < NCH; ch++)
                    tma_load_3d(VS + ch * (KC / 8) * BN * 8, &tmW, 0, wrow, i * (NB / 8) + ch * (KC / 8), ks_full);
                // sum_{j<i} L_ij V_j
                for (int j = i0; j < i; j++) {
                    if (j == i - 1) mbar_wait(step_done, (uint32_t)((step - 1) & 1));  // V_{i-1} is in HBM/L2
                    for (int ch = 0; ch < NCH; ch++) {
                        mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                        unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[ps.stage], Cfg::STAGE_BYTES);
                        const int kout = j * (NB / 8) + ch * (KC / 8);
                        tma_load_3d(st, &tmL, 0, lrow + i * NB, kout, &full[ps.stage]);
                        tma_load_3d(st + Cfg::A_BYTES, &tmW, 0, wrow, kout, &full[ps.stage]);
                        ps.advance();
                    }
                }
                // inv(L_ii)
                for (int ch = 0; ch < NCH; ch++) {
                    mbar_wait(&empty[ps.stage], ps.phase ^ 1u);
                    unsigned char* st = base + ps.stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[ps.stage], Cfg::A_BYTES);
                    tma_load_3d(st, &tmD, 0, lrow + i * NB, ch * (KC / 8), &full[ps.stage]);
                    ps.advance();
                }
            }
        }
        return;
    }

    // =========================== DMMA consumers ===========================
    reg_alloc<232>();
    const int wm = warp >> 1, wn = warp & 1;
    const int g = lane >> 2, t = lane & 3;
    const int arow0 = wm * 32, bcol0 = wn * 8 * NT;
    double colsum[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) colsum[nt][0] = colsum[nt][1] = 0.0;

    PipeState<NS> ps;
    for (int i = i0; i < T; i++) {
        const int step = i - i0;
        double acc[2][NT][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[mt][nt][e] = 0.0;

        for (int c = 0; c < step * NCH; c++) {
            mbar_wait(&full[ps.stage], ps.phase);
            const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
            const double* Bs = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES + Cfg::A_BYTES);
            mma_stage<2, NT, KC>(acc, As, NB, arow0, Bs, BN, bcol0, g, t);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance();
        }

        // rhs = K*_i - acc, in place in VS (element (k = L row r, n = test point c) at VS[r/8][c][r%8])
        mbar_wait(ks_full, (uint32_t)(step & 1));
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int r = arow0 + mt * 16 + g + ((e >> 1) << 3);
                    const int c = bcol0 + nt * 8 + 2 * t + (e & 1);
                    double* q = VS + ((size_t)((r >> 3) * BN + c) * 8 + (r & 7));
                    *q = *q - acc[mt][nt][e];
                    acc[mt][nt][e] = 0.0;
                }
        named_bar_sync(1, Cfg::NCW * 32);

        // V_i = inv(L_ii) * rhs
        for (int ch = 0; ch < NCH; ch++) {
            mbar_wait(&full[ps.stage], ps.phase);
            const double* As = reinterpret_cast<const double*>(base + ps.stage * Cfg::STAGE_BYTES);
            mma_stage<2, NT, KC>(acc, As, NB, arow0, VS + ch * (KC / 8) * BN * 8, BN, bcol0, g, t);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]);
            ps.advance();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(vs_free);

        // epilogue: norms + in-place store of V_i (test-major) for the later block rows
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const double v = acc[mt][nt][e];
                    colsum[nt][e & 1] = fma(v, v, colsum[nt][e & 1]);
                }
        if (i + 1 < T || p.tri_rhs) {   // the last block row is only needed when V itself is the result
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e1 = 0; e1 < 2; e1++) {
                    const int c = bcol0 + nt * 8 + 2 * t + e1;
                    double* wr = p.W + ((int64_t)wrow + c) * p.n_pad + (int64_t)i * NB + arow0 + g;
#pragma unroll
                    for (int mt = 0; mt < 2; mt++) {
                        wr[mt * 16] = acc[mt][nt][e1];
                        wr[mt * 16 + 8] = acc[mt][nt][2 + e1];
                    }
                }
            // generic-proxy global writes -> L2 -> visible to the TMA (async proxy) loads of the next steps
            if (i + 1 < T) {
                __threadfence();
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(step_done);
            }
        }
    }

    // var_c = max(sigma2 (+ nugget) - ||V_c||^2, 0)
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
#pragma unroll
        for (int e1 = 0; e1 < 2; e1++) {
            double v = colsum[nt][e1];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) nred[wm * BN + bcol0 + nt * 8 + 2 * t + e1] = v;
        }
    named_bar_sync(1, Cfg::NCW * 32);
    const double* hyp = p.hyper + (int64_t)o * p.hyper_stride;
    const double top = hyp[p.d] + (p.include_nugget ? hyp[p.d + 1] : 0.0);
    for (int c = threadIdx.x; c < BN; c += Cfg::NCW * 32) {
        const int64_t cg = (int64_t)c0 + c;
        if (cg < p.m) {
            const double nrm = (nred[c] + nred[BN + c]) + (nred[2 * BN + c] + nred[3 * BN + c]);
            p.var[(int64_t)o * p.var_stride + cg] = fmax(top - nrm, 0.0);
        }
    }
}

template <int NT>
static int launch_pred(const TrsmPlan& plan, int count, const CUtensorMap& tmL, const CUtensorMap& tmD,
                       const CUtensorMap& tmW, const PredParams& p, cudaStream_t st) {
    using Cfg = PredCfg<NT>;
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(predict_trsm_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) !=
            cudaSuccess)
            return 1;
        attr_done = true;
    }
    dim3 grid((unsigned)plan.panels, (unsigned)count);
    predict_trsm_kernel<NT><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmL, tmD, tmW, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int predict_init() { return 0; }

// Pick the panel width that minimises waves * width on n_sms SMs (one CTA per SM).
TrsmPlan predict_plan(int64_t m, int n_outputs, int n_sms) {
    TrsmPlan best{64, (int)((m + 63) / 64)};
    double best_cost = 1e300;
    for (int nt = 4; nt <= 8; nt++) {
        const int bn = 16 * nt;
        const int64_t panels = (m + bn - 1) / bn;
        const int64_t ctas = panels * n_outputs;
        const int64_t waves = (ctas + n_sms - 1) / n_sms;
        const double cost = (double)waves * bn;
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best.nw = bn;
            best.panels = (int)panels;
        }
    }
    return best;
}

// identity right-hand side (n_pad columns): widths that divide 128 so a panel never straddles a block row
TrsmPlan predict_plan_square(int64_t n_pad, int n_sms) {
    TrsmPlan best{128, (int)(n_pad / 128)};
    double best_cost = 1e300;
    const int widths[3] = {32, 64, 128};
    for (int bn : widths) {
        const int64_t panels = n_pad / bn;
        const int64_t waves = (panels + n_sms - 1) / n_sms;
        const double cost = (double)waves * bn;
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best.nw = bn;
            best.panels = (int)panels;
        }
    }
    return best;
}

int predict_trsm(const TrsmPlan& plan, const int* outs, int count, const CUtensorMap& tmL, const CUtensorMap& tmD,
                 const CUtensorMap& tmW, double* W, int64_t w_stride, const double* hyper, int d, int include_nugget,
                 int64_t n_pad, int64_t m, double* var, int64_t var_stride, int tri_rhs, cudaStream_t st) {
    PredParams p{};
    p.W = W; p.w_stride = w_stride; p.n_pad = n_pad; p.m = m; p.T = (int)(n_pad / NB);
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    p.hyper = hyper; p.hyper_stride = d + 2; p.d = d; p.include_nugget = include_nugget; p.var = var;
    p.var_stride = var_stride;
    p.tri_rhs = tri_rhs;
    switch (plan.nw / 16) {
        case 2: return launch_pred<2>(plan, count, tmL, tmD, tmW, p, st);
        case 4: return launch_pred<4>(plan, count, tmL, tmD, tmW, p, st);
        case 5: return launch_pred<5>(plan, count, tmL, tmD, tmW, p, st);
        case 6: return launch_pred<6>(plan, count, tmL, tmD, tmW, p, st);
        case 7: return launch_pred<7>(plan, count, tmL, tmD, tmW, p, st);
        case 8: return launch_pred<8>(plan, count, tmL, tmD, tmW, p, st);
    }
    return 2;
}

}  // namespace mogp
