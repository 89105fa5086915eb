// Predict TRSM on the tcgen05 tensor cores: FP64-equivalent blocked forward substitution from int8 slices
// (Ozaki-style error-free splitting), for emulators with many right-hand sides (C3 / C5 shapes).
//
// What it computes is what predict.cu computes -- V = L^-1 K* and var_c = sigma2 [+ nugget] - ||V_c||^2
// (GaussianProcess.predict, GaussianProcess.py:896-920) -- reorganised so that the O(n^2 m) part is integer GEMM:
//
//     V_i = inv(L_ii) K*_i - sum_{j<i} (inv(L_ii) L_ij) V_j  =  K~*_i - sum_{j<i} L~_ij V_j
//
//   * L~ = blockdiag(L_ii)^-1 L is formed once per fit in FP64 (i8_lprep_kernel) and stored as S signed 7-bit planes per
//     element with one power-of-two scale per row (a row of L~ is a fixed-point number with 7 S fractional bits);
//   * K~* = blockdiag(L_ii)^-1 K* is formed in place in FP64 (i8_ktilde_kernel, DMMA);
//   * every solved block row V_i is kept only as S int8 planes with one scale per output (||V_c||^2 <= sigma2, so
//     |V| <= sqrt(sigma2 + nugget)): S bytes per element instead of 8, never re-read in FP64;
//   * the products L~_ij V_j are tcgen05.mma.kind::i8 (M = 128, N = 64, K = 32) into s32 accumulators in TMEM: the
//     S (S + 1) / 2 plane pairs (t, u) with t + u <= S + 1 of a K step go to S accumulators, one per weight 2^-7(t+u);
//     pairs of smaller weight are dropped.  Integer accumulation is exact: |digit| <= 64, so a column of n = 16384 terms
//     times 7 pairs stays below 2^30;
//   * recombination (TMEM -> FP64, S weights, row and output scales), the subtraction from K~*_i, the column norms and the
//     slicing of V_i are the epilogue of the consumer warps; the MMA warp already works on the next tile meanwhile.
//
// S = 7 is the library default (>= 100 x inside the variance tolerance on every case measured), S = 6 is opt-in
// (MOGP_TRSM_I8=6; at the edge of the tolerance on small ill-conditioned problems) -- DESIGN.md section 3 has the table.
//
// One launch per block row i (all panels of all outputs of the call are independent inside a launch): a persistent grid,
// per CTA a loader warp (cp.async.bulk of contiguous plane blocks into a 5- or 6-stage mbarrier ring that takes all the
// shared memory), an MMA warp (one elected thread issues, tcgen05.commit frees ring slots / publishes the accumulators)
// and 8 consumer warps.  The kernel is bound by the operand feed out of L2 (5.4 - 5.9 TB/s, profiles/r01_c3_i8_row*.txt).
// tools/ozaki_study.py (profiles/r01_ozaki_study.txt) is the error study, oracle/i8_emulation.py the exact CPU emulation of
// this arithmetic, tools/probe_i8*.cu measured the instruction (exact s32 results; the M128 N64 K32 shape issues at 3.0 POP/s).
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace mogp {

constexpr int I8_BITS = 7;                    // bits per plane (signed digit in [-64, 64])
constexpr int I8_BN = 64;                     // test points per tile (MMA N)
constexpr int I8_APLANE = NB * 32;            // bytes of one plane of a K = 32 step of L~ (128 rows)
constexpr int I8_BPLANE = I8_BN * 32;         // ... of V (64 columns)
constexpr int I8_THREADS = 320;               // 8 consumer warps + MMA warp + loader warp

// S planes per operand: S = 6 keeps the pairs t + u <= 7 (21 MMAs per K step, products resolved to 2^-49 of the row and
// output scales), S = 7 the pairs t + u <= 8 (28 MMAs, 2^-56): the accurate default, see the error table in DESIGN.md
template <int S>
struct I8Cfg {
    static constexpr int NS = (S == 6) ? 6 : 5;             // ring stages: the feed is latency-bound, all shared memory goes to the ring
    static constexpr int ASTAGE = S * I8_APLANE;
    static constexpr int BSTAGE = S * I8_BPLANE;
    static constexpr int STAGE = ASTAGE + BSTAGE;
    static constexpr int LBLOCK = 4 * ASTAGE;               // one 128 x 128 block of L~: 4 K steps
    static constexpr int VBLOCK = 4 * BSTAGE;               // one 128-row block of V for one panel
    static constexpr int SMEM = NS * STAGE + 4 * I8_BN * 8 + 256 + 128;
    static_assert(S * I8_BN <= 512, "one s32 accumulator group per weight must fit TMEM");
};

// byte offset of element (row r, k in [0, 32)) inside a plane of a K = 32 step: K-major, no swizzle, 8 x 16-byte core
// matrices; leading (K half) byte offset 128, stride (8-row group) byte offset 256 (checked by tools/probe_i8.cu)
__device__ __forceinline__ int i8_plane_off(int r, int kk) {
    return (r >> 3) * 256 + ((kk >> 4) & 1) * 128 + (r & 7) * 16 + (kk & 15);
}

__device__ __forceinline__ uint64_t i8_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
// D = s32, A = B = signed 8 bit, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t i8_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(NB >> 4) << 24);
}

__device__ __forceinline__ void i8_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void i8_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void i8_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void i8_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void i8_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// mbarrier wait that traps instead of hanging the GPU if the pipeline protocol is ever violated
__device__ __forceinline__ void i8_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = globaltimer_ns();
    while (!mbar_try_wait(bar, parity))
        if (globaltimer_ns() - t0 > 10000000000ull) __trap();   // 10 s: never in a correct run
}
__device__ __forceinline__ void i8_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
        "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// S signed 7-bit digits of x in (-0.5, 0.5):  x = sum_t d_t 2^-7t + O(2^-(7S+1)).  Every step is exact in FP64.
template <int S>
__device__ __forceinline__ void i8_digits(double x, int8_t (&dig)[S]) {
    x = fmin(fmax(x, -0.99), 0.99);        // in-range data has |x| < 0.5; out-of-range input degrades instead of wrapping int8
    double y = x;
#pragma unroll
    for (int t = 0; t < S; t++) {
        y *= 128.0;
        const double dd = rint(y);
        y -= dd;
        dig[t] = (int8_t)(int)dd;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// L~ = blockdiag(L_ii)^-1 L, strictly lower blocks, FP64; PASS 0: row maxima, PASS 1: digits
// ------------------------------------------------------------------------------------------------------------------
struct I8PrepParams {
    const double* A;       // L slab [E][n_pad][n_pad]
    const double* Dinv;    // [E][n_pad][128]
    int64_t n_pad;
    int outs[MAXG];
    unsigned long long* rowmax;   // [count][n_pad] bit patterns of non-negative doubles
    int8_t* Lq;            // [E][lq_stride]
    int64_t lq_stride;
    int* eL;               // [E][n_pad]
    double* scratch;       // optional [count][blocks][256 threads][64]: pass 0 parks L~ here, pass 1 only slices it
};

template <int PASS, int S>
__global__ void __launch_bounds__(256) i8_lprep_kernel(const I8PrepParams p) {
    using Cfg = I8Cfg<S>;
    __shared__ double As[NB][17];
    __shared__ double Bs[16][NB];
    const int b = blockIdx.x;
    int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)b)) * 0.5f);
    while (i * (i - 1) / 2 > b) i--;
    while ((i + 1) * i / 2 <= b) i++;
    const int j = b - i * (i - 1) / 2;
    const int o = p.outs[blockIdx.y];
    const double* L = p.A + (size_t)o * p.n_pad * p.n_pad;
    const double* D = p.Dinv + (size_t)o * p.n_pad * NB;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
        for (int v = 0; v < 8; v++) acc[u][v] = 0.0;
    double* park = p.scratch ? p.scratch + (((size_t)blockIdx.y * gridDim.x + b) * 256 + tid) * 64 : nullptr;
    if (PASS == 1 && park) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int v = 0; v < 8; v += 2) {
                const double2 t2 = *reinterpret_cast<const double2*>(park + u * 8 + v);
                acc[u][v] = t2.x;
                acc[u][v + 1] = t2.y;
            }
    }
    for (int k0 = 0; k0 < ((PASS == 1 && park) ? 0 : NB); k0 += 16) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int idx = tid + q * 256;
            // inv(L_ii) is lower triangular: entries above the diagonal are taken as exact zeros whatever the slab holds
            As[idx >> 4][idx & 15] = (k0 + (idx & 15) <= (idx >> 4)) ? D[(size_t)(i * NB + (idx >> 4)) * NB + k0 + (idx & 15)] : 0.0;
            Bs[idx >> 7][idx & 127] = L[(size_t)(i * NB + k0 + (idx >> 7)) * p.n_pad + (size_t)j * NB + (idx & 127)];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            double a[8], bb[8];
#pragma unroll
            for (int u = 0; u < 8; u++) a[u] = As[ty * 8 + u][kk];
#pragma unroll
            for (int v = 0; v < 8; v++) bb[v] = Bs[kk][v * 16 + tx];      // columns strided by 16: conflict-free
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int v = 0; v < 8; v++) acc[u][v] = fma(a[u], bb[v], acc[u][v]);
        }
        __syncthreads();
    }
    unsigned long long* rm = p.rowmax + (size_t)blockIdx.y * p.n_pad + (size_t)i * NB;
    if (PASS == 0) {
        if (park) {
#pragma unroll
            for (int u = 0; u < 8; u++)
#pragma unroll
                for (int v = 0; v < 8; v += 2) *reinterpret_cast<double2*>(park + u * 8 + v) = make_double2(acc[u][v], acc[u][v + 1]);
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            double mx = 0.0;
#pragma unroll
            for (int v = 0; v < 8; v++) mx = fmax(mx, fabs(acc[u][v]));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
            if (tx == 0 && mx > 0.0) atomicMax(rm + ty * 8 + u, (unsigned long long)__double_as_longlong(mx));
        }
        return;
    }
    int8_t* blk = p.Lq + (size_t)o * p.lq_stride + (size_t)b * Cfg::LBLOCK;
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int r = ty * 8 + u;
        const double mx = __longlong_as_double((long long)rm[r]);
        int e = 0;
        if (mx > 0.0) frexp(mx, &e);
        e += 1;                                   // scaled row in (-0.5, 0.5)
        if (j == 0 && tx == 0) p.eL[(size_t)o * p.n_pad + (size_t)i * NB + r] = e;
        const double sc = ldexp(1.0, -e);
#pragma unroll
        for (int v = 0; v < 8; v++) {
            const int c = v * 16 + tx;            // column of the block = K index of the MMA
            int8_t dig[S];
            i8_digits<S>(acc[u][v] * sc, dig);
            int8_t* dst = blk + (size_t)(c >> 5) * Cfg::ASTAGE + i8_plane_off(r, c & 31);
#pragma unroll
            for (int t = 0; t < S; t++) dst[(size_t)t * I8_APLANE] = dig[t];
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// K~* = blockdiag(L_ii)^-1 K*, in place in the test-major workspace (FP64 DMMA)
//
// Unit of work = (output, block row i): inv(L_ii) is loaded ONCE into shared memory as the A operand (128 KB, K-blocked) and
// every 32-test-point tile of K*_i streams through a two-slot TMA ring as the B operand -- 32 KB in, one
// mma_stage<1,4,128> per consumer warp (16 rows each), 32 KB out.  (The dataflow kernel of predict.cu, run with an empty
// history, reloads inv(L_ii) for every tile and serialises the right-hand-side load with the product: 16 ms at C3.)
// ------------------------------------------------------------------------------------------------------------------
constexpr int KT_BN = 32;
constexpr int KT_A_BYTES = NB * NB * 8, KT_B_BYTES = KT_BN * NB * 8;
constexpr int KT_SMEM = KT_A_BYTES + 2 * KT_B_BYTES + 128 + 128;

struct KtParams {
    double* W;
    int64_t w_stride, n_pad;
    int T, tiles, count;        // tiles of KT_BN test points per output
    int outs[MAXG];
};

__global__ void __launch_bounds__(288, 1)
i8_ktilde_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmW, const KtParams p) {
    extern __shared__ __align__(128) unsigned char kt_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(kt_smem_raw) + 127) & ~uintptr_t(127));
    double* As = reinterpret_cast<double*>(base);                                  // [16][128][8]
    unsigned char* Bring = base + KT_A_BYTES;                                      // 2 x [16][32][8]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(Bring + 2 * KT_B_BYTES);
    uint64_t* a_empty = a_full + 1;
    uint64_t* b_full = a_empty + 1;   // [2]
    uint64_t* b_empty = b_full + 2;   // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = p.count * p.T;
    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 8);
        for (int s = 0; s < 2; s++) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 8);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == 8) {
        if (lane == 0) {
            prefetch_tmap(&tmD);
            prefetch_tmap(&tmW);
            int it = 0, k = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x, k++) {
                const int o = u / p.T, i = u - o * p.T;
                if (k > 0) i8_wait(a_empty, (uint32_t)((k - 1) & 1));       // consumers are done with the previous inv(L_ii)
                mbar_arrive_expect_tx(a_full, KT_A_BYTES);
                for (int ch = 0; ch < NB / KC; ch++)
                    tma_load_3d(reinterpret_cast<unsigned char*>(As) + ch * (KC / 8) * NB * 64, &tmD, 0,
                                (int)(p.outs[o] * p.n_pad) + i * NB, ch * (KC / 8), a_full);
                for (int tl = 0; tl < p.tiles; tl++, it++) {
                    const int slot = it & 1;
                    if (it >= 2) i8_wait(&b_empty[slot], (uint32_t)(((it >> 1) - 1) & 1));
                    mbar_arrive_expect_tx(&b_full[slot], KT_B_BYTES);
                    for (int ch = 0; ch < NB / KC; ch++)
                        tma_load_3d(Bring + slot * KT_B_BYTES + ch * (KC / 8) * KT_BN * 64, &tmW, 0,
                                    (int)(o * p.w_stride) + tl * KT_BN, i * (NB / 8) + ch * (KC / 8), &b_full[slot]);
                }
            }
        }
        return;
    }
    const int g = lane >> 2, t4 = lane & 3;
    const int arow0 = warp * 16;
    int it = 0, k = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, k++) {
        const int o = u / p.T, i = u - o * p.T;
        i8_wait(a_full, (uint32_t)(k & 1));
        for (int tl = 0; tl < p.tiles; tl++, it++) {
            const int slot = it & 1;
            i8_wait(&b_full[slot], (uint32_t)((it >> 1) & 1));
            double acc[1][4][4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[0][nt][e] = 0.0;
            mma_stage<1, 4, NB>(acc, As, NB, arow0, reinterpret_cast<const double*>(Bring + slot * KT_B_BYTES), KT_BN, 0, g, t4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&b_empty[slot]);
            // W[test point c][i*128 + row] <- acc (row = arow0 + g (+8), c = 8 nt + 2 t4 (+1))
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int e1 = 0; e1 < 2; e1++) {
                    const int c = nt * 8 + 2 * t4 + e1;
                    double* wr = p.W + ((int64_t)o * p.w_stride + (int64_t)tl * KT_BN + c) * p.n_pad + (int64_t)i * NB + arow0 + g;
                    wr[0] = acc[0][nt][e1];
                    wr[8] = acc[0][nt][2 + e1];
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(a_empty);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// one block row of the forward substitution
// ------------------------------------------------------------------------------------------------------------------
struct I8RowParams {
    const int8_t* Lq;
    int64_t lq_stride;
    const int* eL;          // [E][n_pad]
    int8_t* Vq;             // [count][panels][T][I8_VBLOCK]
    const double* W;        // K~* (test-major): [count][w_stride][n_pad]
    int64_t w_stride, n_pad, m;
    int T, panels, count, i;
    int outs[MAXG];
    int eV[MAXG];           // scale exponent of V per local output
    const double* hyper;
    int hyper_stride, d, include_nugget, no_clip;
    int wide;               // wide-N MMA form (default); MOGP_I8_WIDE=0 keeps one MMA per plane pair
    double* var;
    int64_t var_stride;
    double* normacc;        // [count][w_stride]
};

template <int S>
__global__ void __launch_bounds__(I8_THREADS, 1) i8_row_kernel(const I8RowParams p) {
    using Cfg = I8Cfg<S>;
    constexpr int I8_NS = Cfg::NS, I8_STAGE = Cfg::STAGE, I8_ASTAGE = Cfg::ASTAGE, I8_BSTAGE = Cfg::BSTAGE;
    constexpr int I8_LBLOCK = Cfg::LBLOCK, I8_VBLOCK = Cfg::VBLOCK, I8_S = S;
    extern __shared__ __align__(128) unsigned char i8_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(i8_smem_raw) + 127) & ~uintptr_t(127));
    double* nred = reinterpret_cast<double*>(base + I8_NS * I8_STAGE);              // [4][64]
    uint64_t* full = reinterpret_cast<uint64_t*>(nred + 4 * I8_BN);                 // [NS]
    uint64_t* empty = full + I8_NS;                                                 // [NS]
    uint64_t* acc_full = empty + I8_NS;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i = p.i;
    const int nst = 4 * i;                          // K = 32 steps per tile
    const int ntiles = p.count * p.panels;
    const bool last = (i + 1 == p.T);

    if (warp == 8 && nst > 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < I8_NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 8);
        fence_mbar_init();
    }
    i8_fence_before();
    __syncthreads();
    i8_fence_after();
    const uint32_t tmem = (nst > 0) ? *tmem_slot : 0u;

    if (warp == 9) {
        // ================================ loader ================================
        if (lane == 0 && nst > 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int o = tile / p.panels;
                const int8_t* a_src = p.Lq + (size_t)p.outs[o] * p.lq_stride + (size_t)(i * (i - 1) / 2) * I8_LBLOCK;
                const int8_t* b_src = p.Vq + (size_t)tile * p.T * I8_VBLOCK;
                for (int st = 0; st < nst; st++, it++) {
                    const int slot = it % I8_NS;
                    if (it >= I8_NS) i8_wait(&empty[slot], (uint32_t)(((it / I8_NS) - 1) & 1));
                    unsigned char* dst = base + slot * I8_STAGE;
                    mbar_arrive_expect_tx(&full[slot], I8_STAGE);
                    i8_bulk_load(dst, a_src + (size_t)st * I8_ASTAGE, I8_ASTAGE, &full[slot]);
                    i8_bulk_load(dst + I8_ASTAGE, b_src + (size_t)st * I8_BSTAGE, I8_BSTAGE, &full[slot]);
                }
            }
        }
    } else if (warp == 8) {
        // ================================ MMA issuer ================================
        if (lane == 0 && nst > 0) {
            int it = 0, k = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
                if (k > 0) {          // the consumers must have drained the accumulators of the previous tile
                    i8_wait(acc_empty, (uint32_t)((k - 1) & 1));
                    i8_fence_after();
                }
                for (int st = 0; st < nst; st++, it++) {
                    const int slot = it % I8_NS;
                    i8_wait(&full[slot], (uint32_t)((it / I8_NS) & 1));
                    i8_fence_after();
                    const uint32_t a0 = smem_u32(base + slot * I8_STAGE), b0 = a0 + I8_ASTAGE;
                    if (p.wide) {
                        // plane t of L~ against planes 1 .. S+1-t of V in one instruction chain: the planes of V are contiguous in N
                        // (a plane is 8 row groups of 256 B) and the accumulators are contiguous in TMEM in weight order, so
                        // D[:, 64 (t-1) ...] += A_t [B_1 | B_2 | ... | B_{S+1-t}] lands every pair (t, u) in the accumulator of weight
                        // t + u.  10 MMAs of N <= 256 per K step instead of 28 of N = 64: the A plane is read from shared memory once
                        // per t instead of once per pair (the N = 64 shape is bound by those reads: 6 KB per 32-cycle MMA).
#pragma unroll
                        for (int t = 1; t <= I8_S; t++) {
                            const int ncols = I8_BN * (I8_S + 1 - t);
                            const uint32_t accum = (st == 0 && t == 1) ? 0u : 1u;
                            const uint32_t d0 = tmem + (uint32_t)(t - 1) * I8_BN;
                            const uint64_t ad = i8_desc(a0 + (t - 1) * I8_APLANE);
                            const int n1 = ncols > 256 ? 256 : ncols;
                            i8_mma(d0, ad, i8_desc(b0), accum, i8_idesc(n1));
                            if (ncols > 256) i8_mma(d0 + 256, ad, i8_desc(b0 + 4 * I8_BPLANE), accum, i8_idesc(ncols - 256));
                        }
                    } else {
#pragma unroll
                    for (int w = 2; w <= I8_S + 1; w++) {
                        uint32_t accum = (st == 0) ? 0u : 1u;
#pragma unroll
                        for (int t = 1; t < w; t++) {
                            const int u = w - t;
                            if (t > I8_S || u > I8_S) continue;
                            i8_mma(tmem + (uint32_t)(w - 2) * I8_BN, i8_desc(a0 + (t - 1) * I8_APLANE),
                                   i8_desc(b0 + (u - 1) * I8_BPLANE), accum, i8_idesc(I8_BN));
                            accum = 1u;
                        }
                    }
                    }
                    i8_commit(&empty[slot]);      // the slot is free once these MMAs have read it
                }
                i8_commit(acc_full);              // every MMA of the tile has completed
            }
        }
    } else {
        // ================================ consumers ================================
        const int q4 = warp & 3, h = warp >> 2;
        const int r = q4 * 32 + lane;             // row of the block row = TMEM lane
        int k = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
            const int o = tile / p.panels, pnl = tile - o * p.panels;
            double acc[32];
#pragma unroll
            for (int c = 0; c < 32; c++) acc[c] = 0.0;
            if (nst > 0) {
                i8_wait(acc_full, (uint32_t)(k & 1));
                i8_fence_after();
#pragma unroll
                for (int g = 0; g < I8_S; g++) {
                    uint32_t v[32];
                    i8_tmem_ld32(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(g * I8_BN + h * 32), v);
                    const double wgt = __longlong_as_double((long long)(1023 - I8_BITS * (g + 2)) << 52);   // 2^-7(g+2)
#pragma unroll
                    for (int c = 0; c < 32; c++) acc[c] = fma((double)(int32_t)v[c], wgt, acc[c]);
                }
                i8_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
            }
            // ---- V_i = K~*_i - 2^(eL + eV) acc ----
            const int ev = p.eV[o];
            const double scale = (nst > 0) ? ldexp(1.0, p.eL[(size_t)p.outs[o] * p.n_pad + (size_t)i * NB + r] + ev) : 0.0;
            const double vinv = ldexp(1.0, -ev);
            const double* wsrc = p.W + ((size_t)o * p.w_stride + (size_t)pnl * I8_BN + h * 32) * p.n_pad + (size_t)i * NB + r;
            // digits of V_i go straight to the planes in global memory (K step of the later products = r / 32); a warp's
            // store of one (column, plane) is two 16-byte runs
            int8_t* dimg = p.Vq + ((size_t)tile * p.T + i) * I8_VBLOCK + q4 * I8_BSTAGE;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const double v = __ldcs(wsrc + (size_t)c * p.n_pad) - acc[c] * scale;
                if (!last) {
                    int8_t dig[I8_S];
                    i8_digits<I8_S>(v * vinv, dig);
                    int8_t* dst = dimg + i8_plane_off(h * 32 + c, lane);
#pragma unroll
                    for (int t = 0; t < I8_S; t++) dst[t * I8_BPLANE] = dig[t];
                }
                double s = v * v;
                s += __shfl_xor_sync(0xffffffffu, s, 16);
                s += __shfl_xor_sync(0xffffffffu, s, 8);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                if (lane == 0) nred[q4 * I8_BN + h * 32 + c] = s;
            }
            named_bar_sync(1, 256);
            if (tid < I8_BN) {
                const int64_t cg = (int64_t)pnl * I8_BN + tid;
                double nrm = (nred[tid] + nred[I8_BN + tid]) + (nred[2 * I8_BN + tid] + nred[3 * I8_BN + tid]);
                double* na = p.normacc + (int64_t)o * p.w_stride + cg;
                if (i > 0) nrm += *na;
                if (!last) {
                    *na = nrm;
                } else if (cg < p.m) {
                    const int og = p.outs[o];
                    const double* hyp = p.hyper + (int64_t)og * p.hyper_stride;
                    const double top = hyp[p.d] + (p.include_nugget ? hyp[p.d + 1] : 0.0);
                    p.var[(int64_t)og * p.var_stride + cg] = p.no_clip ? (top - nrm) : fmax(top - nrm, 0.0);
                }
            }
            named_bar_sync(1, 256);      // nred is reused by the next tile
        }
    }
    i8_fence_before();
    __syncthreads();
    if (warp == 8 && nst > 0) {
        i8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
int i8_ktilde(const int* outs, int count, const CUtensorMap& tmD, const CUtensorMap& tmW32, double* W, int64_t w_stride,
              int64_t n_pad, int64_t m_rows, int n_sms, cudaStream_t st) {
    KtParams p{};
    p.W = W; p.w_stride = w_stride; p.n_pad = n_pad; p.T = (int)(n_pad / NB); p.count = count;
    p.tiles = (int)((m_rows + KT_BN - 1) / KT_BN);
    for (int k = 0; k < count; k++) p.outs[k] = outs[k];
    const int units = count * p.T;
    i8_ktilde_kernel<<<(unsigned)(units < n_sms ? units : n_sms), 288, KT_SMEM, st>>>(tmD, tmW32, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int i8_init() {
    if (cudaFuncSetAttribute(i8_ktilde_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KT_SMEM) != cudaSuccess) return 1;
    if (cudaFuncSetAttribute(i8_row_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<6>::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(i8_row_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8Cfg<7>::SMEM) != cudaSuccess)
        return 1;
    return 0;
}

static size_t lblock(int S) { return S == 6 ? I8Cfg<6>::LBLOCK : I8Cfg<7>::LBLOCK; }
static size_t vblock(int S) { return S == 6 ? I8Cfg<6>::VBLOCK : I8Cfg<7>::VBLOCK; }
size_t i8_lq_bytes(int T, int S) { return (size_t)T * (T - 1) / 2 * lblock(S); }
size_t i8_vq_bytes(int count, int panels, int T, int S) { return (size_t)count * panels * T * vblock(S); }
int i8_panel_width() { return I8_BN; }

size_t i8_scratch_bytes(int count, int T) { return (size_t)(count < MAXG ? count : MAXG) * (T * (T - 1) / 2) * NB * NB * 8; }

int i8_prepare_L(int S, const double* A_slab, const double* Dinv_slab, int64_t n_pad, const int* outs, int count, int8_t* Lq,
                 int64_t lq_stride, int* eL, unsigned long long* rowmax, double* scratch, cudaStream_t st) {
    const int T = (int)(n_pad / NB);
    if (T < 2 || count < 1) return 0;
    for (int g0 = 0; g0 < count; g0 += MAXG) {
        const int cnt = count - g0 < MAXG ? count - g0 : MAXG;
        I8PrepParams p{};
        p.A = A_slab; p.Dinv = Dinv_slab; p.n_pad = n_pad; p.rowmax = rowmax; p.Lq = Lq; p.lq_stride = lq_stride; p.eL = eL;
        p.scratch = scratch;
        for (int k = 0; k < cnt; k++) p.outs[k] = outs[g0 + k];
        if (cudaMemsetAsync(rowmax, 0, sizeof(unsigned long long) * (size_t)cnt * n_pad, st) != cudaSuccess) return 1;
        const dim3 grid((unsigned)(T * (T - 1) / 2), (unsigned)cnt);
        i8_lprep_kernel<0, 6><<<grid, 256, 0, st>>>(p);
        if (S == 6) i8_lprep_kernel<1, 6><<<grid, 256, 0, st>>>(p);
        else i8_lprep_kernel<1, 7><<<grid, 256, 0, st>>>(p);
        if (cudaGetLastError() != cudaSuccess) return 1;
    }
    return 0;
}

// W holds K~* = blockdiag(L_ii)^-1 K* on entry (test-major); var receives the variances after the last block row
int i8_trsm(int S, const int* outs, int count, int panels, const int8_t* Lq, int64_t lq_stride, const int* eL, int8_t* Vq,
            const double* W, int64_t w_stride, const double* hyper, const double* h_hyper, int d, int include_nugget,
            int no_clip, int64_t n_pad, int64_t m, double* var, int64_t var_stride, double* normacc, int n_sms,
            cudaStream_t st) {
    I8RowParams p{};
    p.Lq = Lq; p.lq_stride = lq_stride; p.eL = eL; p.Vq = Vq; p.W = W; p.w_stride = w_stride; p.n_pad = n_pad; p.m = m;
    p.T = (int)(n_pad / NB); p.panels = panels; p.count = count;
    p.hyper = hyper; p.hyper_stride = d + 2; p.d = d; p.include_nugget = include_nugget; p.no_clip = no_clip;
    p.var = var; p.var_stride = var_stride; p.normacc = normacc;
    for (int k = 0; k < count; k++) {
        p.outs[k] = outs[k];
        // |V| <= sqrt(k(x*, x*)) = sqrt(sigma2); one more binade of head-room for rounding
        const double bound = sqrt(h_hyper[(size_t)outs[k] * (d + 2) + d] + h_hyper[(size_t)outs[k] * (d + 2) + d + 1]);
        int e = 0;
        frexp(bound > 0.0 ? bound : 1.0, &e);
        p.eV[k] = e + 1;
    }
    {
        const char* e = getenv("MOGP_I8_WIDE");
        p.wide = (e && e[0] == '0') ? 0 : 1;
    }
    const int ntiles = count * panels;
    const unsigned grid = (unsigned)(ntiles < n_sms ? ntiles : n_sms);
    for (int i = 0; i < p.T; i++) {
        p.i = i;
        if (S == 6) i8_row_kernel<6><<<grid, I8_THREADS, I8Cfg<6>::SMEM, st>>>(p);
        else i8_row_kernel<7><<<grid, I8_THREADS, I8Cfg<7>::SMEM, st>>>(p);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mogp
