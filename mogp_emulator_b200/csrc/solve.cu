// Single right-hand-side triangular solves for libmogp_b200:  z = L^-1 y,  alpha = L^-T z,  quad = z^T z.
// Replaces cusolverDnDpotrs for invQt (reference mogp_gpu/src/densegp_gpu.hpp:585-591) and the cublasDdot
// for y^T alpha (:604-611); CPU semantics: ChoInv.solve = cho_solve((L, True), y), linalg/cholesky.py:22-42,
// used at GaussianProcess.py:666-672.
//
// Blocked substitution over 128-row blocks, one thread-block CLUSTER per right-hand side.  The off-diagonal
// part of every step is a streaming GEMV over a block row (forward) / block column (backward) of L whose
// 128-wide column (row) blocks are dealt round-robin to the CTAs of the cluster, so the cluster pulls L through
// C SMs' worth of L2->SM bandwidth instead of one.  Each CTA keeps only its own blocks of the solution vector
// in shared memory (those are the only ones its share of the GEMV touches).  Per step the C partial sums land
// in the owner CTA's shared memory through DSMEM (st.shared::cluster), one barrier.cluster publishes them, and
// the owner applies the precomputed inverse of the diagonal block (Dinv, from the Cholesky panel kernel) -- no
// scalar dependency chain.  Partials are summed in rank order: results are deterministic.  The GEMV of step i+1 over
// the blocks already known runs between the arrive and the wait of step i's barrier (look-ahead): only the product
// with the newest block and the Dinv product are on the chain, and their operands are prefetched into L2 a step ahead.
#include "common.cuh"
#include "kernels.h"

namespace mogp {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the two 128 x 128 blocks the NEXT step has on its chain -- the product with the newest block of the solution and the inverse
// of the next diagonal block -- are pulled into L2 one step ahead (they come from HBM otherwise: the factor of 32 outputs is
// 4 GB); every thread touches two 128-byte lines of a block
__device__ __forceinline__ void prefetch_block_l2(const double* blk, int64_t row_stride, int tid) {
    const double* p = blk + (int64_t)(tid >> 2) * row_stride + (tid & 3) * 32;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 16));
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// store a double into the shared memory of CTA `rank` of this cluster at the address `local` has in this CTA
__device__ __forceinline__ void dsmem_store(double* local, uint32_t rank, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

constexpr int SOLVE_THREADS = 512;

struct SolveParams {
    const double* A;      // matrix slab [E*n_pad][n_pad] holding the factors
    const double* Dinv;   // [E*n_pad][128]
    const double* Y;      // [E][n_pad]
    double* z;            // [E][n_pad]
    double* alpha;        // [E][n_pad]
    double* scal;         // [E][2]: scal[2*o+1] receives y^T K^-1 y
    const int* info;      // [E]
    int64_t n_pad;
    int T;
    int outs[MAXG];       // slab index of the output solved by cluster k
};

__global__ void __launch_bounds__(SOLVE_THREADS, 1)
solve_alpha_kernel(const SolveParams sp) {
    extern __shared__ __align__(16) double sv[];
    const int C = (int)cluster_nctarank(), me = (int)cluster_ctarank();
    const int o = sp.outs[blockIdx.x / C];
    // every CTA of the cluster takes the same branch: no barrier is left dangling
    if (sp.info[o] != 0) return;
    const int64_t ld = sp.n_pad;
    const int T = sp.T;
    const double* __restrict__ A = sp.A + (int64_t)o * ld * ld;
    const double* __restrict__ Dinv = sp.Dinv + (int64_t)o * ld * NB;
    const double* __restrict__ y = sp.Y + (int64_t)o * ld;
    double* __restrict__ z_out = sp.z + (int64_t)o * ld;
    double* __restrict__ alpha_out = sp.alpha + (int64_t)o * ld;
    double* __restrict__ quad = sp.scal + 2 * o + 1;
    const int nown = (T + C - 1) / C;        // blocks of the vector kept by one CTA (block j lives in CTA j % C, slot j / C)
    double* v = sv;                          // [nown][128]
    double* slots = v + (size_t)nown * NB;   // [C][128] partial sums received from the cluster
    double* acc = slots + (size_t)C * NB;    // [128]
    double* red = acc + NB;                  // [4][128]
    double* qs = red + 4 * NB;               // [C] per-CTA sums of z^2 (CTA 0's copy is the one used)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int s = 0; s < nown; s++) {
        const int j = s * C + me;
        for (int r = tid; r < NB; r += SOLVE_THREADS) v[s * NB + r] = (j < T) ? y[(int64_t)j * NB + r] : 0.0;
    }
    __syncthreads();
    cluster_sync_all();   // all CTAs are running before the first remote store

    // ---- forward substitution:  z_i = Dinv_i (y_i - sum_{j<i} L_ij z_j) ----
    // Look-ahead: the part of row i's sum over the blocks j <= i-2 was computed during step i-1, between the arrive and the wait
    // of that step's cluster barrier (while its owner applied Dinv); only the product with the newest block z_{i-1} -- held by the
    // CTA that has just computed it -- is on the chain.  One cluster barrier per step; every value a CTA reads from its own
    // shared memory was written by itself (each CTA owns its blocks of the solution), only the partial sums travel (DSMEM).
    // Before, a step began with the whole GEMV over block row i (i / C blocks of 128 KB from HBM per CTA) and had two barriers:
    // 12.5 us per step at n = 16384.
    {
        auto gemv_block = [&](double (&s)[8], int i, int j, int sl) {      // s += rows (8 per warp) of L_ij times block sl of v
            const double* rowp = A + (int64_t)(i * NB + warp * 8) * ld + (int64_t)j * NB;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const int cl = half * 64 + lane * 2;
                const double2 vv = *reinterpret_cast<const double2*>(v + sl * NB + cl);
                const double* p = rowp + cl;
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {
                    const double2 l = *reinterpret_cast<const double2*>(p + (int64_t)rr * ld);
                    s[rr] = fma(l.x, vv.x, s[rr]);
                    s[rr] = fma(l.y, vv.y, s[rr]);
                }
            }
        };
        double pre[8];
#pragma unroll
        for (int rr = 0; rr < 8; rr++) pre[rr] = 0.0;
        for (int i = 0; i < T; i++) {
            const int owner = i % C;
            if (i + 1 < T && C >= 8) {     // (few, wide clusters: latency-bound; with many narrow ones the prefetches only add traffic)
                if (me == owner) prefetch_block_l2(A + (int64_t)(i + 1) * NB * ld + (int64_t)i * NB, ld, tid);      // L_{i+1,i}
                if (me == (i + 1) % C) prefetch_block_l2(Dinv + (int64_t)(i + 1) * NB * NB, NB, tid);
            }
            if (i > 0 && me == (i - 1) % C) gemv_block(pre, i, i - 1, (i - 1) / C);   // the newest block (CTA-uniform branch)
#pragma unroll
            for (int rr = 0; rr < 8; rr++) {
                const double t = warp_sum(pre[rr]);
                if (lane == 0) dsmem_store(slots + me * NB + warp * 8 + rr, (uint32_t)owner, t);
                pre[rr] = 0.0;
            }
            cluster_arrive();
            if (i + 1 < T)
                for (int j = me, sl = 0; j <= i - 1; j += C, sl++) gemv_block(pre, i + 1, j, sl);
            cluster_wait();
            if (me == owner) {   // CTA-uniform
                const int sl = i / C;
                if (tid < NB) {
                    double t = 0.0;
                    for (int c = 0; c < C; c++) t += slots[c * NB + tid];
                    acc[tid] = v[sl * NB + tid] - t;
                }
                __syncthreads();
                const double* Db = Dinv + (int64_t)i * NB * NB;
                const double2 a0 = *reinterpret_cast<const double2*>(acc + lane * 4);
                const double2 a1 = *reinterpret_cast<const double2*>(acc + lane * 4 + 2);
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {
                    const double* dr = Db + (int64_t)(warp * 8 + rr) * NB + lane * 4;
                    const double2 d0 = *reinterpret_cast<const double2*>(dr);
                    const double2 d1 = *reinterpret_cast<const double2*>(dr + 2);
                    double t = d0.x * a0.x;
                    t = fma(d0.y, a0.y, t);
                    t = fma(d1.x, a1.x, t);
                    t = fma(d1.y, a1.y, t);
                    t = warp_sum(t);
                    if (lane == 0) {
                        v[sl * NB + warp * 8 + rr] = t;
                        z_out[(int64_t)i * NB + warp * 8 + rr] = t;
                    }
                }
                __syncthreads();
            }
        }
    }

    // ---- quad = z^T z: per-CTA sums over own blocks, combined in rank order by CTA 0 ----
    {
        double q = 0.0;
        for (int idx = tid; idx < nown * NB; idx += SOLVE_THREADS) q = fma(v[idx], v[idx], q);   // slots past T hold zeros
        q = warp_sum(q);
        if (lane == 0) red[warp] = q;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < SOLVE_THREADS / 32; w++) t += red[w];
            dsmem_store(qs + me, 0u, t);
        }
        cluster_sync_all();
        if (me == 0 && tid == 0) {
            double t = 0.0;
            for (int c = 0; c < C; c++) t += qs[c];
            *quad = t;
        }
    }

    // ---- backward substitution:  alpha_i = Dinv_i^T (z_i - sum_{j>i} L_ji^T alpha_j) ----
    // The same look-ahead, mirrored: block column i over the blocks j >= i+2 during step i+1, the newest block alpha_{i+1} on
    // the chain.
    const int c = tid & 127, q4 = tid >> 7;
    {
        auto colsum_block = [&](int i, int j) -> double {      // sum over this thread's rows r of L_ji[r][c] alpha_j[r]
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double* vb = v + (j / C) * NB;
            const double* lp = A + (int64_t)i * NB + c + (int64_t)j * NB * ld;
#pragma unroll
            for (int r = q4; r < NB; r += 16) {
                s0 = fma(lp[(int64_t)r * ld], vb[r], s0);
                s1 = fma(lp[(int64_t)(r + 4) * ld], vb[r + 4], s1);
                s2 = fma(lp[(int64_t)(r + 8) * ld], vb[r + 8], s2);
                s3 = fma(lp[(int64_t)(r + 12) * ld], vb[r + 12], s3);
            }
            return (s0 + s1) + (s2 + s3);
        };
        double preb = 0.0;
        for (int i = T - 1; i >= 0; i--) {
            const int owner = i % C;
            if (i >= 1 && C >= 8) {
                if (me == owner) prefetch_block_l2(A + (int64_t)i * NB * ld + (int64_t)(i - 1) * NB, ld, tid);      // L_{i,i-1}
                if (me == (i - 1) % C) prefetch_block_l2(Dinv + (int64_t)(i - 1) * NB * NB, NB, tid);
            }
            if (i + 1 < T && me == (i + 1) % C) preb += colsum_block(i, i + 1);   // the newest block (CTA-uniform branch)
            red[q4 * NB + c] = preb;
            preb = 0.0;
            __syncthreads();
            if (tid < NB)
                dsmem_store(slots + me * NB + tid, (uint32_t)owner,
                            (red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]));
            cluster_arrive();
            if (i >= 1) {
                // own row blocks j >= i + 1, j = me (mod C), of block column i - 1
                int j = i + 1 + ((me - (i + 1)) % C + C) % C;
                for (; j < T; j += C) preb += colsum_block(i - 1, j);
            }
            cluster_wait();
            if (me == owner) {
                const int sl = i / C;
                __syncthreads();       // (every thread of this CTA is past its read of red above)
                if (tid < NB) {
                    double t = 0.0;
                    for (int cc = 0; cc < C; cc++) t += slots[cc * NB + tid];
                    acc[tid] = v[sl * NB + tid] - t;
                }
                __syncthreads();
                {
                    const double* Db = Dinv + (int64_t)i * NB * NB;
                    double s = 0.0;
                    for (int r = q4; r < NB; r += 4)
                        if (r >= c) s = fma(Db[r * NB + c], acc[r], s);
                    red[q4 * NB + c] = s;
                }
                __syncthreads();
                if (tid < NB) {
                    const double a = (red[tid] + red[NB + tid]) + (red[2 * NB + tid] + red[3 * NB + tid]);
                    v[sl * NB + tid] = a;
                    alpha_out[(int64_t)i * NB + tid] = a;
                }
            }
            __syncthreads();   // red is rewritten by the next step; alpha_i is in place for the owner's next product
        }
    }
}

static int g_max_cluster = 0;

int solve_init() {
    if (cudaFuncSetAttribute(solve_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return 1;
    // clusters of 16 CTAs need the non-portable opt-in; fall back to the portable 8
    g_max_cluster = 8;
    if (cudaFuncSetAttribute(solve_alpha_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(16);
        cfg.blockDim = dim3(SOLVE_THREADS);
        cfg.dynamicSmemBytes = 64 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 16;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, solve_alpha_kernel, &cfg) == cudaSuccess && ncl >= 1) g_max_cluster = 16;
    }
    cudaGetLastError();
    return 0;
}

// z = L^-1 y, alpha = L^-T z, quad = z^T z for `count` outputs (slab indices outs[]): one cluster per output, one launch.
int solve_alpha(const double* A_slab, int64_t n_pad, const double* Dinv_slab, const double* Y, double* z, double* alpha,
                double* scal, const int* info, const int* outs, int count, int n_sms, cudaStream_t st) {
    if (count < 1 || count > MAXG) return 1;
    const int T = (int)(n_pad / NB);
    int C = 1;
    while (C * 2 <= g_max_cluster && C * 4 <= T) C *= 2;   // at least two blocks of the vector per CTA
    // Wide clusters have to find their SMs inside one GPC, so fewer of them are resident at once than the SM count suggests
    // (tools/solve_sweep.py, n = 4096, ms per launch: 8 outputs 0.99 at width 16 / 0.62 at 8; 16 outputs 1.17 at 8 / 0.84 at 4;
    // 32 outputs 0.98 at 4 / 1.27 at 2): beyond width 4, keep clusters x width within 64 CTAs.
    while (C > 4 && C * count > 64) C /= 2;
    while (C > 1 && C * count > n_sms) C /= 2;             // many outputs: all clusters resident at once beats wide clusters
    {
        const char* e = getenv("MOGP_SOLVE_CLUSTER");      // (tuning aid: tools/solve_sweep.py)
        if (e && atoi(e) >= 1 && atoi(e) <= g_max_cluster && (atoi(e) & (atoi(e) - 1)) == 0) C = atoi(e);
    }
    const int nown = (T + C - 1) / C;
    const size_t smem = (size_t)(nown * NB + C * NB + NB + 4 * NB + 16) * sizeof(double);
    if (smem > 200 * 1024) return 2;
    SolveParams sp{};
    sp.A = A_slab; sp.Dinv = Dinv_slab; sp.Y = Y; sp.z = z; sp.alpha = alpha; sp.scal = scal; sp.info = info;
    sp.n_pad = n_pad; sp.T = T;
    for (int i = 0; i < count; i++) sp.outs[i] = outs[i];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(C * count);
    cfg.blockDim = dim3(SOLVE_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, solve_alpha_kernel, sp);
    return e == cudaSuccess ? 0 : 1;
}

}  // namespace mogp
