// Shared device/host helpers for libmogp_b200 (sm_100a only).
//
// Everything here is FP64.  The B200 FP64 tensor pipe is the warp-level DMMA
// (mma.sync.aligned.m16n8k8.f64); tcgen05.mma has no f64 kind (CUDA 12.9 PTX), so the dense
// contractions of this library are DMMA kernels whose operand tiles are staged by TMA
// (cp.async.bulk.tensor) into a K-blocked shared-memory layout and consumed through an
// mbarrier full/empty ring by warp-specialised consumer warps.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda.h>

namespace mogp {

constexpr int NB = 128;            // Cholesky block size / tile edge (rows of L per block row)
constexpr int KC = 16;             // K-chunk (doubles) per pipeline stage
constexpr int KSLAB = 8;           // inner K slab: smem tiles are [K/8][rows][8]

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 3-D tiled TMA load: coordinates (c0 = k_in, c1 = row, c2 = k_out)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// 2-D tiled TMA load: coordinates (c0 = fastest dim, c1)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// generic-proxy writes (st.global / st.shared) -> visible to subsequent async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// register re-balancing between the TMA producer warpgroup and the DMMA consumer warpgroups
template <int R>
__device__ __forceinline__ void reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R));
}
template <int R>
__device__ __forceinline__ void reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R));
}

// One lane of a converged warp, chosen by elect.sync: ptxas then knows a single thread runs the branch and feeds the uniform
// datapath of TMA / tcgen05 instructions with plain R2UR (behind `if (lane == 0)` every such instruction is wrapped in an
// ELECT / R2UR.BROADCAST / branch loop -- profiles/r02_sass_evidence.txt).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- cross-CTA dataflow helpers (ticket counters / ready flags in global memory) ----
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin until *p >= target.  A producer thread that then issues TMA loads of the published data must follow up
// with fence_proxy_async() (the data was written through the generic proxy).
__device__ __forceinline__ void wait_counter(const int* p, int target) {
    if (ld_acquire_gpu(p) >= target) return;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_gpu(p) < target) {
        __nanosleep(100);
        if (globaltimer_ns() - t0 > 20000000000ull) __trap();   // 20 s: never in a correct run
    }
}

// D(16x8) += A(16x8, row) * B(8x8, col), FP64 tensor pipe (SASS: DMMA.16x8x8)
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], double a0, double a1, double a2, double a3, double b0,
                                            double b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

// ------------------------------------------------------------------------------------------
// Warp-level DMMA over one K-blocked smem stage.
//
// Operand tiles live in shared memory as [KCH/8][rows][8] doubles (each 8-wide K slab of a row is
// 64 contiguous bytes) -- exactly what one 3-D TMA box (8, rows, KCH/8) writes.  Lane (g = lane/4,
// t = lane%4) reads the 16-byte pair at K offsets (2t, 2t+1) of its row: the DMMA K index is
// permuted (logical t <-> physical 2t, logical t+4 <-> physical 2t+1) identically for A and B, which
// leaves the contraction unchanged and makes every fragment load a conflict-free LDS.128.
//
// acc[mt][nt][0..3] <-> (row = arow0 + 16*mt + g (+8 for 2,3), col = bcol0 + 8*nt + 2t (+1 for 1,3))
// ------------------------------------------------------------------------------------------
template <int MT, int NT, int KCH>
__device__ __forceinline__ void mma_stage(double (&acc)[MT][NT][4], const double* __restrict__ As, int a_rows,
                                          int arow0, const double* __restrict__ Bs, int b_rows, int bcol0, int g,
                                          int t) {
#pragma unroll
    for (int ks = 0; ks < KCH / 8; ks++) {
        double2 af[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            const double* p = As + ((size_t)(ks * a_rows + arow0 + mt * 16 + g) * 8 + 2 * t);
            af[mt][0] = *reinterpret_cast<const double2*>(p);
            af[mt][1] = *reinterpret_cast<const double2*>(p + 64);
        }
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
            const double2 bf =
                *reinterpret_cast<const double2*>(Bs + ((size_t)(ks * b_rows + bcol0 + nt * 8 + g) * 8 + 2 * t));
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
                dmma_16x8x8(acc[mt][nt], af[mt][0].x, af[mt][1].x, af[mt][0].y, af[mt][1].y, bf.x, bf.y);
        }
    }
}

// pipeline bookkeeping for an mbarrier ring of NS stages
template <int NS>
struct PipeState {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++stage == NS) {
            stage = 0;
            phase ^= 1u;
        }
    }
};

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
#define MOGP_CUDA_OK(expr)                                                                     \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            mogp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                            __LINE__);                                                         \
            return MOGP_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

void set_error(const char* fmt, ...);

// K-blocked 3-D tensor map over a row-major FP64 matrix M[rows][ld]: dims (8, rows, ld/8),
// box (8, box_rows, box_kslabs).  Returns 0 on success.
int make_kblocked_tmap(CUtensorMap* tm, const double* base, int64_t rows, int64_t ld, int box_rows, int box_kslabs = KC / 8);
// 2-D tensor map over a row-major FP64 matrix M[rows][cols] (cols contiguous): box (box_cols, box_rows)
int make_2d_tmap(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                 int box_cols);

}  // namespace mogp
