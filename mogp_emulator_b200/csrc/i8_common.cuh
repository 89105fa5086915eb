// Helpers shared by the int8 (tcgen05) kernels of libmogp_b200: the predict TRSM (trsm_i8.cu) and the Cholesky whose trailing
// updates run on the same tensor cores (chol.cu, chol_i8_kernel).  Plane layout, instruction / descriptor encodings (checked by
// tools/probe_i8.cu), TMEM access, digit extraction.
#pragma once
#include "common.cuh"

namespace mogp {

constexpr int I8_BITS = 7;                    // bits per plane (signed digit in [-64, 64])
constexpr int I8_BN = 64;                     // test points per tile (MMA N per accumulator)
constexpr int I8_APLANE = NB * 32;            // bytes of one plane of a K = 32 step of L (128 rows)
constexpr int I8_BPLANE = I8_BN * 32;         // ... of V (64 columns)
constexpr int I8_NCW = 8;                     // consumer warps
constexpr int I8_THREADS = (I8_NCW + 3) * 32; // + MMA warp + loader warp + inv(L_ii) loader warp
constexpr int I8_QN = 4;                      // ticket queue depth
constexpr int I8_DCHUNK = NB * KC * 8;        // one K chunk (16 columns) of inv(L_ii): 16 KB, K-blocked
constexpr int I8_NCHECK = 32;                 // test points per output re-solved in FP64 by the a-posteriori accuracy check

constexpr int I8_LP = 8;                      // planes of L STORED per element (the Cholesky's tcgen05 path produces and reads 8,
                                              // the predict TRSM reads the leading S <= 8 of them: a truncation to S signed digits
                                              // has the error bound of a rounding to S digits, 2^-(7S+1))
constexpr int I8_LSTAGE = I8_LP * I8_APLANE;  // bytes of one K = 32 step of a 128 x 128 block of L (all stored planes)
constexpr int I8_LBLOCK = 4 * I8_LSTAGE;      // bytes of one 128 x 128 block of L

// byte offset of element (row r, k in [0, 32)) inside a plane of a K = 32 step: K-major, no swizzle, 8 x 16-byte core
// matrices; leading (K half) byte offset 128, stride (8-row group) byte offset 256 (checked by tools/probe_i8.cu)
__host__ __device__ __forceinline__ int i8_plane_off(int r, int kk) {
    return (r >> 3) * 256 + ((kk >> 4) & 1) * 128 + (r & 7) * 16 + (kk & 15);
}

__device__ __forceinline__ uint64_t i8_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
// D = s32, A = B = signed 8 bit, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t i8_idesc(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(NB >> 4) << 24);
}

// The issuing thread is chosen with elect.sync (common.cuh elect_one_sync): ptxas then knows a single lane is active and
// moves the operands to uniform registers with plain R2UR; behind `if (lane == 0)` every tcgen05.mma was wrapped in an ELECT /
// R2UR.BROADCAST / branch loop and the issue thread, not the tensor pipe, set the pace (profiles/r02_i8_timeline_before.txt).
__device__ __forceinline__ bool i8_elect_one() { return elect_one_sync(); }
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
__device__ __forceinline__ void i8_bulk_store(void* dst_global, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_global), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void i8_bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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

// S signed 7-bit digits of v 2^-es in [-0.99, 0.99] in INTEGER arithmetic (FP64 instructions share the tensor datapath with the
// int8 MMAs on this chip -- tools/probe_concurrency.cu -- so the epilogue keeps them to a minimum): q = round(v 2^(7S - es))
// from the bits of v, then d_t = round(q / 2^(7(S-t))) top down, q -= d_t 2^(7(S-t)).  sum_t d_t 2^-7t = q 2^-7S exactly;
// |d_t| <= 64 (|d_1| <= 127 for out-of-range input, which degrades instead of wrapping int8).  Differs from the FP64 form
// below only in how exact ties round.
template <int S>
__device__ __forceinline__ void i8_digits_int(double v, int es, int8_t (&dig)[S]) {
    const long long bits = __double_as_longlong(v);
    const int e = (int)((bits >> 52) & 0x7FF);
    const unsigned long long mant = ((unsigned long long)bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
    const int sh = 1075 + es - 7 * S - e;                  // v 2^(7S - es) = mant 2^-sh; |v 2^-es| < 1: sh >= 53 - 7 S
    long long q = 0;
    if (e != 0 && sh < 64) {
        const long long qmax = 127ll << (7 * (S - 1));
        if (sh >= 1) q = (long long)((mant + (1ull << (sh - 1))) >> sh);
        else if (sh >= -9) q = (long long)(mant << (-sh));  // S = 8: the grid 2^-56 is finer than the ulp of the large entries (exact)
        else q = qmax;                                     // far out of range (or NaN / inf)
        q = q > qmax ? qmax : q;
        q = bits < 0 ? -q : q;
    }
#pragma unroll
    for (int t = 1; t < S; t++) {
        const int shift = 7 * (S - t);
        const long long d = (q + (1ll << (shift - 1))) >> shift;
        q -= d << shift;
        dig[t - 1] = (int8_t)d;
    }
    dig[S - 1] = (int8_t)q;
}

// The same digits in FP64 (the slicing pass of L, a separate kernel):  x = sum_t d_t 2^-7t + O(2^-(7S+1)), every step exact.
template <int S>
__device__ __forceinline__ void i8_digits(double x, int8_t (&dig)[S]) {
    x = fmin(fmax(x, -0.99), 0.99);        // in-range data has |x| <= 0.99; out-of-range input degrades instead of wrapping int8
    double y = x;
#pragma unroll
    for (int t = 0; t < S; t++) {
        y *= 128.0;
        const double dd = rint(y);
        y -= dd;
        dig[t] = (int8_t)(int)dd;
    }
}

}  // namespace mogp
