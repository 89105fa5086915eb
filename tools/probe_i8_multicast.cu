// Hardware probe (not product code) for round 2: is the operand-feed ceiling of i8_row_kernel (5.4 - 5.9 TB/s of TMA loads
// out of L2, profiles/r01_c3_i8_row*.txt) L2 read bandwidth or crossbar delivery?
//   Same per-CTA pipeline as csrc/trsm_i8.cu (S = 7 planes, M128 N64 K32, 28 kind::i8 MMAs per K step, 4-stage ring of
//   cp.async.bulk loads), run by thread-block clusters of CS CTAs that all need the SAME A (L~) planes and their own B (V)
//   planes:  mode 0: every CTA loads A itself (what the kernel does today);
//            mode 1: CTA r of the cluster loads the r-th 1/CS slice of A with .multicast::cluster to all CS CTAs.
//   If mode 1 is faster at equal delivered bytes, the ceiling is L2 read bandwidth and cluster multicast of the L~ stream is
//   the next step for the kernel; if not, it is delivery and only a larger tile per operand byte helps.
//   Operand values are irrelevant (timing only); A planes are shared by all clusters (L2-resident), B planes are per CTA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_i8_multicast tools/probe_i8_multicast.cu
// mode bit 1 (value 2): the wide-N MMA form (10 instructions per K step instead of 28).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int S = 7, NS = 4;
constexpr int ABYTES = S * 128 * 32, BBYTES = S * 64 * 32, STAGE = ABYTES + BBYTES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t idesc_n(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }
constexpr uint32_t IDESC = idesc_n(64);
__device__ __forceinline__ void mma_i8n(uint32_t d, uint64_t a, uint64_t b, uint32_t acc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) { mma_i8n(d, a, b, acc, IDESC); }
__device__ __forceinline__ void commit_local(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void commit_multicast(uint64_t* bar, uint16_t mask) {   // arrives on the barrier at this offset in every CTA of mask
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try(bar, parity)) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 5000000000ull) __trap();      // 5 s: a protocol bug must not hang the GPU
    }
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS>
__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* __restrict__ Aq, const int8_t* __restrict__ Bq, int nstage, int mode) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t full[NS], empty[NS], done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CS > 1) ? cluster_rank() : 0u;
    const uint16_t mask = (uint16_t)((1u << CS) - 1u);
    const bool mc = ((mode & 1) && CS > 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], mc ? CS : 1); }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) cluster_sync();          // every CTA's barriers exist before a peer multicasts into them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int8_t* b_src = Bq + (size_t)blockIdx.x * (size_t)BBYTES * 64;      // 64 distinct stages per CTA, reused cyclically
    if (warp == 1 && lane == 0) {
        for (int it = 0; it < nstage; it++) {
            const int slot = it % NS;
            if (it >= NS) mbar_wait(&empty[slot], (uint32_t)(((it / NS) - 1) & 1));
            unsigned char* dst = smem + slot * STAGE;
            mbar_expect(&full[slot], STAGE);
            const int8_t* a_src = Aq + (size_t)(it % 128) * ABYTES;
            if (mc) {
                const uint32_t part = ABYTES / CS;
                bulk_load_multicast(dst + rank * part, a_src + rank * part, part, &full[slot], mask);
            } else {
                bulk_load(dst, a_src, ABYTES, &full[slot]);
            }
            bulk_load(dst + ABYTES, b_src + (size_t)(it % 64) * BBYTES, BBYTES, &full[slot]);
        }
    } else if (warp == 0 && lane == 0) {
        for (int it = 0; it < nstage; it++) {
            const int slot = it % NS;
            mbar_wait(&full[slot], (uint32_t)((it / NS) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(smem + slot * STAGE), b0 = a0 + ABYTES;
            if (mode & 2) {
                // wide form: plane t of A against planes 1 .. S+1-t of V in ONE instruction chain (the planes of V are contiguous
                // in N, the accumulators contiguous in TMEM in weight order): 10 MMAs of N <= 256 instead of 28 of N = 64
#pragma unroll
                for (int t = 1; t <= S; t++) {
                    const int ncols = 64 * (S + 1 - t);
                    const uint32_t acc = (it == 0 && t == 1) ? 0u : 1u;
                    const uint32_t d0 = tmem + (uint32_t)(t - 1) * 64;
                    const uint64_t ad = make_desc(a0 + (t - 1) * 128 * 32);
                    const int n1 = ncols > 256 ? 256 : ncols;
                    mma_i8n(d0, ad, make_desc(b0), acc, idesc_n(n1));
                    if (ncols > 256) mma_i8n(d0 + 256, ad, make_desc(b0 + 4 * 64 * 32), acc, idesc_n(ncols - 256));
                }
            } else
#pragma unroll
            for (int w = 2; w <= S + 1; w++) {
                uint32_t acc = it == 0 ? 0u : 1u;
#pragma unroll
                for (int t = 1; t < w; t++) {
                    const int u = w - t;
                    if (t > S || u > S) continue;
                    mma_i8(tmem + (uint32_t)(w - 2) * 64, make_desc(a0 + (t - 1) * 128 * 32), make_desc(b0 + (u - 1) * 64 * 32), acc);
                    acc = 1u;
                }
            }
            if (mc) commit_multicast(&empty[slot], mask);      // the slot is rewritten by every CTA of the cluster
            else commit_local(&empty[slot]);
        }
        commit_local(&done);
        mbar_wait(&done, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) cluster_sync();          // no CTA leaves while a peer may still multicast into it / arrive on its barriers
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

template <int CS>
static void run(const int8_t* A, const int8_t* B, int nsm, int nstage, int mode) {
    const size_t smem = (size_t)NS * STAGE + 256;
    CK(cudaFuncSetAttribute(probe_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8) CK(cudaFuncSetAttribute(probe_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nsm / CS * CS));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, probe_kernel<CS>, A, B, nstage, mode));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, probe_kernel<CS>, A, B, nstage, mode));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double ctas = cfg.gridDim.x;
    const double delivered = ctas * nstage * (double)STAGE;
    const double l2_reads = ctas * nstage * (((mode & 1) && CS > 1 ? (double)ABYTES / CS : (double)ABYTES) + BBYTES);
    const double ops = ctas * nstage * 28.0 * 2.0 * 128 * 64 * 32;
    printf("cluster %d, %s, %s: %.3f ms  delivered %.2f TB/s  L2 requests %.2f TB/s  %.0f TOP/s int8 (%.1f cycles per MMA at 1.965 GHz)\n",
           CS, (mode & 1) ? "A multicast" : "A unicast  ", (mode & 2) ? "wide N (10 MMAs)  " : "narrow N (28 MMAs)", ms, delivered / ms * 1e-9, l2_reads / ms * 1e-9, ops / ms * 1e-9,
           ms * 1e-3 * 1.965e9 / (nstage * 28.0));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    printf("device %s SMs=%d\n", p.name, nsm);
    int8_t *A, *B;
    CK(cudaMalloc(&A, (size_t)128 * ABYTES));                       // 3.7 MB shared by everybody: L2-resident
    CK(cudaMalloc(&B, (size_t)nsm * 64 * BBYTES));                   // 136 MB: per-CTA streams
    CK(cudaMemset(A, 1, (size_t)128 * ABYTES)); CK(cudaMemset(B, 1, (size_t)nsm * 64 * BBYTES));
    const int nstage = 4096;
    for (int wide = 0; wide <= 2; wide += 2) {
        run<1>(A, B, nsm, nstage, wide);
        run<2>(A, B, nsm, nstage, wide);
        run<2>(A, B, nsm, nstage, wide | 1);
        run<4>(A, B, nsm, nstage, wide);
        run<4>(A, B, nsm, nstage, wide | 1);
        run<8>(A, B, nsm, nstage, wide | 1);
    }
    printf("probe done\n");
    return 0;
}
