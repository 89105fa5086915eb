// Hardware probe (not product code) for a possible round-2 kernel: tcgen05.mma kind::i8 on B200.
//   1. exactness + layout check: D(128 x N, s32 in TMEM) = A(128 x K, s8) * B(N x K, s8)^T with hand-built shared-memory
//      descriptors (K-major, no swizzle, 8x16-byte core matrices) against a host reference
//   2. issue-rate measurement: back-to-back MMAs from resident operands, per N, all SMs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_i8 tools/probe_i8.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (between the two 16-byte K halves of one MMA),
//   [32,46) stride byte offset >> 4 (between 8-row groups), [46,48) version = 1, [61,64) layout type = 0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor (InstrDescriptor): c_format [4,6) = 2 (S32), a_format [7,10) = 1 (signed 8 bit), b_format [10,13) = 1,
// a_major/b_major = 0 (K), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// operand tile in shared memory: [k32 step][row group of 8][k half (16 B)][8 rows][16 B]
__device__ __forceinline__ int tile_off(int rows, int r, int k) {
    const int ks = k >> 5, kh = (k >> 4) & 1, kb = k & 15;
    return ks * rows * 32 + (r >> 3) * 256 + kh * 128 + (r & 7) * 16 + kb;
}

template <int N>
__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* A, const int8_t* B, int K, int32_t* D, int reps, long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    int8_t* sA = reinterpret_cast<int8_t*>(smem);
    int8_t* sB = sA + 128 * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * K; i += 128) sA[tile_off(128, i / K, i % K)] = A[i];
    for (int i = tid; i < N * K; i += 128) sB[tile_off(N, i / K, i % K)] = B[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // generic-proxy smem writes -> visible to the tensor core's async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = clock64();
    if (tid == 0) {
        for (int rep = 0; rep < reps; rep++) {
            for (int ks = 0; ks < K / 32; ks++) {
                const uint64_t ad = make_desc(smem_u32(sA + ks * 128 * 32), 128, 256);
                const uint64_t bd = make_desc(smem_u32(sB + ks * N * 32), 128, 256);
                mma_i8(tmem, ad, bd, idesc, (rep > 0 || ks > 0) ? 1u : 0u);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (accumulator rows), 8 columns at a time
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (blockIdx.x == 0)
            for (int j = 0; j < 8; j++) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
}

template <int N>
void run(int nsm) {
    const int K = 128;
    std::vector<int8_t> hA(128 * K), hB(N * K);
    for (size_t i = 0; i < hA.size(); i++) hA[i] = (int8_t)((int)((i * 37 + 11) % 129) - 64);
    for (size_t i = 0; i < hB.size(); i++) hB[i] = (int8_t)((int)((i * 53 + 7) % 127) - 63);
    int8_t *dA, *dB; int32_t* dD; long long* dC;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, sizeof(int32_t) * 128 * N));
    CK(cudaMalloc(&dC, sizeof(long long) * nsm));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(128 + N) * K + 128;
    CK(cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_kernel<N><<<1, 128, smem>>>(dA, dB, K, dD, 1, nullptr);
    CK(cudaDeviceSynchronize());
    std::vector<int32_t> hD(128 * N);
    CK(cudaMemcpy(hD.data(), dD, sizeof(int32_t) * hD.size(), cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < N; j++) {
            int32_t ref = 0;
            for (int k = 0; k < K; k++) ref += (int32_t)hA[i * K + k] * (int32_t)hB[j * K + k];
            if (ref != hD[i * N + j]) { if (bad < 3) printf("  mismatch (%d,%d): got %d want %d\n", i, j, hD[i * N + j], ref); bad++; }
        }
    printf("kind::i8 M=128 N=%3d K=%d: %s (%lld mismatches)\n", N, K, bad ? "WRONG" : "exact", bad);
    // issue rate: reps x (K/32) MMAs back to back on every SM
    const int reps = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    probe_kernel<N><<<nsm, 128, smem>>>(dA, dB, K, dD, reps, dC);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    probe_kernel<N><<<nsm, 128, smem>>>(dA, dB, K, dD, reps, dC);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> hc(nsm);
    CK(cudaMemcpy(hc.data(), dC, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    const double ops = 2.0 * 128 * N * K * (double)reps * nsm;
    const double ops_sm = 2.0 * 128 * N * K * (double)reps;
    printf("   %d SMs: %.3f ms -> %.1f TOP/s whole kernel; SM0 MMA span %lld cycles -> %.0f op/clk/SM\n", nsm, ms, ops / ms * 1e-9,
           hc[0], ops_sm / (double)hc[0]);
    cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    run<64>(p.multiProcessorCount);
    run<128>(p.multiProcessorCount);
    run<256>(p.multiProcessorCount);
    printf("probe done\n");
    return 0;
}
