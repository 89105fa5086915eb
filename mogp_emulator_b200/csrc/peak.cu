// Device-peak probe used by bench.py for the roofline denominator: issue-bound DMMA (mma.sync m16n8k8 f64)
// from registers on every SM.  MEASURED_PEAKS.json carries HBM and bf16 figures only, so the FP64 tensor
// peak is measured in the same process as the benchmark.
#include "../../include/mogp_b200.h"
#include "common.cuh"

namespace mogp {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double acc[8][4];
    const double a0 = 1e-3 * threadIdx.x, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3;
    const double b0 = 1e-3 * (threadIdx.x & 7), b1 = b0 + 1e-3;
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[j][e] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) dmma_16x8x8(acc[j], a0, a1, a2, a3, b0, b1);
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) s += acc[j][e];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// int8 tcgen05 issue peak: back-to-back kind::i8 MMAs (M = 128, K = 32, N = NN) from resident shared-memory operands into
// TMEM, one issuing thread per SM.  The operand values do not matter (s32 accumulation wraps silently).
template <int NN>
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
    __shared__ __align__(128) unsigned char ops[NB * 32 + NN * 32];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NB * 32 + NN * 32; i += blockDim.x) ops[i] = (unsigned char)(i * 7 + 1);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(&done, 1);
        fence_mbar_init();
    }
    fence_proxy_async();      // generic-proxy operand writes -> async-proxy (tensor core) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(ops), b0 = a0 + NB * 32;
        const uint64_t ad = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t bd = (uint64_t)((b0 >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(NB >> 4) << 24);
        for (int it = 0; it < iters; it++)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 0 ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
        const unsigned long long t0 = globaltimer_ns();
        while (!mbar_try_wait(&done, 0))
            if (globaltimer_ns() - t0 > 10000000000ull) __trap();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
    }
}

template <int NN>
static float time_i8_peak(int sms, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    i8_peak_kernel<NN><<<sms, 128>>>(iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        i8_peak_kernel<NN><<<sms, 128>>>(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

}  // namespace mogp

// tops[0]: int8 tensor peak of the chip (N = 256 MMAs), tops[1]: issue-bound rate of the M128 N64 K32 shape the predict
// kernel uses (its six / seven s32 accumulators of a 128 x 64 tile are all TMEM holds) -- both in TOP/s
extern "C" int mogp_peak_i8(int32_t device, int32_t iters, double* tops) {
    using namespace mogp;
    if (!tops || iters < 1) return MOGP_ERR_ARG;
    int sms = 0;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        set_error("mogp_peak_i8: device %d not available", device);
        return MOGP_ERR_CUDA;
    }
    const float ms256 = time_i8_peak<256>(sms, iters), ms64 = time_i8_peak<64>(sms, iters);
    if (cudaGetLastError() != cudaSuccess) {
        set_error("mogp_peak_i8: kernel failed");
        return MOGP_ERR_CUDA;
    }
    tops[0] = 2.0 * NB * 256 * 32 * (double)iters * sms / (ms256 * 1e-3) * 1e-12;
    tops[1] = 2.0 * NB * 64 * 32 * (double)iters * sms / (ms64 * 1e-3) * 1e-12;
    return MOGP_OK;
}

extern "C" int mogp_peak_dmma(int32_t device, int32_t iters, double* tflops) {
    using namespace mogp;
    if (!tflops || iters < 1) return MOGP_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("mogp_peak_dmma: device %d not available", device);
        return MOGP_ERR_CUDA;
    }
    const int blocks = prop.multiProcessorCount * 2, threads = 256;
    double* out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return MOGP_ERR_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dmma_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) {
        set_error("mogp_peak_dmma: kernel failed");
        return MOGP_ERR_CUDA;
    }
    const double flop = 2.0 * 1024.0 * 8.0 * (double)iters * blocks * (threads / 32);
    *tflops = flop / (best * 1e-3) * 1e-12;
    return MOGP_OK;
}
