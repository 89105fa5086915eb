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

}  // namespace mogp

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
