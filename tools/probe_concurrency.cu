// Hardware probe (not product code), round 2: can FP64 work run beside a saturated int8 tcgen05 stream on the same SM?
//   The fused epilogue of csrc/trsm_i8.cu computes V_i = inv(L_ii) T_i in FP64 while the MMA warp issues kind::i8 MMAs of the
//   next tile.  ncu showed the DMMA instructions of that epilogue stalled on "math pipe throttle" at < 10 % DMMA utilisation.
//   This probe runs, per SM, one thread issuing back-to-back kind::i8 MMAs (M128 N256 K32, resident operands) and 8 warps of
//   FP64 work from registers, alone and together:
//     mode 1: only the int8 stream         mode 2: only DMMA (m16n8k8)        mode 4: only DFMA
//     mode 3: int8 + DMMA                  mode 5: int8 + DFMA                mode 6: DMMA + DFMA (4 warps each)
//   and prints the time of each part (the kernel ends when both parts are done; each part records its own elapsed clock64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_concurrency tools/probe_concurrency.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dmma(double (&c)[4], double a0, double a1, double a2, double a3, double b0, double b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

// warps 0-7: FP64 work (fp_kind: 0 none, 1 DMMA, 2 DFMA, 3 warps 0-3 DMMA + warps 4-7 DFMA); warp 8: int8 MMA issuer
__global__ void __launch_bounds__(288, 1) probe(double* out, long long* clocks, int mma_iters, int fp_iters, int fp_kind) {
    __shared__ __align__(128) unsigned char ops[128 * 32 + 256 * 32];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128 * 32 + 256 * 32; i += blockDim.x) ops[i] = (unsigned char)(i * 7 + 1);
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const long long t0 = clock64();
    if (warp == 8) {
        if (lane == 0 && mma_iters > 0) {
            const uint32_t a0 = smem_u32(ops), b0 = a0 + 128 * 32;
            const uint64_t ad = (uint64_t)((a0 >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
            const uint64_t bd = (uint64_t)((b0 >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int it = 0; it < mma_iters; it++)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(it > 0 ? 1u : 0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&done)), "r"(0) : "memory");
            clocks[blockIdx.x * 2] = clock64() - t0;
        }
    } else {
        const int kind = (fp_kind == 3) ? (warp < 4 ? 1 : 2) : fp_kind;
        double s = 0.0;
        if (kind == 1) {
            double acc[8][4];
            const double a0 = 1e-3 * threadIdx.x, a1 = a0 + 1e-3, a2 = a0 + 2e-3, a3 = a0 + 3e-3, b0 = 1e-3 * (threadIdx.x & 7), b1 = b0 + 1e-3;
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) acc[j][e] = 0.0;
            for (int it = 0; it < fp_iters; it++)
#pragma unroll
                for (int j = 0; j < 8; j++) dmma(acc[j], a0, a1, a2, a3, b0, b1);
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) s += acc[j][e];
        } else if (kind == 2) {
            double acc[16];
            const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9;
#pragma unroll
            for (int j = 0; j < 16; j++) acc[j] = j;
            // same flop count per iteration as the DMMA branch: 8 x 2048 flop per warp = 8 x 32 DFMA per thread ... 256 DFMA
            for (int it = 0; it < fp_iters; it++)
#pragma unroll
                for (int r = 0; r < 16; r++)
#pragma unroll
                    for (int j = 0; j < 16; j++) acc[j] = fma(acc[j], a, b);
#pragma unroll
            for (int j = 0; j < 16; j++) s += acc[j];
        }
        if (kind != 0 && lane == 0 && warp == 0) clocks[blockIdx.x * 2 + 1] = clock64() - t0;
        if (kind != 0) out[blockIdx.x * 256 + threadIdx.x] = s;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; long long* clk;
    CK(cudaMalloc(&out, sizeof(double) * sms * 256));
    CK(cudaMalloc(&clk, sizeof(long long) * sms * 2));
    const int mma_iters = 40000;                 // x 128 cycles each at the N = 256 rate
    const int fp_iters = 10000;                  // x 8 DMMA m16n8k8 per warp (8 warps): 1.31 GFLOP per SM
    struct { const char* name; int mi, fi, kind; } runs[] = {
        {"int8 MMA stream alone            ", mma_iters, 0, 0},
        {"DMMA alone (8 warps)             ", 0, fp_iters, 1},
        {"DFMA alone (8 warps)             ", 0, fp_iters, 2},
        {"int8 MMA stream + DMMA (8 warps) ", mma_iters, fp_iters, 1},
        {"int8 MMA stream + DFMA (8 warps) ", mma_iters, fp_iters, 2},
        {"DMMA (4 warps) + DFMA (4 warps)  ", 0, fp_iters, 3},
        {"DMMA alone (4 warps: fp_kind 3 without DFMA is not run; reference = half of the 8-warp run)", 0, 0, 0},
    };
    for (int r = 0; r < 6; r++) {
        CK(cudaMemset(clk, 0, sizeof(long long) * sms * 2));
        probe<<<sms, 288>>>(out, clk, runs[r].mi, runs[r].fi, runs[r].kind);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMemset(clk, 0, sizeof(long long) * sms * 2));
        CK(cudaEventRecord(e0));
        probe<<<sms, 288>>>(out, clk, runs[r].mi, runs[r].fi, runs[r].kind);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long h[2 * 200];
        CK(cudaMemcpy(h, clk, sizeof(long long) * sms * 2, cudaMemcpyDeviceToHost));
        double cm = 0, cf = 0;
        for (int i = 0; i < sms; i++) { cm += h[2 * i]; cf += h[2 * i + 1]; }
        cm /= sms; cf /= sms;
        const double mma_cyc = runs[r].mi ? cm / runs[r].mi : 0.0;                         // cycles per N=256 MMA (floor 128)
        const double warps = 8.0;
        const double fp_flop_per_clk = runs[r].fi ? (warps * runs[r].fi * 8.0 * 2048.0) / cf : 0.0;   // per SM (peak ~ 128)
        printf("%s: %.3f ms | int8: %.1f cycles per N=256 MMA | FP64: %.1f flop/clk/SM (%.1f TFLOP/s at 1.965 GHz x %d SMs)\n", runs[r].name, ms,
               mma_cyc, fp_flop_per_clk, fp_flop_per_clk * 1.965e9 * sms * 1e-12, sms);
    }
    printf("probe done\n");
    return 0;
}
