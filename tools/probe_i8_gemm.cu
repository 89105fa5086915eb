// Hardware probe (not product code), step 2 of the round-2 study: an FP64-equivalent GEMM tile from int8 slices on tcgen05.
//   C (128 x 128, FP64) = A (128 x K) * B (128 x K)^T,  K = 128 * J, operands pre-sliced on the host into 6 signed 7-bit planes
//   (one power-of-two scale per row), planes stored in the canonical K-major core-matrix order so that one k32 stage of one
//   operand is a contiguous block (cp.async.bulk, no tensor map).  TMEM holds 3 weight groups of s32 accumulators
//   (3 x 128 columns) at a time, so the slice pairs are issued in two passes over the operands:
//     pass A: t + u <= 4  (6 MMAs per k32 step, planes 1..3)      pass B: 5 <= t + u <= 7  (15 MMAs per k32 step, planes 1..6)
//   Each pass ends with a TMEM -> FP64 recombination.  Checks the result against an FP64 reference and times the MMA/load
//   pipeline on every SM (all CTAs read the same L2-resident operands: the measured rate is min(MMA issue, L2 -> SM feed)).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_i8_gemm tools/probe_i8_gemm.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int S = 6, BITS = 7, NS = 4;
constexpr int PLANE = 128 * 32;            // bytes of one plane of one k32 stage (128 rows x 32 k)
constexpr int STAGE = 2 * S * PLANE;       // A planes then B planes

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, no swizzle: LBO 128 B, SBO 256 B, version 1
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// planes in global memory: [j*4 + ks][plane t][PLANE bytes]  (A and B separately)
__global__ void __launch_bounds__(320, 1)
gemm_kernel(const int8_t* __restrict__ Aq, const int8_t* __restrict__ Bq, const int* __restrict__ eA, const int* __restrict__ eB,
            int J, double* __restrict__ C, long long* __restrict__ cycles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t full[NS], empty[NS], acc_full, acc_empty;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const int nstage = J * 4;
    const long long t0 = clock64();

    if (warp == 9 && lane == 0) {
        // ---- loader: one contiguous block of planes per operand and k32 stage ----
        int it = 0;
        for (int pass = 0; pass < 2; pass++) {
            const uint32_t np = pass == 0 ? 3 : S;          // planes needed by this pass
            for (int st = 0; st < nstage; st++, it++) {
                const int slot = it % NS;
                if (it >= NS) mbar_wait(&empty[slot], (uint32_t)(((it / NS) - 1) & 1));
                unsigned char* dst = smem + slot * STAGE;
                mbar_expect(&full[slot], 2 * np * PLANE);
                bulk_load(dst, Aq + (size_t)st * S * PLANE, np * PLANE, &full[slot]);
                bulk_load(dst + S * PLANE, Bq + (size_t)st * S * PLANE, np * PLANE, &full[slot]);
            }
        }
    } else if (warp == 8 && lane == 0) {
        // ---- MMA issuer ----
        int it = 0;
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) {   // the recombination of pass A must have drained the accumulators
                mbar_wait(&acc_empty, 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const int wlo = pass == 0 ? 2 : 5, whi = pass == 0 ? 4 : 7;
            for (int st = 0; st < nstage; st++, it++) {
                const int slot = it % NS;
                mbar_wait(&full[slot], (uint32_t)((it / NS) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(smem + slot * STAGE), b0 = a0 + S * PLANE;
                for (int w = wlo; w <= whi; w++) {
                    bool first = (st == 0);
                    for (int t = 1; t < w; t++) {
                        const int u = w - t;
                        if (t > S || u > S) continue;
                        mma_i8(tmem + (uint32_t)(w - wlo) * 128, make_desc(a0 + (t - 1) * PLANE), make_desc(b0 + (u - 1) * PLANE),
                               first ? 0u : 1u);
                        first = false;
                    }
                }
                mma_commit(&empty[slot]);      // frees the smem slot when these MMAs have read it
            }
            mma_commit(&acc_full);             // all MMAs of the pass are complete
        }
    } else if (warp < 8) {
        // ---- recombination: thread = (accumulator row, 64-column half) ----
        const int r = (warp & 3) * 32 + lane, h = warp >> 2;
        const int ea = eA[r];
        for (int pass = 0; pass < 2; pass++) {
            mbar_wait(&acc_full, (uint32_t)pass);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int wlo = pass == 0 ? 2 : 5;
            for (int c0 = 0; c0 < 64; c0 += 8) {
                double acc[8];
#pragma unroll
                for (int j = 0; j < 8; j++) acc[j] = 0.0;
                for (int g = 0; g < 3; g++) {
                    uint32_t v[8];
                    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(g * 128 + h * 64 + c0);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const double wgt = ldexp(1.0, -BITS * (wlo + g));
#pragma unroll
                    for (int j = 0; j < 8; j++) acc[j] = fma((double)(int32_t)v[j], wgt, acc[j]);
                }
                if (blockIdx.x == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const int c = h * 64 + c0 + j;
                        const double val = ldexp(acc[j], ea + eB[c]);
                        double* dst = C + (size_t)r * 128 + c;
                        *dst = (pass == 0) ? val : (*dst + val);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty);
        }
    }
    __syncthreads();
    if (tid == 0 && cycles) cycles[blockIdx.x] = clock64() - t0;
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// host slicing (same scheme as tools/ozaki_study.py), planes written in the canonical core-matrix order
static void slice_operand(const std::vector<double>& X, int K, std::vector<int8_t>& out, std::vector<int>& ex) {
    const int J4 = K / 32;
    out.assign((size_t)J4 * S * PLANE, 0);
    ex.resize(128);
    for (int r = 0; r < 128; r++) {
        double mx = 0;
        for (int k = 0; k < K; k++) mx = fmax(mx, fabs(X[(size_t)r * K + k]));
        int e; frexp(mx + 1e-300, &e);
        e += 1;               // scaled values in (-0.5, 0.5): every signed digit fits int8 (|d| <= 64)
        ex[r] = e;
        for (int k = 0; k < K; k++) {
            double rem = ldexp(X[(size_t)r * K + k], -e);
            const int st = k / 32, kk = k % 32;
            const size_t off = (size_t)(r >> 3) * 256 + ((kk >> 4) & 1) * 128 + (r & 7) * 16 + (kk & 15);
            for (int t = 1; t <= S; t++) {
                const double d = nearbyint(ldexp(rem, BITS * t));
                out[((size_t)st * S + (t - 1)) * PLANE + off] = (int8_t)d;
                rem -= ldexp(d, -BITS * t);
            }
        }
    }
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    printf("device %s SMs=%d\n", p.name, nsm);
    for (int J : {4, 32}) {
        const int K = 128 * J;
        std::vector<double> A((size_t)128 * K), B((size_t)128 * K);
        uint64_t s = 12345;
        auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)((s >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 53) - 0.5; };
        for (int r = 0; r < 128; r++)
            for (int k = 0; k < K; k++) {
                A[(size_t)r * K + k] = rnd() * exp(-0.002 * abs(r * 32 - k)) * (1.0 + r);     // decaying rows, different scales
                B[(size_t)r * K + k] = rnd() * (0.1 + 0.01 * r);
            }
        std::vector<int8_t> Aq, Bq; std::vector<int> ea, eb;
        slice_operand(A, K, Aq, ea);
        slice_operand(B, K, Bq, eb);
        int8_t *dA, *dB; int *dea, *deb; double* dC; long long* dcy;
        CK(cudaMalloc(&dA, Aq.size())); CK(cudaMalloc(&dB, Bq.size())); CK(cudaMalloc(&dea, 512)); CK(cudaMalloc(&deb, 512));
        CK(cudaMalloc(&dC, sizeof(double) * 128 * 128)); CK(cudaMalloc(&dcy, sizeof(long long) * nsm));
        CK(cudaMemcpy(dA, Aq.data(), Aq.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bq.data(), Bq.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dea, ea.data(), 512, cudaMemcpyHostToDevice)); CK(cudaMemcpy(deb, eb.data(), 512, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)NS * STAGE + 256;
        CK(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gemm_kernel<<<nsm, 320, smem>>>(dA, dB, dea, deb, J, dC, dcy);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        gemm_kernel<<<nsm, 320, smem>>>(dA, dB, dea, deb, J, dC, dcy);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<double> Cd(128 * 128); std::vector<long long> cy(nsm);
        CK(cudaMemcpy(Cd.data(), dC, sizeof(double) * Cd.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cy.data(), dcy, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
        double max_err = 0, max_rel_bound = 0;
        for (int i = 0; i < 128; i++)
            for (int j = 0; j < 128; j++) {
                long double ref = 0, absum = 0;
                for (int k = 0; k < K; k++) { ref += (long double)A[(size_t)i * K + k] * B[(size_t)j * K + k]; absum += fabsl((long double)A[(size_t)i * K + k] * B[(size_t)j * K + k]); }
                const double err = fabs((double)(ref - Cd[i * 128 + j]));
                max_err = fmax(max_err, err);
                max_rel_bound = fmax(max_rel_bound, err / (double)absum);
            }
        const double flop = 2.0 * 128 * 128 * K * (double)nsm;
        printf("K=%5d: max |C - ref| %.3e, max error / sum|a_k b_k| %.3e;  %d CTAs in %.3f ms = %.1f TFLOP/s FP64-equivalent (SM0: %lld cycles, %.0f cycles per 128-k block)\n",
               K, max_err, max_rel_bound, nsm, ms, flop / ms * 1e-9, cy[0], (double)cy[0] / J);
        cudaFree(dA); cudaFree(dB); cudaFree(dea); cudaFree(deb); cudaFree(dC); cudaFree(dcy);
    }
    printf("probe done\n");
    return 0;
}
