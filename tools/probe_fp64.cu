// Hardware probe (not product code): FP64 pipe facts on B200 that the design depends on.
//   - DFMA and DMMA (mma.sync f64) throughput per shape, exp() throughput
//   - cuBLAS DGEMM TF/s (the measured "FP64 dense peak" denominator)
//   - a TMA 3-D tensor-map load of a K-blocked operand tile (layout sanity)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe tools/probe_fp64.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dfma_kernel(double* out, int iters) {
    double a[8]; double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
    }
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_kernel(double* out, int iters) {
    double a[4];
    for (int i = 0; i < 4; i++) a[i] = -(threadIdx.x * 1e-3 + i);
    double s = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) { s += exp(a[i]); a[i] -= 1e-3; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
__global__ void dmma_kernel(double* out, int iters) {
    // SHAPE: 0 = m8n8k4, 1 = m16n8k4, 2 = m16n8k8, 3 = m16n8k16
    double acc[NACC][4];
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = 1e-3 * (threadIdx.x + i);
    for (int i = 0; i < 4; i++) b[i] = 1e-3 * (threadIdx.x - i);
    for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) acc[j][i] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < NACC; j++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[j][0]), "+d"(acc[j][1]) : "d"(a[0]), "d"(b[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(acc[j][0]), "+d"(acc[j][1]), "+d"(acc[j][2]), "+d"(acc[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(acc[j][0]), "+d"(acc[j][1]), "+d"(acc[j][2]), "+d"(acc[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(acc[j][0]), "+d"(acc[j][1]), "+d"(acc[j][2]), "+d"(acc[j][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    double s = 0;
    for (int j = 0; j < NACC; j++) for (int i = 0; i < 4; i++) s += acc[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// correctness probe of the fragment layout assumption for m16n8k8 (row.col):
// A frag a0:(g,t) a1:(g+8,t) a2:(g,t+4) a3:(g+8,t+4); B frag b0:(k=t,n=g) b1:(k=t+4,n=g);
// C frag c0:(g,2t) c1:(g,2t+1) c2:(g+8,2t) c3:(g+8,2t+1)
__global__ void dmma_layout_kernel(const double* A, const double* B, double* C) {
    int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    double b0 = B[g * 8 + t], b1 = B[g * 8 + t + 4];   // B stored [n][k]
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    C[g * 8 + 2 * t] = c0; C[g * 8 + 2 * t + 1] = c1; C[(g + 8) * 8 + 2 * t] = c2; C[(g + 8) * 8 + 2 * t + 1] = c3;
}

// ---- TMA probe: 3-D tensor map over a row-major matrix M[rows][ld] viewed as (k_in=8, row, k_out) ----
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tm, double* out, int row0, int k0, int* status) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* tile = reinterpret_cast<double*>(smem);            // [4][128][8]
    __shared__ __align__(8) unsigned long long bar;
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    uint32_t tile_a = (uint32_t)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(4 * 128 * 8 * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(tile_a), "l"(&tm), "r"(0), "r"(row0), "r"(k0 / 8), "r"(bar_a) : "memory");
    }
    // bounded wait
    uint32_t done = 0; long long spins = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar_a), "r"(0) : "memory");
        if (++spins > 2000000) { if (threadIdx.x == 0) *status = 1; break; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 128 * 8; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F> float time_kernel(F f, int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d SMs=%d smem/block optin=%zu L2=%d MB clock=%d kHz\n", p.name, p.major, p.minor,
           p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize >> 20, p.clockRate);
    int nsm = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));

    {   // DFMA
        for (int warps : {4, 8, 16, 32}) {
            int iters = 20000; int blocks = nsm * 2;
            float ms = time_kernel([&] { dfma_kernel<<<blocks, warps * 32 / 2>>>(out, iters); });
            double flop = 2.0 * 8 * iters * (double)blocks * (warps * 32 / 2);
            printf("DFMA warps/SM=%2d : %.2f TFLOP/s\n", warps, flop / ms * 1e-9);
        }
    }
    {   // exp
        int iters = 4000, blocks = nsm * 2, thr = 512;
        float ms = time_kernel([&] { exp_kernel<<<blocks, thr>>>(out, iters); });
        double n = 4.0 * iters * blocks * thr;
        printf("exp(double): %.1f Gexp/s  (= %.1f DFMA-equiv per exp at 18.6T DFMA/s nominal)\n", n / ms * 1e-6, 18.6e12 / (n / ms * 1e3));
    }
#define RUN_DMMA(SHAPE, NACC, NAME, FMA) \
    for (int warps : {4, 8, 16}) { int iters = 4000; int blocks = nsm; \
        float ms = time_kernel([&] { dmma_kernel<SHAPE, NACC><<<blocks, warps * 32>>>(out, iters); }); \
        double flop = 2.0 * FMA * NACC * iters * (double)blocks * warps; \
        printf("DMMA %-9s nacc=%2d warps/SM=%2d : %.2f TFLOP/s\n", NAME, NACC, warps, flop / ms * 1e-9); }
    RUN_DMMA(0, 8, "m8n8k4", 256.0)
    RUN_DMMA(0, 16, "m8n8k4", 256.0)
    RUN_DMMA(1, 8, "m16n8k4", 512.0)
    RUN_DMMA(2, 8, "m16n8k8", 1024.0)
    RUN_DMMA(2, 16, "m16n8k8", 1024.0)
    RUN_DMMA(3, 8, "m16n8k16", 2048.0)
    RUN_DMMA(3, 16, "m16n8k16", 2048.0)
    RUN_DMMA(2, 1, "m16n8k8", 1024.0)
    RUN_DMMA(2, 2, "m16n8k8", 1024.0)
    RUN_DMMA(2, 4, "m16n8k8", 1024.0)

    {   // fragment layout check
        std::vector<double> A(16 * 8), B(8 * 8), C(16 * 8), Cr(16 * 8, 0.0);
        for (int i = 0; i < 16 * 8; i++) A[i] = (i * 7 % 13) - 6;
        for (int i = 0; i < 8 * 8; i++) B[i] = (i * 5 % 11) - 5;
        for (int i = 0; i < 16; i++) for (int j = 0; j < 8; j++) for (int k = 0; k < 8; k++) Cr[i * 8 + j] += A[i * 8 + k] * B[j * 8 + k];
        double *dA, *dB, *dC; cudaMalloc(&dA, 16 * 8 * 8); cudaMalloc(&dB, 8 * 8 * 8); cudaMalloc(&dC, 16 * 8 * 8);
        cudaMemcpy(dA, A.data(), 16 * 8 * 8, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), 8 * 8 * 8, cudaMemcpyHostToDevice);
        dmma_layout_kernel<<<1, 32>>>(dA, dB, dC); CK(cudaDeviceSynchronize());
        cudaMemcpy(C.data(), dC, 16 * 8 * 8, cudaMemcpyDeviceToHost);
        double err = 0; for (int i = 0; i < 16 * 8; i++) err = fmax(err, fabs(C[i] - Cr[i]));
        printf("m16n8k8 fragment layout check: max err %.3g (%s)\n", err, err == 0 ? "OK" : "MISMATCH");
    }
    {   // cuBLAS DGEMM
        cublasHandle_t h; cublasCreate(&h);
        for (int n : {2048, 4096, 8192}) {
            double *A, *B, *C; size_t bytes = sizeof(double) * n * n;
            CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
            CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
            double one = 1.0, zero = 0.0;
            float ms = time_kernel([&] { cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, 5);
            printf("cublasDgemm TN n=%d: %.3f ms  %.2f TFLOP/s\n", n, ms, 2.0 * n * n * (double)n / ms * 1e-9);
            float ms2 = time_kernel([&] { cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n, n, &one, A, n, &zero, C, n); }, 5);
            printf("cublasDsyrk n=%d k=%d: %.3f ms  %.2f TFLOP/s\n", n, n, ms2, 1.0 * n * n * (double)n / ms2 * 1e-9);
            cudaFree(A); cudaFree(B); cudaFree(C);
        }
        cublasDestroy(h);
    }
    {   // TMA probe
        const int rows = 512, ld = 256;
        std::vector<double> M((size_t)rows * ld);
        for (int r = 0; r < rows; r++) for (int c = 0; c < ld; c++) M[(size_t)r * ld + c] = r * 1000.0 + c;
        double* dM; CK(cudaMalloc(&dM, M.size() * 8)); CK(cudaMemcpy(dM, M.data(), M.size() * 8, cudaMemcpyHostToDevice));
        EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
        CUtensorMap tm;
        cuuint64_t gdim[3] = {8, (cuuint64_t)rows, (cuuint64_t)ld / 8};
        cuuint64_t gstr[2] = {(cuuint64_t)ld * 8, 64};
        cuuint32_t box[3] = {8, 128, 4};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, dM, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
        if (r == CUDA_SUCCESS) {
            double* dT; CK(cudaMalloc(&dT, 4 * 128 * 8 * 8)); int* dS; CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4));
            int row0 = 128, k0 = 64;
            tma_probe_kernel<<<1, 128, 4 * 128 * 8 * 8>>>(tm, dT, row0, k0, dS);
            cudaError_t e = cudaDeviceSynchronize();
            printf("tma kernel: %s\n", cudaGetErrorString(e));
            std::vector<double> T(4 * 128 * 8); int st = 0;
            cudaMemcpy(T.data(), dT, T.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int ko = 0; ko < 4; ko++) for (int rr = 0; rr < 128; rr++) for (int ki = 0; ki < 8; ki++) {
                double want = (row0 + rr) * 1000.0 + (k0 + ko * 8 + ki);
                if (T[(ko * 128 + rr) * 8 + ki] != want) bad++;
            }
            printf("TMA 3D K-blocked tile: timeout=%d mismatches=%d (%s)\n", st, bad, (bad == 0 && st == 0) ? "OK" : "FAIL");
            // OOB probe: rows beyond extent must be zero-filled
            CK(cudaMemset(dS, 0, 4));
            tma_probe_kernel<<<1, 128, 4 * 128 * 8 * 8>>>(tm, dT, 448, 240, dS);
            e = cudaDeviceSynchronize();
            cudaMemcpy(T.data(), dT, T.size() * 8, cudaMemcpyDeviceToHost); cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
            bad = 0;
            for (int ko = 0; ko < 4; ko++) for (int rr = 0; rr < 128; rr++) for (int ki = 0; ki < 8; ki++) {
                int R = 448 + rr, Cc = 240 + ko * 8 + ki;
                double want = (R < rows && Cc < ld) ? R * 1000.0 + Cc : 0.0;
                if (T[(ko * 128 + rr) * 8 + ki] != want) bad++;
            }
            printf("TMA OOB zero-fill: %s timeout=%d mismatches=%d\n", cudaGetErrorString(e), st, bad);
        }
    }
    printf("probe done\n");
    return 0;
}
