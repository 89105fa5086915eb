// Kernel-matrix assembly for libmogp_b200 (replaces the scalar one-thread-per-element kernels
// sqexp_cov_batch_kernel / mat52_cov_batch_kernel, reference mogp_gpu/src/kernel.cu:55-65,251-261;
// arithmetic follows the CPU reference StationaryKernel.calc_r2 + calc_K, mogp_emulator/Kernel.py:476-485,
// 787-791, 878-882, and GaussianProcess.get_cov_matrix, GaussianProcess.py:517-558).
//
//   r2[i][j] = sum_d exp(theta_d) (x_i,d - x'_j,d)^2 ;  K = sigma2 * k(r2)
//
// The design matrices are kept transposed in HBM (XT: d x n_pad), so a 128-point tile is a TMA 2-D box
// (128 points x up-to-16 dims) landing in shared memory as [dim][point]: lanes read consecutive points
// (conflict-free) and the row operand is a broadcast.  Each thread owns an 8x8 sub-tile in registers
// and writes 128-byte row segments of K.
//   SYM   : lower 128x128 tiles of K(X,X) + nugget*I into the Cholesky workspace (identity in padding)
//   CROSS : test-major K*(Xs,X) into the predict workspace, with the fused posterior-mean partial
//           dot  sum_r K*[c][r] alpha[r]  (deterministic per-tile partials, no atomics).
#include "common.cuh"
#include "kernels.h"
#include "../../include/mogp_b200.h"

namespace mogp {

constexpr int DCH = 16;                                // dims per TMA box / smem chunk
constexpr int KM_TILE = DCH * 128 * 8;                 // one 128-point tile of one chunk of dims (16 KB)
constexpr int KM_STAGE = 2 * KM_TILE;                  // ring stage: row tile + column tile
constexpr int KM_SMEM = 2 * KM_STAGE + 128;
constexpr int KM_THREADS = 256;
constexpr int KM_ROWS = 64;                            // rows of a tile (a tile is 64 x 128)
constexpr int KM_RA = KM_ROWS / (KM_THREADS / 16);     // rows per thread (4)

// 2^(j/32), correctly rounded
__device__ const double kExp2Table[32] = {
    1, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237,
    1.0905077326652577, 1.1143867425958924, 1.1387886347566916, 1.1637248587775775,
    1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332,
    1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228,
    1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.6457554781539649,
    1.681792830507429, 1.7186192981224779, 1.7562521603732995, 1.7947090750031072,
    1.8340080864093424, 1.8741676341103, 1.9152065613971474, 1.9571441241754002};

// exp(x) for x <= 0 to ~1.5 ulp:  x = (32 e + j) ln2/32 + r, |r| <= ln2/64;  exp(x) = 2^e 2^(j/32) (1 + expm1(r)) with a
// degree-6 expm1 and the 32-entry table above -- 11 FP64 instructions instead of the ~23 DFMA-equivalents of exp()
// (profiles/r01_probe_fp64.txt; the kernel is bound by the FP64 pipe).  x is clamped at -708 (the smallest normal results):
// below that the reference's numpy exp returns denormals or 0, this returns 3e-308 -- both below 1e-300.
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab) {
    const double magic = 6755399441055744.0;                // 1.5 * 2^52
    // x = max(x, -708) on the high word (x <= 0: a more negative double has the larger unsigned high word; -708.0 is
    // 0xC086200000000000): one integer instruction instead of a DSETP and two selects.  NaN (high word above 0xFFF0...) is
    // clamped too -- the caller has flagged it (inf_flag) and the result is discarded.
    x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC0862000u), __double2loint(x));
    const double t = fma(x, 46.16624130844683 /* 32 / ln2 */, magic);
    const int ni = __double2loint(t);                       // round(32 x / ln2), |ni| < 2^16
    const double nd = t - magic;
    double r = fma(nd, -0.02166084938653512, x);            // ln2/32, upper 32 bits: nd * hi is exact
    r = fma(nd, -5.9631716539705866e-12, r);                // ln2/32 - hi
    double q = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    q = fma(r, q, 1.0 / 24.0);
    q = fma(r, q, 1.0 / 6.0);
    q = fma(r, q, 0.5);
    const double pm1 = fma(r * r, q, r);                    // expm1(r)
    const double tj = tab[ni & 31];
    const double v = fma(tj, pm1, tj);                      // in [1, 2)
    const int e = ni >> 5;
    return __hiloint2double(__double2hiint(v) + (e << 20), __double2loint(v));
}

template <int KT>
__device__ __forceinline__ double kfun(double r2, const double* __restrict__ tab) {
    if (KT == MOGP_KERNEL_SQEXP) {
        return exp_neg(-0.5 * r2, tab);
    } else {
        const double s = sqrt(5.0 * r2);
        return (1.0 + s + (5.0 / 3.0) * r2) * exp_neg(-s, tab);
    }
}

struct KmatParams {
    int64_t n, n_pad;       // training points (real / padded)
    int64_t rows_pad;       // CROSS: padded number of test points (m_pad)
    int d, dbox;            // input dims, dims per TMA box
    int hyper_stride;       // doubles per output in hyper (d + 2)
    const double* hyper;    // [count][d+2]
    double* out;            // SYM: matrix slab base; CROSS: workspace slab base
    int64_t out_stride;     // CROSS: rows per output in the workspace slab
    const double* alpha;    // CROSS: [count][n_pad] or null
    int64_t alpha_stride;
    double* part;           // CROSS: [count][n_tiles][rows_pad] or null
    int outs[MAXG];         // global output index handled by local output k (hyper/alpha rows)
    int store;              // CROSS: write the matrix (0 when only the mean is wanted)
    int add_nugget;         // SYM: add hyper[d+1] (the nugget) on the diagonal
    int* inf_flag;          // set to 1 when a squared distance is +inf (Kernel.py:482-483 raises FloatingPointError); may be null
    int tiles_i, tiles_j;   // CROSS: tiles of test / training points; SYM: tiles_i = lower tiles per output
    int count;
};

// Persistent CTAs, TWO per SM, walk the 64 x 128 tiles (output, I, J); the point tiles of X of the NEXT unit (tile, chunk of
// 16 dims) are TMA-loaded into the other stage of a two-stage ring while the current one is computed.  256 threads, each
// owning a 4 x 8 register sub-tile: rows ty + 16 a, columns 2 tx + 32 b' + {0, 1} (pairs of adjacent columns: 16-byte loads
// of the column operand and 16-byte stores of K).  Two co-resident CTAs because a tile ends with a 64 KB store burst: with
// one CTA per SM (every warp behind the same barriers) the FP64 pipe idled while the stores drained and the kernel ran at
// the SUM of its FP64 time and its HBM-write time; two CTAs drift out of phase and overlap them.  The loaded tiles are scaled by sqrt(exp(theta_d)) in shared memory first, so the
// inner loop is one subtraction and one FMA per dimension and element: r2 += (x~ - x~')^2.
template <int KT, int CROSS>
__global__ void __launch_bounds__(KM_THREADS, 2)
kmat_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmC, const KmatParams p) {
    extern __shared__ __align__(128) unsigned char km_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(km_smem_raw) + 127) & ~uintptr_t(127));
    __shared__ double sw_s[256];     // sqrt of the weights of the current output (up to 256 dims)
    __shared__ double al_s[128];
    __shared__ double tab_s[32];
    __shared__ __align__(8) uint64_t full[2];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    if (tid < 32) tab_s[tid] = kExp2Table[tid];
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int nchunk = (p.d + p.dbox - 1) / p.dbox;
    const int tiles_per_out = CROSS ? p.tiles_i * p.tiles_j : p.tiles_i;
    const int n_tiles = p.count * tiles_per_out;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int units = my_tiles * nchunk;

    auto decode = [&](int u, int& o, int& I, int& J, int& ch) {
        const int tl = blockIdx.x + (u / nchunk) * gridDim.x;
        ch = u % nchunk;
        o = tl / tiles_per_out;
        const int id = tl - o * tiles_per_out;
        if (CROSS) {
            I = id / p.tiles_j;      // tile of 64 test points
            J = id - I * p.tiles_j;  // tile of 128 training points
        } else {
            // lower tiles: the row tiles 2k and 2k+1 (64 rows each) of block row k have k + 1 column tiles each
            int k = (int)((sqrtf(4.0f * (float)id + 1.0f) - 1.0f) * 0.5f);
            while ((k + 1) * (k + 2) <= id) k++;
            while (k * (k + 1) > id) k--;
            const int rem = id - k * (k + 1);
            I = 2 * k + (rem > k ? 1 : 0);
            J = rem > k ? rem - (k + 1) : rem;
        }
    };
    auto issue = [&](int u) {
        int o, I, J, ch;
        decode(u, o, I, J, ch);
        unsigned char* st = base + (u & 1) * KM_STAGE;
        mbar_arrive_expect_tx(&full[u & 1], 2u * 128u * (uint32_t)p.dbox * 8u);
        tma_load_2d(st, &tmR, I * KM_ROWS, ch * p.dbox, &full[u & 1]);      // (a box of 128 points: the upper 64 are not used)
        tma_load_2d(st + KM_TILE, &tmC, J * 128, ch * p.dbox, &full[u & 1]);
    };

    if (tid == 0 && units > 0) issue(0);
    double r2[KM_RA][8];
    int cur_o = -1;
    for (int u = 0; u < units; u++) {
        int o, I, J, ch;
        decode(u, o, I, J, ch);
        const double* hyp = p.hyper + (int64_t)p.outs[o] * p.hyper_stride;
        if (o != cur_o) {            // (the previous unit ended with a barrier: nobody reads sw_s any more)
            for (int i = tid; i < p.d; i += KM_THREADS) sw_s[i] = sqrt(hyp[i]);
            cur_o = o;
        }
        if (CROSS && ch == 0 && tid < 128)
            al_s[tid] = p.alpha ? p.alpha[(int64_t)p.outs[o] * p.alpha_stride + (int64_t)J * 128 + tid] : 0.0;
        if (tid == 0 && u + 1 < units) issue(u + 1);      // the other stage was released by the barrier that ended unit u - 1
        if (ch == 0) {
#pragma unroll
            for (int a = 0; a < KM_RA; a++)
#pragma unroll
                for (int b = 0; b < 8; b++) r2[a][b] = 0.0;
        }
        mbar_wait(&full[u & 1], (uint32_t)((u >> 1) & 1));
        __syncthreads();             // sw_s / al_s written
        double* Xr = reinterpret_cast<double*>(base + (u & 1) * KM_STAGE);
        double* Xc = Xr + DCH * 128;
        const int dlim = min(p.dbox, p.d - ch * p.dbox);
        // scale both tiles by sqrt(w_d) in place: 2 * dlim * 128 elements
        for (int idx = tid; idx < dlim * 256; idx += KM_THREADS) {
            const int dd = idx >> 8, e = idx & 255;
            double* q = (e < 128) ? (Xr + dd * 128 + e) : (Xc + dd * 128 + e - 128);
            *q *= sw_s[ch * p.dbox + dd];
        }
        __syncthreads();
        for (int dd = 0; dd < dlim; dd++) {
            double xr[KM_RA];
            double2 xc[4];
#pragma unroll
            for (int a = 0; a < KM_RA; a++) xr[a] = Xr[dd * 128 + ty + (KM_THREADS / 16) * a];
#pragma unroll
            for (int b = 0; b < 4; b++) xc[b] = *reinterpret_cast<const double2*>(Xc + dd * 128 + 2 * tx + 32 * b);
#pragma unroll
            for (int a = 0; a < KM_RA; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double d0 = xr[a] - xc[b].x, d1 = xr[a] - xc[b].y;
                    r2[a][2 * b] = fma(d0, d0, r2[a][2 * b]);
                    r2[a][2 * b + 1] = fma(d1, d1, r2[a][2 * b + 1]);
                }
        }
        if (ch == nchunk - 1) {
            if (p.inf_flag) {
                // the reference refuses infinite distances (calc_r2, Kernel.py:482-483); integer test of the bit pattern: the
                // kernel is bound by the FP64 pipe
                // r2 >= 0 (or NaN): the largest high word of the sub-tile tells whether any entry is inf / NaN; the exact test runs
                // only then (one integer max per entry instead of two compares and a predicate)
                int hmax = 0;
#pragma unroll
                for (int a = 0; a < KM_RA; a++)
#pragma unroll
                    for (int b = 0; b < 8; b++) hmax = max(hmax, __double2hiint(r2[a][b]) & 0x7fffffff);
                if (hmax >= 0x7ff00000) {
                    bool inf = false;
#pragma unroll
                    for (int a = 0; a < KM_RA; a++)
#pragma unroll
                        for (int b = 0; b < 8; b++)
                            inf |= (__double2hiint(r2[a][b]) == 0x7ff00000) && (__double2loint(r2[a][b]) == 0);
                    if (inf) atomicOr(p.inf_flag, 1);
                }
            }
            const double sigma2 = hyp[p.d];
            if (!CROSS) {
                const double nugget = p.add_nugget ? hyp[p.d + 1] : 0.0;
                const int64_t row_base = (int64_t)p.outs[o] * p.out_stride;
#pragma unroll
                for (int a = 0; a < KM_RA; a++) {
                    const int64_t row = (int64_t)I * KM_ROWS + ty + (KM_THREADS / 16) * a;
                    double* orow = p.out + (row_base + row) * p.n_pad + (int64_t)J * 128;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        double v[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int64_t col = (int64_t)J * 128 + 2 * tx + 32 * b + e;
                            double w = sigma2 * kfun<KT>(r2[a][2 * b + e], tab_s);
                            if (row == col) w += nugget;
                            if (row >= p.n || col >= p.n) w = (row == col) ? 1.0 : 0.0;
                            v[e] = w;
                        }
                        *reinterpret_cast<double2*>(orow + 2 * tx + 32 * b) = make_double2(v[0], v[1]);
                    }
                }
            } else {
                const int nvalid = (int)min((int64_t)128, p.n - (int64_t)J * 128);     // training points of this tile
                const bool ragged = nvalid < 128;                                      // (CTA-uniform: only the last tile of a ragged n)
#pragma unroll
                for (int a = 0; a < KM_RA; a++) {
                    const int64_t row = (int64_t)I * KM_ROWS + ty + (KM_THREADS / 16) * a;  // test point
                    double* orow = p.out + ((int64_t)o * p.out_stride + row) * p.n_pad + (int64_t)J * 128;
                    double sdot = 0.0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        double v[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int c = 2 * tx + 32 * b + e;           // training point inside the tile
                            double w = sigma2 * kfun<KT>(r2[a][2 * b + e], tab_s);
                            if (ragged && c >= nvalid) w = 0.0;
                            sdot = fma(w, al_s[c], sdot);
                            v[e] = w;
                        }
                        if (p.store) *reinterpret_cast<double2*>(orow + 2 * tx + 32 * b) = make_double2(v[0], v[1]);
                    }
                    if (p.part) {
                        sdot += __shfl_xor_sync(0xffffffffu, sdot, 8);
                        sdot += __shfl_xor_sync(0xffffffffu, sdot, 4);
                        sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
                        sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
                        if (tx == 0) p.part[((int64_t)o * p.tiles_j + J) * p.rows_pad + row] = sdot;
                    }
                }
            }
        }
        fence_proxy_async();         // the tiles were scaled in place through the generic proxy; a TMA load overwrites them next
        __syncthreads();             // this stage (and sw_s / al_s) may be overwritten
    }
}

struct OutList {
    int outs[MAXG];
};

__global__ void mean_reduce_kernel(const double* __restrict__ part, int n_tiles, int64_t m_pad, int64_t m,
                                   double* __restrict__ mean, int64_t mean_stride, const OutList ol) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int o = blockIdx.y;
    if (c >= m) return;
    double s = 0.0;
    for (int J = 0; J < n_tiles; J++) s += part[((int64_t)o * n_tiles + J) * m_pad + c];
    mean[(int64_t)ol.outs[o] * mean_stride + c] = s;
}



int kmat_dbox(int d) { return d < DCH ? d : DCH; }

// SMs of the current device (persistent grids); cached per device
static int kmat_sms() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = cache[dev & 63];
    if (c == 0 && cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) c = 148;
    return c;
}

// K + nugget*I (lower 128x128 tiles) of `count` outputs in one launch.  Output k uses hyper row outs[k] and is
// written at slab rows slab_idx[k]*n_pad.. of A_slab (slab_idx == nullptr: same as outs).
int kmat_sym(const CUtensorMap& tmXT, int kernel, int64_t n, int64_t n_pad, int d, const double* hyper, const int* outs,
             int count, int add_nugget, double* A_slab, int64_t slab_rows_per_output, cudaStream_t st, int* inf_flag) {
    if (count < 1 || count > MAXG) return 1;
    KmatParams p{};
    p.n = n; p.n_pad = n_pad; p.rows_pad = n_pad; p.d = d; p.dbox = kmat_dbox(d); p.hyper_stride = d + 2;
    p.hyper = hyper; p.out = A_slab; p.out_stride = slab_rows_per_output; p.add_nugget = add_nugget; p.inf_flag = inf_flag;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    const int T = (int)(n_pad / 128);
    p.tiles_i = T * (T + 1); p.tiles_j = 1; p.count = count;       // lower 64 x 128 tiles per output
    const int64_t tiles = (int64_t)p.tiles_i * count;
    const unsigned grid = (unsigned)(tiles < 2 * kmat_sms() ? tiles : 2 * kmat_sms());
    if (kernel == MOGP_KERNEL_SQEXP)
        kmat_kernel<MOGP_KERNEL_SQEXP, 0><<<grid, KM_THREADS, KM_SMEM, st>>>(tmXT, tmXT, p);
    else
        kmat_kernel<MOGP_KERNEL_MATERN52, 0><<<grid, KM_THREADS, KM_SMEM, st>>>(tmXT, tmXT, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int kmat_cross(const CUtensorMap& tmXsT, const CUtensorMap& tmXT, int kernel, int64_t n, int64_t n_pad, int64_t m_pad,
               int d, const int* outs, int count, const double* hyper, double* W_slab, int64_t w_stride, int store,
               const double* alpha, int64_t alpha_stride, double* part, cudaStream_t st, int* inf_flag) {
    KmatParams p{};
    p.inf_flag = inf_flag;
    p.n = n; p.n_pad = n_pad; p.rows_pad = m_pad; p.d = d; p.dbox = kmat_dbox(d); p.hyper_stride = d + 2;
    p.hyper = hyper; p.out = W_slab; p.out_stride = w_stride; p.alpha = alpha; p.alpha_stride = alpha_stride;
    p.part = part; p.store = store;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    p.tiles_i = (int)(m_pad / KM_ROWS); p.tiles_j = (int)(n_pad / 128); p.count = count;
    const int64_t tiles = (int64_t)p.tiles_i * p.tiles_j * count;
    const unsigned grid = (unsigned)(tiles < 2 * kmat_sms() ? tiles : 2 * kmat_sms());
    if (kernel == MOGP_KERNEL_SQEXP)
        kmat_kernel<MOGP_KERNEL_SQEXP, 1><<<grid, KM_THREADS, KM_SMEM, st>>>(tmXsT, tmXT, p);
    else
        kmat_kernel<MOGP_KERNEL_MATERN52, 1><<<grid, KM_THREADS, KM_SMEM, st>>>(tmXsT, tmXT, p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int mean_reduce(const double* part, const int* outs, int count, int n_tiles, int64_t m_pad, int64_t m, double* mean,
                int64_t mean_stride, cudaStream_t st) {
    dim3 grid((unsigned)((m + 255) / 256), (unsigned)count);
    OutList ol{};
    for (int i = 0; i < count; i++) ol.outs[i] = outs[i];
    mean_reduce_kernel<<<grid, 256, 0, st>>>(part, n_tiles, m_pad, m, mean, mean_stride, ol);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ------------------------------------------------------------------------------------------
// Predictive derivatives d mean / d x*  (replaces cov_deriv_x_batch_gpu + cublasDgemv, reference
// mogp_gpu/src/kernel.cu:69-100, 264-302 and densegp_gpu.hpp:411-448):
//     dmu_c/dx*_q = sum_i alpha_i * sigma2 * k'(r2_ci) * 2 w_q (x*_cq - x_iq),   k' = dk/dr2
// One thread per test point streams the training points (tiles of 32 staged in shared memory, broadcast reads)
// and keeps 16 derivative components in registers (its own coordinates come through L1); more than 16 input
// dimensions take further passes.  No
// (m x n x d) derivative tensor is materialised (the reference allocates exactly that in work_mat_d).
// ------------------------------------------------------------------------------------------
constexpr int KD_TB = 32;    // training points per staged tile
constexpr int KD_Q = 16;     // derivative components per pass
constexpr int KD_MAXD = 256;   // == the library's limit on input dimensions

struct DerivParams {
    const double* XsT;     // [d][xs_stride] test points, transposed
    int64_t xs_stride;
    const double* XT;      // [d][n_pad]
    const double* alpha;   // [E][n_pad]
    const double* hyper;   // [E][d+2]
    double* out;           // [count][m][d]
    int64_t n, n_pad, m;
    int d;
    int outs[MAXG];
};

template <int KT>
__global__ void __launch_bounds__(128) kderiv_kernel(const DerivParams p) {
    extern __shared__ __align__(16) double kd_smem[];
    const int d = p.d;
    double* xt = kd_smem;              // [d][KD_TB]
    double* al = xt + d * KD_TB;       // [KD_TB]
    double* w = al + KD_TB;            // [d]
    const int tid = threadIdx.x;
    const int o = p.outs[blockIdx.y];
    const int64_t c = (int64_t)blockIdx.x * 128 + tid;
    const double* hyp = p.hyper + (int64_t)o * (d + 2);
    for (int i = tid; i < d; i += 128) w[i] = hyp[i];
    // this thread's test point: coalesced, L1-resident reads of XsT (zero-padded up to xs_stride)
    const double* xs = p.XsT + ((c < p.xs_stride) ? c : 0);
    const int64_t xss = p.xs_stride;
    const double sigma2 = hyp[d];
    __syncthreads();
    for (int d0 = 0; d0 < d; d0 += KD_Q) {
        double g[KD_Q];
#pragma unroll
        for (int q = 0; q < KD_Q; q++) g[q] = 0.0;
        for (int64_t i0 = 0; i0 < p.n; i0 += KD_TB) {
            __syncthreads();
            for (int idx = tid; idx < d * KD_TB; idx += 128) {
                const int dd = idx / KD_TB, ii = idx - dd * KD_TB;
                xt[idx] = (i0 + ii < p.n) ? p.XT[(int64_t)dd * p.n_pad + i0 + ii] : 0.0;
            }
            if (tid < KD_TB) al[tid] = (i0 + tid < p.n) ? p.alpha[(int64_t)o * p.n_pad + i0 + tid] : 0.0;
            __syncthreads();
#pragma unroll 2
            for (int ii = 0; ii < KD_TB; ii++) {
                double r2 = 0.0;
                for (int dd = 0; dd < d; dd++) {
                    const double df = __ldg(xs + dd * xss) - xt[dd * KD_TB + ii];
                    r2 = fma(w[dd], df * df, r2);
                }
                double dk;
                if (KT == MOGP_KERNEL_SQEXP) {
                    dk = -0.5 * exp(-0.5 * r2);
                } else {
                    const double sq = sqrt(5.0 * r2);
                    dk = -(5.0 / 6.0) * (1.0 + sq) * exp(-sq);
                }
                const double kp = sigma2 * dk * al[ii];   // al is zero past n
#pragma unroll
                for (int q = 0; q < KD_Q; q++)
                    if (d0 + q < d) g[q] = fma(kp, __ldg(xs + (d0 + q) * xss) - xt[(d0 + q) * KD_TB + ii], g[q]);
            }
        }
        if (c < p.m) {
            double* orow = p.out + ((int64_t)blockIdx.y * p.m + c) * d;
#pragma unroll
            for (int q = 0; q < KD_Q; q++)
                if (d0 + q < d) orow[d0 + q] = 2.0 * w[d0 + q] * g[q];
        }
    }
}

int kderiv_max_dims() { return KD_MAXD; }

int kmat_init() {
    if (cudaFuncSetAttribute(kmat_kernel<MOGP_KERNEL_SQEXP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(kmat_kernel<MOGP_KERNEL_SQEXP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(kmat_kernel<MOGP_KERNEL_MATERN52, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(kmat_kernel<MOGP_KERNEL_MATERN52, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, KM_SMEM) != cudaSuccess)
        return 1;
    const int maxs = (int)(sizeof(double) * ((size_t)KD_MAXD * KD_TB + KD_TB + KD_MAXD));
    if (cudaFuncSetAttribute(kderiv_kernel<MOGP_KERNEL_SQEXP>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs) != cudaSuccess ||
        cudaFuncSetAttribute(kderiv_kernel<MOGP_KERNEL_MATERN52>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs) != cudaSuccess)
        return 1;
    return 0;
}

int kmat_deriv(int kernel, const double* XsT, int64_t xs_stride, const double* XT, int64_t n, int64_t n_pad, int64_t m,
               int d, const int* outs, int count, const double* hyper, const double* alpha, double* out, cudaStream_t st) {
    if (count < 1 || count > MAXG || d > KD_MAXD) return 1;
    const size_t smem = sizeof(double) * ((size_t)d * KD_TB + KD_TB + d);
    DerivParams p{};
    p.XsT = XsT; p.xs_stride = xs_stride; p.XT = XT; p.alpha = alpha; p.hyper = hyper; p.out = out;
    p.n = n; p.n_pad = n_pad; p.m = m; p.d = d;
    for (int i = 0; i < count; i++) p.outs[i] = outs[i];
    dim3 grid((unsigned)((m + 127) / 128), (unsigned)count);
    if (kernel == MOGP_KERNEL_SQEXP)
        kderiv_kernel<MOGP_KERNEL_SQEXP><<<grid, 128, smem, st>>>(p);
    else
        kderiv_kernel<MOGP_KERNEL_MATERN52><<<grid, 128, smem, st>>>(p);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace mogp
