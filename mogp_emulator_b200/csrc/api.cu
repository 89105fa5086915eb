// C ABI of libmogp_b200 (see include/mogp_b200.h): handle management, the fit / predict orchestration
// (every phase one batched launch over the handle's outputs on one stream) and getters.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges are no-ops unless a profiler injects its library

#include "../../include/mogp_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "nccl_dyn.h"
#include "pool.h"

namespace mogp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_kblocked_tmap(CUtensorMap* tm, const double* base, int64_t rows, int64_t ld, int box_rows, int box_kslabs) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 1;
    cuuint64_t gdim[3] = {8, (cuuint64_t)rows, (cuuint64_t)(ld / 8)};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 8, 64};
    cuuint32_t box[3] = {8, (cuuint32_t)box_rows, (cuuint32_t)box_kslabs};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

int make_2d_tmap(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                 int box_cols) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return 1;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

static inline int64_t round_up(int64_t v, int64_t q) { return (v + q - 1) / q * q; }

// MOGP_TRACE=1: host-side phase timings of the API calls on stderr
static bool trace_on() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("MOGP_TRACE");
        on = (e && *e && *e != '0') ? 1 : 0;
    }
    return on == 1;
}
struct TraceClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const char* what;
    explicit TraceClock(const char* w) : what(w) {}
    void mark(const char* label) {
        if (!trace_on()) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[mogp trace] %s: %s +%.3f ms\n", what, label, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// NVTX range over an API call or one of its phases (the reference has no tracing of its own; nsys / ncu --nvtx attribute the
// kernels of a fit or predict to these names)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

}  // namespace mogp

using namespace mogp;

constexpr int MAXM = 32;  // mean-function vectors per output (grad_max_mean())
constexpr int I8_DEFAULT_PLANES = 7;   // default of MOGP_TRSM_I8 (see mogp_create)
// MOGP_CHOL_I8 unset: the Cholesky takes the tcgen05 path when outputs x (block rows)^2 of the launch reaches this.  The
// history products are ~2 x faster there (profiles/r02b_chol_i8_check.txt: 32 x n=4096 24.7 -> 11.9 ms, one n=16384 43.5 -> 23.0 ms),
// but a launch of a few small matrices is bound by the chain D(j) -> ROW(j+1, ., j) -> DIAG(j+1) -> D(j+1), whose tiles have the
// longer epilogue on that path (one n=4096 matrix: 2.64 ms FP64, 2.79 ms int8; four: 3.52 / 3.31 ms).  Work / chain ~ outputs x T^3 / T.
constexpr int64_t CHOL_I8_MIN_WORK = 4096;
enum { T_KMAT = 0, T_CHOL, T_SOLVE, T_KSTAR, T_TRSM, T_GRAD, T_NTRSM, T_NLAUNCH, T_FIT, T_PRED_HOST, T_PRED_D2H,
       T_I8_PREP, T_I8_CHECK, T_I8_ROWS, T_I8_NROWS, T_I8_NFALLBACK, T_CHOL_I8_N, T_CHOL_I8_RECHECKED, T_CHOL_I8_OVERTURNED, T_COUNT };

struct mogp_handle {
    int device = 0, n_sms = 148;
    int64_t n = 0, n_pad = 0;
    int d = 0, E = 0, kernel = 0, nug_type = 0;
    double nug_fixed = 0.0;
    cudaStream_t main = nullptr, side = nullptr;   // side: the FP64 accuracy check of the int8 predict path runs beside the integer kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_d = nullptr, ev_d2 = nullptr, ev_e = nullptr, ev_f = nullptr;
    // device slabs
    double *XT = nullptr, *Y = nullptr, *A = nullptr, *Dinv = nullptr, *alpha = nullptr, *z = nullptr;
    double *hyper = nullptr, *scal = nullptr;  // scal: [E][2] = logdet, quad
    int* info = nullptr;
    // pinned host mirrors
    double *h_hyper = nullptr, *h_scal = nullptr;
    int* h_info = nullptr;
    CUtensorMap tmXT;
    CholMaps maps;
    std::vector<char> fitted;
    // predict workspace (grown on demand)
    double *XsT = nullptr, *W = nullptr, *part = nullptr, *res = nullptr, *h_res = nullptr, *h_XsT = nullptr;
    double *sync = nullptr, *normacc = nullptr;   // TRSM ticket/flag words (used as int) and running column norms
    double* csync = nullptr;                       // Cholesky ticket/progress words (used as int)
    // int8 (tcgen05) predict TRSM: planes of L~ per output (valid until the output is fitted again), planes of V per call
    int8_t* Lq = nullptr;
    double* Vq = nullptr;                          // (bytes; typed double for grow())
    size_t Vq_cap = 0;
    std::vector<char> lq_valid;
    std::vector<char> i8_bad;                      // output failed the a-posteriori accuracy check since its last fit: FP64 path
    int use_i8 = 0;                                // planes per operand of the tcgen05 path (6 or 7), 0 = FP64 DMMA path only
    int i8_check = 1;                              // MOGP_I8_CHECK=0 skips the a-posteriori check (diagnostic: to time it)
    int chol_i8 = 2;                               // Cholesky history products on tcgen05: 0 never, 1 whenever possible, 2 auto (MOGP_CHOL_I8)
    // a-posteriori check of the int8 path: sampled K* rows, their FP64 variances, ticket words / norms of that solve, ratios
    double *chk_W = nullptr, *chk_var = nullptr, *chk_sync = nullptr, *chk_norm = nullptr, *chk_ratio = nullptr, *h_chk_ratio = nullptr;
    size_t chk_W_cap = 0, chk_var_cap = 0, chk_sync_cap = 0, chk_norm_cap = 0, chk_ratio_cap = 0, h_chk_ratio_cap = 0;
    size_t csync_cap = 0;
    // analytic mean function (set by the host front-end after a fit, cleared by every fit of that output)
    double* U = nullptr;                           // [E][MAXM][n_pad]: u_q with K^-1 H A^-1 H^T K^-1 = sum_q u_q u_q^T
    std::vector<int> n_u;                          // vectors stored per output
    double* aux = nullptr;                         // scratch for mogp_solve_list / mogp_kstar_dot
    size_t aux_cap = 0;
    size_t XsT_cap = 0, W_cap = 0, part_cap = 0, res_cap = 0, h_res_cap = 0, h_XsT_cap = 0, sync_cap = 0, normacc_cap = 0;
    // grad workspace
    double* G = nullptr;
    size_t G_cap = 0;
    double* h_grad = nullptr;
    size_t h_grad_cap = 0;
    double timings[T_COUNT] = {0};
};

struct mogp_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;
    double *sendbuf = nullptr, *recvbuf = nullptr, *h_recv = nullptr;
    size_t send_cap = 0, recv_cap = 0, h_cap = 0;
    double* dscalar = nullptr;
};

#define API_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);  \
            return (e__ == cudaErrorMemoryAllocation) ? MOGP_ERR_NOMEM : MOGP_ERR_CUDA;              \
        }                                                                                            \
    } while (0)

// Copy-out of the posteriors from the pinned staging buffer into the caller's arrays: rows[k] = {dst, src}, `bytes` each.  The
// destination is usually freshly allocated (numpy.empty): its first touch faults every page, which bounds a single thread at
// ~8 GB/s (0.65 ms for the 5 MB of a C3 predict, 5 % of an 8-GPU step); a few threads take the faults in parallel.
static void copy_rows(const std::vector<std::pair<double*, const double*>>& rows, size_t bytes) {
    const size_t total = rows.size() * bytes;
    const int nt = total >= (size_t)1 << 20 ? (int)std::min<size_t>(4, rows.size()) : 1;
    auto work = [&](int t) {
        for (size_t k = t; k < rows.size(); k += nt) memcpy(rows[k].first, rows[k].second, bytes);
    };
    if (nt == 1) {
        work(0);
        return;
    }
    // (no exception may cross the C ABI: a thread that cannot be started leaves its share to the caller's thread)
    std::vector<std::thread> th;
    std::vector<int> todo{0};
    for (int t = 1; t < nt; t++) {
        try {
            th.emplace_back(work, t);
        } catch (...) {
            todo.push_back(t);
        }
    }
    for (int t : todo) work(t);
    for (auto& x : th) x.join();
}

static int grow(double** p, size_t* cap, size_t bytes, int device) {
    // device == -1: pinned host
    if (*cap >= bytes) return MOGP_OK;
    if (*p) {
        pool_free(*p);
        *p = nullptr;
        *cap = 0;
    }
    *p = (double*)pool_alloc(bytes, device);
    if (!*p) {
        set_error("allocation of %zu bytes failed (device %d)", bytes, device);
        return MOGP_ERR_NOMEM;
    }
    *cap = bytes;
    return MOGP_OK;
}

extern "C" {

int mogp_version(int32_t* major, int32_t* minor) {
    if (major) *major = 0;
    if (minor) *minor = 1;
    return MOGP_OK;
}

const char* mogp_last_error(void) { return g_err; }

// cudaGetDeviceProperties costs 2-35 ms per call on this driver (it refreshes clocks / power state): ask the two
// attributes we need instead, once per device and process.
struct DevInfo {
    int major = -1, sms = 0;
};
static const DevInfo& device_info(int device) {
    static DevInfo cache[64];
    DevInfo& d = cache[device & 63];
    if (d.major < 0) {
        int major = 0, sms = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) {
            d.major = major;
            d.sms = sms;
        } else {
            cudaGetLastError();
            static DevInfo none;
            none.major = 0;
            return none;
        }
    }
    return d;
}

int mogp_device_count(int32_t* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    int usable = 0;
    for (int i = 0; i < c && i < 64; i++)
        if (device_info(i).major == 10) usable++;
    if (count) *count = usable;
    return MOGP_OK;
}

int mogp_destroy(mogp_handle* h) {
    if (!h) return MOGP_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->main) cudaStreamDestroy(h->main);
    if (h->side) cudaStreamDestroy(h->side);
    cudaEvent_t evs[9] = {h->ev_a, h->ev_b, h->ev_c, h->ev_d, h->ev_d2, h->ev_e, h->ev_f, h->ev_fork, h->ev_join};
    for (auto e : evs)
        if (e) cudaEventDestroy(e);
    void* bufs[] = {h->XT, h->Y, h->A, h->Dinv, h->alpha, h->z, h->hyper, h->scal, h->XsT, h->W, h->part, h->res, h->G,
                    h->sync, h->normacc, h->csync, h->U, h->aux, h->Lq, h->Vq, h->chk_W, h->chk_var, h->chk_sync, h->chk_norm, h->chk_ratio, h->h_chk_ratio, h->info, h->h_hyper, h->h_scal, h->h_res, h->h_XsT, h->h_info, h->h_grad};
    for (auto p : bufs) pool_free(p);
    cudaGetLastError();
    delete h;
    return MOGP_OK;
}

int mogp_create(const double* X, int64_t n, int32_t d, const double* Y, int32_t n_out, int32_t kernel,
                int32_t nugget_type, double nugget, int32_t device, int32_t n_streams, mogp_handle** out) {
    if (!X || !Y || !out || n < 1 || d < 1 || d > 256 || n_out < 1) {
        set_error("mogp_create: bad shape (n=%lld d=%d n_out=%d; need n>=1, 1<=d<=256, n_out>=1)", (long long)n, d, n_out);
        return MOGP_ERR_ARG;
    }
    if (kernel != MOGP_KERNEL_SQEXP && kernel != MOGP_KERNEL_MATERN52) {
        set_error("mogp_create: unknown kernel %d", kernel);
        return MOGP_ERR_ARG;
    }
    if (nugget_type < 0 || nugget_type > 2 || (nugget_type == MOGP_NUG_FIXED && !(nugget >= 0.0))) {
        set_error("mogp_create: bad nugget specification");
        return MOGP_ERR_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_error("mogp_create: CUDA device %d not available (%d visible)", device, ndev);
        return MOGP_ERR_CUDA;
    }
    TraceClock tc("mogp_create");
    API_CUDA(cudaSetDevice(device));
    tc.mark("cudaSetDevice");
    const DevInfo& prop = device_info(device);
    tc.mark("device attributes");
    if (prop.major != 10) {
        set_error("mogp_create: device %d has compute capability major %d; this library contains sm_100a code only", device,
                  prop.major);
        return MOGP_ERR_CUDA;
    }
    {
        // kernel attributes (opt-in shared memory, non-portable cluster size) are per device
        static bool inited[64] = {false};
        if (!inited[device & 63]) {
            if (chol_init() || solve_init() || kmat_init() || predict_init() || grad_init() || i8_init()) {
                set_error("kernel attribute setup failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MOGP_ERR_CUDA;
            }
            inited[device & 63] = true;
        }
    }
    tc.mark("kernel attribute setup");
    mogp_handle* h = new mogp_handle();
    h->device = device;
    h->n_sms = prop.sms;
    h->n = n;
    h->n_pad = round_up(n, NB);
    h->d = d;
    h->E = n_out;
    h->kernel = kernel;
    h->nug_type = nugget_type;
    h->nug_fixed = nugget;
    h->fitted.assign(n_out, 0);
    h->n_u.assign(n_out, 0);
    h->lq_valid.assign(n_out, 0);
    h->i8_bad.assign(n_out, 0);
    {
        const char* e = getenv("MOGP_I8_CHECK");
        h->i8_check = (e && e[0] == '0') ? 0 : 1;
    }
    {
        // MOGP_TRSM_I8 = 0: FP64 DMMA path only; 6 / 7: planes per operand of the int8 tcgen05 path
        const char* e = getenv("MOGP_TRSM_I8");
        h->use_i8 = I8_DEFAULT_PLANES;
        if (e && e[0] == '0') h->use_i8 = 0;
        else if (e && e[0] == '6') h->use_i8 = 6;
        else if (e && (e[0] == '7' || e[0] == '1')) h->use_i8 = 7;
    }
    {
        // MOGP_CHOL_I8 = 0: FP64 DMMA Cholesky only; 1: int8 tcgen05 history products whenever there is a history (n > 128);
        // default: when the launch is bound by its arithmetic rather than by the chain of diagonal tiles (enqueue_attempt)
        const char* e = getenv("MOGP_CHOL_I8");
        h->chol_i8 = (e && e[0] == '0') ? 0 : ((e && e[0] == '1') ? 1 : 2);
    }
    const int64_t np = h->n_pad;
    (void)n_streams;   // kept in the ABI: every phase is one batched launch on the handle's stream, there is nothing to tune
    int rc = MOGP_OK;
#define CREATE_CUDA(expr)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__);  \
            rc = (e__ == cudaErrorMemoryAllocation) ? MOGP_ERR_NOMEM : MOGP_ERR_CUDA;                \
            cudaGetLastError();                                                                      \
            mogp_destroy(h);                                                                         \
            return rc;                                                                               \
        }                                                                                            \
    } while (0)
    CREATE_CUDA(cudaStreamCreateWithFlags(&h->main, cudaStreamNonBlocking));
    {
        int prio_lo = 0, prio_hi = 0;
        CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CREATE_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_hi));
    }
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreate(&h->ev_a));
    CREATE_CUDA(cudaEventCreate(&h->ev_b));
    CREATE_CUDA(cudaEventCreate(&h->ev_c));
    CREATE_CUDA(cudaEventCreate(&h->ev_d));
    CREATE_CUDA(cudaEventCreate(&h->ev_d2));
    CREATE_CUDA(cudaEventCreate(&h->ev_e));
    CREATE_CUDA(cudaEventCreate(&h->ev_f));
    tc.mark("streams + events");
#define CREATE_ALLOC(ptr, type, bytes, dev)                                                          \
    do {                                                                                             \
        ptr = (type*)pool_alloc((bytes), (dev));                                                     \
        if (!ptr) {                                                                                  \
            set_error("mogp_create: allocation of %zu bytes failed", (size_t)(bytes));               \
            mogp_destroy(h);                                                                         \
            return MOGP_ERR_NOMEM;                                                                   \
        }                                                                                            \
    } while (0)
    CREATE_ALLOC(h->XT, double, sizeof(double) * d * np, device);
    CREATE_ALLOC(h->Y, double, sizeof(double) * n_out * np, device);
    CREATE_ALLOC(h->A, double, sizeof(double) * (size_t)n_out * np * np, device);
    CREATE_ALLOC(h->Dinv, double, sizeof(double) * (size_t)n_out * np * NB, device);
    CREATE_ALLOC(h->alpha, double, sizeof(double) * n_out * np, device);
    CREATE_ALLOC(h->z, double, sizeof(double) * n_out * np, device);
    CREATE_ALLOC(h->hyper, double, sizeof(double) * n_out * (d + 2), device);
    CREATE_ALLOC(h->scal, double, sizeof(double) * n_out * 2, device);
    CREATE_ALLOC(h->info, int, sizeof(int) * (n_out + 1), device);      // [n_out] = infinite-distance flag of the kernel-matrix kernels
    CREATE_ALLOC(h->h_hyper, double, sizeof(double) * n_out * (d + 2), -1);
    CREATE_ALLOC(h->h_scal, double, sizeof(double) * n_out * 2, -1);
    CREATE_ALLOC(h->h_info, int, sizeof(int) * (n_out + 1), -1);
#undef CREATE_ALLOC
    tc.mark("allocations");
    {
        // transposed, zero-padded design matrix and zero-padded targets
        std::vector<double> xt((size_t)d * np, 0.0), yp((size_t)n_out * np, 0.0);
        for (int64_t i = 0; i < n; i++)
            for (int k = 0; k < d; k++) xt[(size_t)k * np + i] = X[i * d + k];
        for (int o = 0; o < n_out; o++) memcpy(&yp[(size_t)o * np], Y + (size_t)o * n, sizeof(double) * n);
        CREATE_CUDA(cudaMemcpy(h->XT, xt.data(), sizeof(double) * xt.size(), cudaMemcpyHostToDevice));
        CREATE_CUDA(cudaMemcpy(h->Y, yp.data(), sizeof(double) * yp.size(), cudaMemcpyHostToDevice));
    }
    tc.mark("H2D of X and Y");
    if (make_2d_tmap(&h->tmXT, h->XT, d, np, np, kmat_dbox(d), 128) ||
        chol_make_maps(&h->maps, h->A, h->Dinv, (int64_t)n_out * np, np)) {
        set_error("cuTensorMapEncodeTiled failed");
        mogp_destroy(h);
        return MOGP_ERR_CUDA;
    }
#undef CREATE_CUDA
    *out = h;
    return MOGP_OK;
}

int mogp_reset(mogp_handle* h, int32_t idx) {
    if (!h || idx >= h->E) return MOGP_ERR_ARG;
    if (idx < 0) std::fill(h->fitted.begin(), h->fitted.end(), 0);
    else h->fitted[idx] = 0;
    return MOGP_OK;
}

int mogp_is_fit(mogp_handle* h, int32_t idx, int32_t* out) {
    if (!h || idx < 0 || idx >= h->E || !out) return MOGP_ERR_ARG;
    *out = h->fitted[idx];
    return MOGP_OK;
}

// enqueue kernel matrix + factorisation + solves for the outputs outs[0..count) on h->main (batched launches), with
// the nugget currently stored in h->hyper[o][d+1]; info / (logdet, quad) land in the pinned mirrors.
static int enqueue_attempt(mogp_handle* h, const int* outs, int count, bool allow_i8 = true) {
    const int64_t np = h->n_pad;
    const int T = (int)(np / NB);
    NvtxRange nvtx_range("mogp fit attempt: kmat + cholesky + solves");
    for (int g0 = 0; g0 < count; g0 += MAXG) {
        const int cnt = std::min(MAXG, count - g0);
        const int* og = outs + g0;
        for (int i = 0; i < cnt; i++) {
            API_CUDA(cudaMemsetAsync(h->info + og[i], 0, sizeof(int), h->main));
            API_CUDA(cudaMemsetAsync(h->scal + 2 * og[i], 0, 2 * sizeof(double), h->main));
        }
        API_CUDA(cudaMemsetAsync(h->info + h->E, 0, sizeof(int), h->main));
        API_CUDA(cudaEventRecord(h->ev_a, h->main));
        if (kmat_sym(h->tmXT, h->kernel, h->n, np, h->d, h->hyper, og, cnt, 1, h->A, np, h->main, h->info + h->E)) {
            set_error("kmat launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MOGP_ERR_CUDA;
        }
        API_CUDA(cudaEventRecord(h->ev_b, h->main));
        int rc;
        if ((rc = grow(&h->csync, &h->csync_cap, chol_sync_bytes(cnt, T), h->device))) return rc;
        // Many or large factorisations are bound by their O(n^3) history products: those run as exact integer GEMM on the
        // tcgen05 tensor cores (chol_i8_kernel, 8 signed 7-bit planes per operand), which also leaves the planes of L the predict
        // TRSM needs.  A few small ones are bound by the chain of diagonal tiles, which is FP64 either way: DMMA kernel.
        bool ci8 = allow_i8 && h->chol_i8 != 0 && T >= 2 && np <= 32768 && (h->chol_i8 == 1 || (int64_t)cnt * T * T >= CHOL_I8_MIN_WORK);
        if (ci8 && !h->Lq) {
            h->Lq = (int8_t*)pool_alloc(i8_lq_bytes(T, chol_i8_planes()) * (size_t)h->E, h->device);
            if (!h->Lq) {
                cudaGetLastError();
                ci8 = false;      // no room for the planes: FP64 path
            }
        }
        int nl;
        if (ci8) {
            std::vector<int> exps(cnt);
            for (int i = 0; i < cnt; i++) {
                const double* hy = h->h_hyper + (size_t)og[i] * (h->d + 2);
                exps[i] = i8_scale_exponent(hy[h->d], hy[h->d + 1]);
            }
            nl = chol_i8_factor_batch(h->maps, h->A, h->Dinv, og, exps.data(), cnt, np, h->Lq, (int64_t)i8_lq_bytes(T, chol_i8_planes()),
                                      h->info, h->scal, (int*)h->csync, h->n_sms, h->main);
        } else {
            nl = chol_factor_batch(h->maps, h->A, h->Dinv, og, cnt, np, h->info, h->scal, (int*)h->csync, h->n_sms, h->main);
        }
        if (nl < 0) {
            set_error("cholesky launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MOGP_ERR_CUDA;
        }
        API_CUDA(cudaEventRecord(h->ev_c, h->main));
        int rs = solve_alpha(h->A, np, h->Dinv, h->Y, h->z, h->alpha, h->scal, h->info, og, cnt, h->n_sms, h->main);
        if (rs) {
            set_error(rs == 2 ? "n too large for the cluster solver" : "solve launch failed");
            return rs == 2 ? MOGP_ERR_ARG : MOGP_ERR_CUDA;
        }
        API_CUDA(cudaEventRecord(h->ev_d, h->main));
        h->timings[T_NLAUNCH] += nl + 2;
        for (int i = 0; i < cnt; i++) {
            API_CUDA(cudaMemcpyAsync(h->h_info + og[i], h->info + og[i], sizeof(int), cudaMemcpyDeviceToHost, h->main));
            API_CUDA(cudaMemcpyAsync(h->h_scal + 2 * og[i], h->scal + 2 * og[i], 2 * sizeof(double), cudaMemcpyDeviceToHost,
                                     h->main));
        }
        API_CUDA(cudaMemcpyAsync(h->h_info + h->E, h->info + h->E, sizeof(int), cudaMemcpyDeviceToHost, h->main));
        API_CUDA(cudaStreamSynchronize(h->main));
        if (h->h_info[h->E]) {
            set_error("Inf enountered in kernel distance computation");      // (sic) the reference's message, Kernel.py:483
            return MOGP_ERR_FPE;
        }
        std::vector<int> recheck;
        if (ci8) {
            for (int i = 0; i < cnt; i++) {
                h->lq_valid[og[i]] = (h->h_info[og[i]] == 0) ? 1 : 0;   // the planes of L came with the factor
                if (h->h_info[og[i]] != 0) recheck.push_back(og[i]);
            }
            h->timings[T_CHOL_I8_N] += cnt;
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
        h->timings[T_KMAT] += ms;
        cudaEventElapsedTime(&ms, h->ev_b, h->ev_c);
        h->timings[T_CHOL] += ms;
        cudaEventElapsedTime(&ms, h->ev_c, h->ev_d);
        h->timings[T_SOLVE] += ms;
        cudaEventElapsedTime(&ms, h->ev_a, h->ev_d);
        h->timings[T_FIT] += ms;
        if (!recheck.empty()) {
            // A "not positive definite" verdict of the tcgen05 factorisation is never reported as such: the outputs concerned are
            // assembled and factorised again by the FP64 kernel, whose verdict (and LAPACK-style info) stands.  Failures are rare
            // and fail early, so this costs next to nothing; it makes every adaptive-nugget decision the FP64 kernel's, whatever
            // the integer path does to a pivot that sits within rounding of zero.
            int rc = enqueue_attempt(h, recheck.data(), (int)recheck.size(), false);
            if (rc) return rc;
            h->timings[T_CHOL_I8_RECHECKED] += (double)recheck.size();
            for (int o : recheck)
                if (h->h_info[o] == 0) h->timings[T_CHOL_I8_OVERTURNED] += 1.0;
        }
    }
    return MOGP_OK;
}

// H2D of the hyperparameter rows of the listed outputs (one copy over their index range)
static int upload_hyper(mogp_handle* h, const int* idx, int count) {
    int lo = idx[0], hi = idx[0];
    for (int i = 1; i < count; i++) {
        lo = std::min(lo, idx[i]);
        hi = std::max(hi, idx[i]);
    }
    const int hs = h->d + 2;
    API_CUDA(cudaMemcpyAsync(h->hyper + (size_t)lo * hs, h->h_hyper + (size_t)lo * hs, sizeof(double) * (hi - lo + 1) * hs,
                             cudaMemcpyHostToDevice, h->main));
    return MOGP_OK;
}

int mogp_fit_list(mogp_handle* h, const int32_t* idx, int32_t count, const double* thetas, int32_t n_params,
                  double* quad_out, double* logdet_out, double* nugget_out, int32_t* status_out) {
    if (!h || !thetas || !idx || count < 1) {
        set_error("mogp_fit: bad arguments");
        return MOGP_ERR_ARG;
    }
    std::vector<char> seen(h->E, 0);
    for (int i = 0; i < count; i++) {
        if (idx[i] < 0 || idx[i] >= h->E || seen[idx[i]]) {
            set_error("mogp_fit: output index %d out of range or repeated", idx[i]);
            return MOGP_ERR_ARG;
        }
        seen[idx[i]] = 1;
    }
    const int d = h->d;
    const int want = d + 1 + (h->nug_type == MOGP_NUG_FIT ? 1 : 0);
    if (n_params != want) {
        set_error("mogp_fit: theta has %d entries, expected %d", n_params, want);
        return MOGP_ERR_ARG;
    }
    API_CUDA(cudaSetDevice(h->device));
    NvtxRange nvtx_range("mogp_fit");
    const int hs = d + 2;
    std::vector<double> nug(count, 0.0);
    std::vector<int> todo(count);
    std::vector<int> pos(h->E, -1);   // output -> position in the call
    for (int i = 0; i < count; i++) {
        const int o = idx[i];
        const double* th = thetas + (size_t)i * n_params;
        double* hy = h->h_hyper + (size_t)o * hs;
        for (int k = 0; k < d; k++) {
            hy[k] = exp(th[k]);                               // CorrTransform: l = exp(-theta/2) <=> weight exp(theta)
            if (std::isinf(hy[k])) {
                // exp(theta) overflows: every squared distance between points that differ in this dimension is +inf and calc_r2
                // raises (Kernel.py:482-483).  Caught here because the device kernels work with sqrt(exp(theta)) (inf - inf = NaN).
                for (int q = 0; q < count; q++) h->fitted[idx[q]] = 0;
                set_error("Inf enountered in kernel distance computation");
                return MOGP_ERR_FPE;
            }
        }
        hy[d] = exp(th[d]);                                   // CovTransform
        if (h->nug_type == MOGP_NUG_FIT) nug[i] = exp(th[d + 1]);
        else if (h->nug_type == MOGP_NUG_FIXED) nug[i] = h->nug_fixed;
        else nug[i] = 0.0;
        hy[d + 1] = nug[i];
        h->fitted[o] = 0;
        h->n_u[o] = 0;
        h->lq_valid[o] = 0;
        h->i8_bad[o] = 0;
        todo[i] = o;
        pos[o] = i;
    }
    int rc = upload_hyper(h, todo.data(), count);
    if (rc) return rc;
    rc = enqueue_attempt(h, todo.data(), count);
    if (rc) return rc;
    // adaptive jitter retries (linalg/cholesky.py:264-279): jitter = mean(diag K) * 1e-6, x10 per failure, 5 tries.
    // diag of a stationary kernel matrix is sigma2 for every entry, so mean(diag K) == sigma2.
    std::vector<int> status(count, MOGP_OK);
    std::vector<int> failed;
    for (int i = 0; i < count; i++)
        if (h->h_info[idx[i]] != 0) {
            status[i] = MOGP_ERR_NOT_PD;
            if (h->nug_type == MOGP_NUG_ADAPTIVE) failed.push_back(idx[i]);
        }
    double scale = 1e-6;
    for (int t = 0; t < 5 && !failed.empty(); t++, scale *= 10.0) {
        std::vector<int> run;
        for (int o : failed) {
            const double jitter = h->h_hyper[(size_t)o * hs + d] * scale;
            if (!std::isfinite(jitter)) continue;
            h->h_hyper[(size_t)o * hs + d + 1] = jitter;
            run.push_back(o);
        }
        if (run.empty()) break;
        if ((rc = upload_hyper(h, run.data(), (int)run.size()))) return rc;
        rc = enqueue_attempt(h, run.data(), (int)run.size());
        if (rc) return rc;
        failed.clear();
        for (int o : run) {
            if (h->h_info[o] == 0) {
                status[pos[o]] = MOGP_OK;
                nug[pos[o]] = h->h_hyper[(size_t)o * hs + d + 1];
            } else {
                failed.push_back(o);
            }
        }
    }
    for (int i = 0; i < count; i++) {
        const int o = idx[i];
        if (status[i] == MOGP_OK) h->fitted[o] = 1;
        h->h_hyper[(size_t)o * hs + d + 1] = nug[i];   // the nugget actually used enters the predictive variance
        if (status_out) status_out[i] = status[i];
        if (nugget_out) nugget_out[i] = nug[i];
        if (logdet_out) logdet_out[i] = status[i] == MOGP_OK ? h->h_scal[2 * o] : std::numeric_limits<double>::quiet_NaN();
        if (quad_out) quad_out[i] = status[i] == MOGP_OK ? h->h_scal[2 * o + 1] : std::numeric_limits<double>::quiet_NaN();
    }
    if ((rc = upload_hyper(h, todo.data(), count))) return rc;
    API_CUDA(cudaStreamSynchronize(h->main));
    return MOGP_OK;
}

int mogp_fit(mogp_handle* h, int32_t first, int32_t count, const double* thetas, int32_t n_params,
             double* quad_out, double* logdet_out, double* nugget_out, int32_t* status_out) {
    if (!h || first < 0 || count < 1 || first + count > h->E) {
        set_error("mogp_fit: bad output range");
        return MOGP_ERR_ARG;
    }
    std::vector<int32_t> idx(count);
    for (int i = 0; i < count; i++) idx[i] = first + i;
    return mogp_fit_list(h, idx.data(), count, thetas, n_params, quad_out, logdet_out, nugget_out, status_out);
}

// Runs predict for all fitted outputs; results land in h->res as [E][2][m] (mean row, variance row),
// rows of unfitted outputs are NaN.  Enqueued on h->main, not synchronised.
static int predict_device(mogp_handle* h, const double* Xs, int64_t m, int want_var, int include_nugget) {
    NvtxRange nvtx_range("mogp predict: kstar + mean + trsm");
    const int64_t np = h->n_pad;
    const int d = h->d;
    std::vector<int> fit_idx;
    for (int o = 0; o < h->E; o++)
        if (h->fitted[o]) fit_idx.push_back(o);
    int rc;
    if ((rc = grow(&h->res, &h->res_cap, sizeof(double) * (size_t)h->E * 2 * m, h->device))) return rc;
    API_CUDA(cudaMemsetAsync(h->res, 0xFF, sizeof(double) * (size_t)h->E * 2 * m, h->main));  // all-ones == NaN
    if (fit_idx.empty() || m == 0) return MOGP_OK;

    // group / chunk sizes under a workspace budget.  cudaMemGetInfo is a slow driver query (it can take tens of ms):
    // when the workspace of the previous call already holds the whole job, reuse it without asking.
    const size_t need_all = (size_t)(round_up(m, 128) + 128) * np * 8 * fit_idx.size();
    size_t budget;
    if (want_var && fit_idx.size() <= (size_t)MAXG && (h->W_cap >= need_all || pool_has_block(need_all, h->device))) {
        budget = need_all;   // the previous call's workspace (this handle's, or a closed handle's in the cache) holds the whole job
    } else {
        size_t free_b = 0, total_b = 0;
        API_CUDA(cudaMemGetInfo(&free_b, &total_b));
        budget = (size_t)((free_b + h->W_cap + pool_cached_bytes()) * 0.6);
    }
    {
        // MOGP_WORKSPACE_MB caps the predict workspace (tests: forces the grouping of outputs / chunking of test points that a
        // full device would cause)
        const char* e = getenv("MOGP_WORKSPACE_MB");
        if (e && atof(e) > 0.0) budget = std::min(budget, (size_t)(atof(e) * 1048576.0));
    }
    int64_t mc_max = m;
    while ((size_t)(round_up(mc_max, 128) + 128) * np * 8 > budget && mc_max > 128) mc_max = (mc_max + 1) / 2;
    for (int64_t m0 = 0; m0 < m; m0 += mc_max) {
        const int64_t mc = std::min<int64_t>(mc_max, m - m0);
        size_t gcap = budget / ((size_t)(round_up(mc, 128) + 128) * np * 8);
        if (gcap < 1) gcap = 1;
        if (gcap > (size_t)MAXG) gcap = MAXG;
        if (gcap > fit_idx.size()) gcap = fit_idx.size();
        const int G = (int)gcap;
        for (size_t g0 = 0; g0 < fit_idx.size(); g0 += G) {
            const int cnt = (int)std::min<size_t>(G, fit_idx.size() - g0);
            const int* outs = fit_idx.data() + g0;
            TrsmPlan plan = predict_plan(mc, cnt, (int)np, h->n_sms);
            // panels * nw < mc + 128, so this stride covers every panel of every plan (and is what the
            // budget above assumed); the kernel-matrix kernel fills all w_stride rows (zero-padded test points)
            const int64_t w_stride = round_up(mc, 128) + 128;
            const int n_tiles = (int)(np / 128);
            if ((rc = grow(&h->XsT, &h->XsT_cap, sizeof(double) * d * w_stride, h->device))) return rc;
            if ((rc = grow(&h->h_XsT, &h->h_XsT_cap, sizeof(double) * d * w_stride, -1))) return rc;
            if (want_var && (rc = grow(&h->W, &h->W_cap, sizeof(double) * (size_t)cnt * w_stride * np, h->device))) return rc;
            if ((rc = grow(&h->part, &h->part_cap, sizeof(double) * (size_t)cnt * n_tiles * w_stride, h->device))) return rc;
            if (g0 == 0) {
                memset(h->h_XsT, 0, sizeof(double) * d * w_stride);
                for (int64_t i = 0; i < mc; i++)
                    for (int k = 0; k < d; k++) h->h_XsT[(size_t)k * w_stride + i] = Xs[(m0 + i) * d + k];
                API_CUDA(cudaMemcpyAsync(h->XsT, h->h_XsT, sizeof(double) * d * w_stride, cudaMemcpyHostToDevice, h->main));
            }
            CUtensorMap tmXsT, tmW;
            if (make_2d_tmap(&tmXsT, h->XsT, d, w_stride, w_stride, kmat_dbox(d), 128)) {
                set_error("tensor map (XsT) failed");
                return MOGP_ERR_CUDA;
            }
            API_CUDA(cudaMemsetAsync(h->info + h->E, 0, sizeof(int), h->main));
            API_CUDA(cudaEventRecord(h->ev_a, h->main));
            if (kmat_cross(tmXsT, h->tmXT, h->kernel, h->n, np, w_stride, d, outs, cnt, h->hyper, h->W, w_stride, want_var,
                           h->alpha, np, h->part, h->main, h->info + h->E) ||
                mean_reduce(h->part, outs, cnt, n_tiles, w_stride, mc, h->res + m0, 2 * m, h->main)) {
                set_error("kstar launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MOGP_ERR_CUDA;
            }
            API_CUDA(cudaEventRecord(h->ev_b, h->main));
            // many right-hand sides (at least one panel chain per SM): the int8 / tcgen05 path.  There is no a-priori accuracy
            // gate: worst-case bounds of the fixed-point error (n 2^(2 es - 7 S) amplified by ||L^-1||) are orders of magnitude
            // above what is observed, so every call is checked a posteriori instead -- I8_NCHECK test points per output are also
            // solved by the FP64 kernel and must agree to 1 % of the parity bar; an output that fails sends its group back to the
            // FP64 path now and until it is fitted again.
            int i8_prep_launches = 0;
            // (n_pad <= 65536: a column of n_pad products of 7-bit digits times S pairs must stay below 2^31 in the s32 accumulators)
            bool i8 = want_var && h->use_i8 && n_tiles >= 2 && np <= 65536 &&
                      (int64_t)cnt * ((mc + i8_panel_width() - 1) / i8_panel_width()) >= (int64_t)h->n_sms;
            for (int k = 0; k < cnt && i8; k++)
                if (h->i8_bad[outs[k]]) i8 = false;
            if (i8) {
                // planes of L (all outputs of the handle, allocated once) and of V (this call); if the device cannot hold them the
                // call stays on the FP64 path
                const int panels8 = (int)((mc + i8_panel_width() - 1) / i8_panel_width());
                if (!h->Lq) h->Lq = (int8_t*)pool_alloc(i8_lq_bytes(n_tiles, h->use_i8) * (size_t)h->E, h->device);
                if (!h->Lq || grow(&h->Vq, &h->Vq_cap, i8_vq_bytes(cnt, panels8, n_tiles, h->use_i8), h->device) != MOGP_OK) {
                    pool_free(h->Lq);
                    h->Lq = nullptr;
                    std::fill(h->lq_valid.begin(), h->lq_valid.end(), 0);
                    cudaGetLastError();
                    h->use_i8 = 0;
                    i8 = false;
                } else {
                    plan = TrsmPlan{i8_panel_width(), panels8};
                }
            }
            if (want_var) {
                if (make_kblocked_tmap(&tmW, h->W, (int64_t)cnt * w_stride, np, plan.nw)) {
                    set_error("tensor map (W) failed");
                    return MOGP_ERR_CUDA;
                }
                if ((rc = grow(&h->sync, &h->sync_cap, predict_sync_bytes(plan, cnt, n_tiles), h->device))) return rc;
                if ((rc = grow(&h->normacc, &h->normacc_cap, sizeof(double) * (size_t)cnt * w_stride, h->device))) return rc;
                if (i8) {
                    const int S8 = h->use_i8;
                    const size_t lq_stride = i8_lq_bytes(n_tiles, S8);
                    if ((rc = grow(&h->sync, &h->sync_cap, i8_sync_bytes(cnt, plan.panels), h->device))) return rc;
                    std::vector<int> exps(cnt), stale, stale_exp;
                    for (int k = 0; k < cnt; k++) {
                        const double* hy = h->h_hyper + (size_t)outs[k] * (d + 2);
                        exps[k] = i8_scale_exponent(hy[d], hy[d + 1]);
                        if (!h->lq_valid[outs[k]]) {
                            stale.push_back(outs[k]);
                            stale_exp.push_back(exps[k]);
                        }
                    }
                    const int npt = i8_check_points();
                    const TrsmPlan cplan{npt, 1};
                    if (h->i8_check) {
                        // FP64 reference variances of the sampled test points (a copy of their K* rows: the FP64 kernel solves in
                        // place) on the side stream (higher priority, enqueued first; the gather runs on the main stream before the
                        // fork).  When its CTAs reach the SMs first they take one SM per output while the persistent integer kernel
                        // starts on the others, whose remaining CTAs join the ticket queue as those SMs free up (0.5 - 2 ms concurrent:
                        // what happens when a slicing pass sits in front of the integer kernel, and at some output counts without
                        // one); when the persistent kernel fills all SMs first, the check runs behind it (0.4 - 0.8 ms serial).
                        CUtensorMap tmWc;
                        if ((rc = grow(&h->chk_W, &h->chk_W_cap, sizeof(double) * (size_t)cnt * npt * np, h->device))) return rc;
                        if ((rc = grow(&h->chk_var, &h->chk_var_cap, sizeof(double) * (size_t)h->E * npt, h->device))) return rc;
                        if ((rc = grow(&h->chk_sync, &h->chk_sync_cap, predict_sync_bytes(cplan, cnt, n_tiles), h->device))) return rc;
                        if ((rc = grow(&h->chk_norm, &h->chk_norm_cap, sizeof(double) * (size_t)cnt * npt, h->device))) return rc;
                        if ((rc = grow(&h->chk_ratio, &h->chk_ratio_cap, sizeof(double) * MAXG, h->device))) return rc;
                        if ((rc = grow(&h->h_chk_ratio, &h->h_chk_ratio_cap, sizeof(double) * MAXG, -1))) return rc;
                        if (make_kblocked_tmap(&tmWc, h->chk_W, (int64_t)cnt * npt, np, npt)) {
                            set_error("tensor map (check workspace) failed");
                            return MOGP_ERR_CUDA;
                        }
                        if (i8_check_gather(outs, cnt, h->W, w_stride, np, mc, h->chk_W, h->main)) {
                            set_error("i8 check launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                            return MOGP_ERR_CUDA;
                        }
                        API_CUDA(cudaEventRecord(h->ev_fork, h->main));
                        API_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
                        API_CUDA(cudaEventRecord(h->ev_e, h->side));
                        if (predict_trsm(cplan, outs, cnt, h->maps.a128, h->maps.d128, tmWc, h->chk_W, npt, h->hyper, d, include_nugget, np,
                                         npt, h->chk_var, npt, 0, (int*)h->chk_sync, h->chk_norm, h->n_sms, h->side, 0,
                                         want_var == 2 ? 1 : 0)) {
                            set_error("i8 check launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                            return MOGP_ERR_CUDA;
                        }
                        API_CUDA(cudaEventRecord(h->ev_f, h->side));
                        API_CUDA(cudaEventRecord(h->ev_join, h->side));
                    }
                    API_CUDA(cudaEventRecord(h->ev_d, h->main));
                    if (!stale.empty()) {
                        if (i8_slice_L(S8, h->A, np, stale.data(), stale_exp.data(), (int)stale.size(), h->Lq, (int64_t)lq_stride, h->main)) {
                            set_error("i8_slice_L launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                            return MOGP_ERR_CUDA;
                        }
                        for (int o : stale) h->lq_valid[o] = 1;
                        i8_prep_launches = 1;
                    }
                    API_CUDA(cudaEventRecord(h->ev_d2, h->main));
                    // the integer forward substitution with its FP64 epilogue (K* in W is only read)
                    if (i8_trsm(S8, outs, exps.data(), cnt, plan.panels, h->Lq, (int64_t)lq_stride, (int8_t*)h->Vq, h->maps.d128, tmW, h->W,
                                w_stride, h->hyper, d, include_nugget, want_var == 2 ? 1 : 0, np, mc, h->res + m + m0, 2 * m,
                                h->normacc, (int*)h->sync, h->n_sms, h->main)) {
                        set_error("i8 predict launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                        return MOGP_ERR_CUDA;
                    }
                    if (h->i8_check) {
                        API_CUDA(cudaStreamWaitEvent(h->main, h->ev_join, 0));
                        if (i8_check_compare(outs, cnt, mc, h->res + m + m0, 2 * m, h->chk_var, h->hyper, d, h->chk_ratio, h->main)) {
                            set_error("i8 check launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                            return MOGP_ERR_CUDA;
                        }
                        API_CUDA(cudaMemcpyAsync(h->h_chk_ratio, h->chk_ratio, sizeof(double) * cnt, cudaMemcpyDeviceToHost, h->main));
                    }
                } else if (predict_trsm(plan, outs, cnt, h->maps.a128, h->maps.d128, tmW, h->W, w_stride, h->hyper, d,
                                 include_nugget, np, mc, h->res + m + m0, 2 * m, 0, (int*)h->sync, h->normacc, h->n_sms,
                                 h->main, 0, want_var == 2 ? 1 : 0)) {
                    set_error("predict_trsm launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    return MOGP_ERR_CUDA;
                }
                h->timings[T_NTRSM] += 1;
            }
            API_CUDA(cudaEventRecord(h->ev_c, h->main));
            h->timings[T_NLAUNCH] += 2 + (want_var ? 1 : 0);
            // per-phase device times (the sync also protects the reused workspace and the pinned XsT buffer)
            API_CUDA(cudaMemcpyAsync(h->h_info + h->E, h->info + h->E, sizeof(int), cudaMemcpyDeviceToHost, h->main));
            API_CUDA(cudaStreamSynchronize(h->main));
            if (h->h_info[h->E]) {
                set_error("Inf enountered in kernel distance computation");
                return MOGP_ERR_FPE;
            }
            float ms1 = 0.f, ms2 = 0.f;
            cudaEventElapsedTime(&ms1, h->ev_a, h->ev_b);
            cudaEventElapsedTime(&ms2, h->ev_b, h->ev_c);
            h->timings[T_KSTAR] += ms1;
            h->timings[T_TRSM] += ms2;
            if (i8) {
                float a = 0.f, b = 0.f, c = 0.f;
                cudaEventElapsedTime(&a, h->ev_d, h->ev_d2);
                if (h->i8_check) cudaEventElapsedTime(&b, h->ev_e, h->ev_f);      // on the side stream, concurrent with the integer kernel
                cudaEventElapsedTime(&c, h->ev_d2, h->ev_c);
                h->timings[T_I8_PREP] += a;
                h->timings[T_I8_CHECK] += b;
                h->timings[T_I8_ROWS] += c;
                h->timings[T_I8_NROWS] += n_tiles;          // block rows solved by the (single) int8 launch
                h->timings[T_NLAUNCH] += i8_prep_launches + (h->i8_check ? 3 : 0);
                bool failed = false;
                for (int k = 0; k < cnt && h->i8_check; k++)
                    if (!(h->h_chk_ratio[k] <= 1.0)) {
                        h->i8_bad[outs[k]] = 1;
                        failed = true;
                    }
                if (failed) {
                    // the fixed-point solve missed the bar on a sampled test point: redo the group in FP64 (K* in W is intact)
                    if (trace_on()) fprintf(stderr, "[mogp trace] int8 predict path failed its accuracy check: group redone in FP64\n");
                    const TrsmPlan fplan = predict_plan(mc, cnt, (int)np, h->n_sms);
                    CUtensorMap tmWf;
                    if (make_kblocked_tmap(&tmWf, h->W, (int64_t)cnt * w_stride, np, fplan.nw)) {
                        set_error("tensor map (W) failed");
                        return MOGP_ERR_CUDA;
                    }
                    if ((rc = grow(&h->sync, &h->sync_cap, predict_sync_bytes(fplan, cnt, n_tiles), h->device))) return rc;
                    if (predict_trsm(fplan, outs, cnt, h->maps.a128, h->maps.d128, tmWf, h->W, w_stride, h->hyper, d, include_nugget, np,
                                     mc, h->res + m + m0, 2 * m, 0, (int*)h->sync, h->normacc, h->n_sms, h->main, 0,
                                     want_var == 2 ? 1 : 0)) {
                        set_error("predict_trsm launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                        return MOGP_ERR_CUDA;
                    }
                    API_CUDA(cudaStreamSynchronize(h->main));
                    h->timings[T_I8_NFALLBACK] += 1;
                    h->timings[T_NLAUNCH] += 1;
                }
            }
        }
    }
    return MOGP_OK;
}

int mogp_predict(mogp_handle* h, const double* Xs, int64_t m, int32_t want_var, int32_t include_nugget,
                 double* mean, double* var, int32_t* status) {
    if (!h || (!Xs && m > 0) || m < 0 || !mean || (want_var && !var)) {
        set_error("mogp_predict: bad arguments");
        return MOGP_ERR_ARG;
    }
    API_CUDA(cudaSetDevice(h->device));
    for (int o = 0; o < h->E; o++)
        if (status) status[o] = h->fitted[o] ? MOGP_OK : MOGP_ERR_NOT_FIT;
    if (m == 0) return MOGP_OK;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = predict_device(h, Xs, m, want_var, include_nugget);
    if (rc) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    if ((rc = grow(&h->h_res, &h->h_res_cap, sizeof(double) * (size_t)h->E * 2 * m, -1))) return rc;
    API_CUDA(cudaMemcpyAsync(h->h_res, h->res, sizeof(double) * (size_t)h->E * 2 * m, cudaMemcpyDeviceToHost, h->main));
    API_CUDA(cudaStreamSynchronize(h->main));
    {
        std::vector<std::pair<double*, const double*>> rows;
        for (int o = 0; o < h->E; o++) {
            rows.emplace_back(mean + (size_t)o * m, h->h_res + (size_t)o * 2 * m);
            if (want_var) rows.emplace_back(var + (size_t)o * m, h->h_res + (size_t)o * 2 * m + m);
        }
        copy_rows(rows, sizeof(double) * m);
    }
    const auto t2 = std::chrono::steady_clock::now();
    h->timings[T_PRED_HOST] += std::chrono::duration<double, std::milli>(t1 - t0).count();
    h->timings[T_PRED_D2H] += std::chrono::duration<double, std::milli>(t2 - t1).count();
    return MOGP_OK;
}

int mogp_predict_deriv(mogp_handle* h, const double* Xs, int64_t m, double* deriv, int32_t* status) {
    if (!h || (!Xs && m > 0) || m < 0 || !deriv) {
        set_error("mogp_predict_deriv: bad arguments");
        return MOGP_ERR_ARG;
    }
    const int d = h->d;
    if (d > kderiv_max_dims()) {
        set_error("mogp_predict_deriv: at most %d input dimensions are supported", kderiv_max_dims());
        return MOGP_ERR_ARG;
    }
    API_CUDA(cudaSetDevice(h->device));
    std::vector<int> fit_idx;
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    for (int o = 0; o < h->E; o++) {
        if (status) status[o] = h->fitted[o] ? MOGP_OK : MOGP_ERR_NOT_FIT;
        if (h->fitted[o]) fit_idx.push_back(o);
        else std::fill(deriv + (size_t)o * m * d, deriv + (size_t)(o + 1) * m * d, qnan);
    }
    if (m == 0 || fit_idx.empty()) return MOGP_OK;
    int rc;
    const int64_t xs_stride = round_up(m, 128);
    if ((rc = grow(&h->XsT, &h->XsT_cap, sizeof(double) * d * xs_stride, h->device))) return rc;
    if ((rc = grow(&h->h_XsT, &h->h_XsT_cap, sizeof(double) * d * xs_stride, -1))) return rc;
    memset(h->h_XsT, 0, sizeof(double) * d * xs_stride);
    for (int64_t i = 0; i < m; i++)
        for (int k = 0; k < d; k++) h->h_XsT[(size_t)k * xs_stride + i] = Xs[i * d + k];
    API_CUDA(cudaMemcpyAsync(h->XsT, h->h_XsT, sizeof(double) * d * xs_stride, cudaMemcpyHostToDevice, h->main));
    for (size_t g0 = 0; g0 < fit_idx.size(); g0 += MAXG) {
        const int cnt = (int)std::min<size_t>(MAXG, fit_idx.size() - g0);
        const size_t bytes = sizeof(double) * (size_t)cnt * m * d;
        if ((rc = grow(&h->W, &h->W_cap, bytes, h->device))) return rc;   // the predict workspace doubles as the result buffer
        if (kmat_deriv(h->kernel, h->XsT, xs_stride, h->XT, h->n, h->n_pad, m, d, fit_idx.data() + g0, cnt, h->hyper,
                       h->alpha, h->W, h->main)) {
            set_error("deriv launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MOGP_ERR_CUDA;
        }
        h->timings[T_NLAUNCH] += 1;
        API_CUDA(cudaStreamSynchronize(h->main));
        for (int k = 0; k < cnt; k++)
            API_CUDA(cudaMemcpy(deriv + (size_t)fit_idx[g0 + k] * m * d, h->W + (size_t)k * m * d, sizeof(double) * m * d,
                                cudaMemcpyDeviceToHost));
    }
    return MOGP_OK;
}

int mogp_predict_cov(mogp_handle* h, int32_t idx, const double* Xs, int64_t m, int32_t include_nugget, double* mean,
                     double* cov) {
    if (!h || idx < 0 || idx >= h->E || !Xs || m < 1 || !mean || !cov) {
        set_error("mogp_predict_cov: bad arguments");
        return MOGP_ERR_ARG;
    }
    if (!h->fitted[idx]) {
        set_error("mogp_predict_cov: output %d has not been fit", idx);
        return MOGP_ERR_NOT_FIT;
    }
    API_CUDA(cudaSetDevice(h->device));
    const int64_t np = h->n_pad;
    const int d = h->d;
    const int n_tiles = (int)(np / 128);
    const int64_t m_pad = round_up(m, 128);
    const int64_t w_stride = m_pad + 128;
    const int outs[1] = {idx};
    int rc;
    // workspace: V = L^-1 K* (test-major, w_stride x n_pad) in W, the m_pad x m_pad covariance in G
    if ((rc = grow(&h->W, &h->W_cap, sizeof(double) * (size_t)w_stride * np, h->device))) return rc;
    if ((rc = grow(&h->G, &h->G_cap, sizeof(double) * (size_t)m_pad * m_pad, h->device))) return rc;
    if ((rc = grow(&h->XsT, &h->XsT_cap, sizeof(double) * d * w_stride, h->device))) return rc;
    if ((rc = grow(&h->h_XsT, &h->h_XsT_cap, sizeof(double) * d * w_stride, -1))) return rc;
    if ((rc = grow(&h->part, &h->part_cap, sizeof(double) * (size_t)n_tiles * w_stride, h->device))) return rc;
    if ((rc = grow(&h->res, &h->res_cap, sizeof(double) * (size_t)h->E * 2 * m, h->device))) return rc;
    if ((rc = grow(&h->normacc, &h->normacc_cap, sizeof(double) * (size_t)w_stride, h->device))) return rc;
    TrsmPlan plan = predict_plan(m, 1, (int)np, h->n_sms);
    if ((rc = grow(&h->sync, &h->sync_cap, predict_sync_bytes(plan, 1, n_tiles), h->device))) return rc;
    memset(h->h_XsT, 0, sizeof(double) * d * w_stride);
    for (int64_t i = 0; i < m; i++)
        for (int k = 0; k < d; k++) h->h_XsT[(size_t)k * w_stride + i] = Xs[i * d + k];
    API_CUDA(cudaMemcpyAsync(h->XsT, h->h_XsT, sizeof(double) * d * w_stride, cudaMemcpyHostToDevice, h->main));
    CUtensorMap tmXsT, tmW, tmW128, tmW64;
    if (make_2d_tmap(&tmXsT, h->XsT, d, w_stride, w_stride, kmat_dbox(d), 128) ||
        make_kblocked_tmap(&tmW, h->W, w_stride, np, plan.nw) || make_kblocked_tmap(&tmW128, h->W, w_stride, np, 128) ||
        make_kblocked_tmap(&tmW64, h->W, w_stride, np, 64)) {
        set_error("tensor map (predict_cov) failed");
        return MOGP_ERR_CUDA;
    }
    double* res_row = h->res + (size_t)idx * 2 * m;   // [mean | variance] rows of this output
    if (kmat_cross(tmXsT, h->tmXT, h->kernel, h->n, np, w_stride, d, outs, 1, h->hyper, h->W, w_stride, 1, h->alpha, np,
                   h->part, h->main) ||
        mean_reduce(h->part, outs, 1, n_tiles, w_stride, m, h->res, 2 * m, h->main) ||
        predict_trsm(plan, outs, 1, h->maps.a128, h->maps.d128, tmW, h->W, w_stride, h->hyper, d, include_nugget, np, m,
                     h->res + m, 2 * m, 0, (int*)h->sync, h->normacc, h->n_sms, h->main, 1) ||
        // sigma2 * k(X*, X*) [+ nugget I] (lower tiles), then minus V^T V
        kmat_sym(tmXsT, h->kernel, m, m_pad, d, h->hyper, outs, 1, include_nugget ? 1 : 0, h->G, 0, h->main) ||
        cov_syrk_sub(tmW128, tmW64, 0, m_pad, np, h->G, h->main)) {
        set_error("predict_cov launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MOGP_ERR_CUDA;
    }
    h->timings[T_NLAUNCH] += 5;
    API_CUDA(cudaMemcpyAsync(mean, res_row, sizeof(double) * m, cudaMemcpyDeviceToHost, h->main));
    API_CUDA(cudaMemcpy2DAsync(cov, sizeof(double) * m, h->G, sizeof(double) * m_pad, sizeof(double) * m, m,
                               cudaMemcpyDeviceToHost, h->main));
    API_CUDA(cudaStreamSynchronize(h->main));
    for (int64_t i = 0; i < m; i++)
        for (int64_t j = i + 1; j < m; j++) cov[i * m + j] = cov[j * m + i];
    return MOGP_OK;
}

// ---------------------------------------------------------------------------------------------
// primitives for the analytic mean function (GaussianProcess.py:657-685, 887-920; linalg_utils.py:5-168): the host
// front-end does the n_mean x n_mean algebra, the device the solves and the kernel-matrix products.
// ---------------------------------------------------------------------------------------------
int mogp_solve_list(mogp_handle* h, const int32_t* idx, int32_t count, const double* rhs, double* out) {
    if (!h || !idx || count < 1 || !rhs || !out) return MOGP_ERR_ARG;
    for (int i = 0; i < count; i++)
        if (idx[i] < 0 || idx[i] >= h->E || !h->fitted[idx[i]]) {
            set_error("mogp_solve_list: output %d out of range or not fit", idx[i]);
            return idx[i] < 0 || idx[i] >= h->E ? MOGP_ERR_ARG : MOGP_ERR_NOT_FIT;
        }
    API_CUDA(cudaSetDevice(h->device));
    const int64_t np = h->n_pad, n = h->n;
    const size_t slab = (size_t)h->E * np;                 // rhs | z | solution, each [E][n_pad]; then [E][2] scalars
    int rc;
    if ((rc = grow(&h->aux, &h->aux_cap, sizeof(double) * (3 * slab + 2 * h->E), h->device))) return rc;
    double *Y2 = h->aux, *z2 = h->aux + slab, *x2 = h->aux + 2 * slab, *scal2 = h->aux + 3 * slab;
    for (int i = 0; i < count; i++) {
        API_CUDA(cudaMemsetAsync(Y2 + (size_t)idx[i] * np, 0, sizeof(double) * np, h->main));
        API_CUDA(cudaMemcpyAsync(Y2 + (size_t)idx[i] * np, rhs + (size_t)i * n, sizeof(double) * n, cudaMemcpyHostToDevice, h->main));
    }
    for (int g0 = 0; g0 < count; g0 += MAXG) {
        const int cnt = std::min(MAXG, count - g0);
        int rs = solve_alpha(h->A, np, h->Dinv, Y2, z2, x2, scal2, h->info, idx + g0, cnt, h->n_sms, h->main);
        if (rs) {
            set_error("solve launch failed");
            return MOGP_ERR_CUDA;
        }
        h->timings[T_NLAUNCH] += 1;
    }
    API_CUDA(cudaStreamSynchronize(h->main));
    for (int i = 0; i < count; i++)
        API_CUDA(cudaMemcpy(out + (size_t)i * n, x2 + (size_t)idx[i] * np, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return MOGP_OK;
}

int mogp_set_alpha_list(mogp_handle* h, const int32_t* idx, int32_t count, const double* alpha) {
    if (!h || !idx || count < 1 || !alpha) return MOGP_ERR_ARG;
    API_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < count; i++) {
        if (idx[i] < 0 || idx[i] >= h->E) return MOGP_ERR_ARG;
        API_CUDA(cudaMemcpy(h->alpha + (size_t)idx[i] * h->n_pad, alpha + (size_t)i * h->n, sizeof(double) * h->n,
                            cudaMemcpyHostToDevice));
    }
    return MOGP_OK;
}

int mogp_set_mean_vectors_list(mogp_handle* h, const int32_t* idx, int32_t count, int32_t n_vec, const double* U) {
    if (!h || !idx || count < 1 || n_vec < 0 || n_vec > MAXM || (n_vec > 0 && !U)) {
        set_error("mogp_set_mean_vectors: at most %d vectors per output", MAXM);
        return MOGP_ERR_ARG;
    }
    API_CUDA(cudaSetDevice(h->device));
    const int64_t np = h->n_pad;
    if (!h->U) {
        h->U = (double*)pool_alloc(sizeof(double) * (size_t)h->E * MAXM * np, h->device);
        if (!h->U) return MOGP_ERR_NOMEM;
        API_CUDA(cudaMemset(h->U, 0, sizeof(double) * (size_t)h->E * MAXM * np));
    }
    for (int i = 0; i < count; i++) {
        if (idx[i] < 0 || idx[i] >= h->E) return MOGP_ERR_ARG;
        for (int q = 0; q < n_vec; q++)
            API_CUDA(cudaMemcpy(h->U + ((size_t)idx[i] * MAXM + q) * np, U + ((size_t)i * n_vec + q) * h->n, sizeof(double) * h->n,
                                cudaMemcpyHostToDevice));
        h->n_u[idx[i]] = n_vec;
    }
    return MOGP_OK;
}

int mogp_kstar_dot(mogp_handle* h, const double* Xs, int64_t m, const double* vecs, int32_t n_vec, double* out) {
    if (!h || !Xs || m < 1 || !vecs || n_vec < 1 || !out) return MOGP_ERR_ARG;
    API_CUDA(cudaSetDevice(h->device));
    const int64_t np = h->n_pad, n = h->n;
    const int d = h->d, E = h->E;
    const int n_tiles = (int)(np / 128);
    const int64_t w_stride = round_up(m, 128);
    std::vector<int> fit_idx;
    for (int o = 0; o < E; o++)
        if (h->fitted[o]) fit_idx.push_back(o);
    const double qnan = std::numeric_limits<double>::quiet_NaN();
    std::fill(out, out + (size_t)E * n_vec * m, qnan);
    if (fit_idx.empty()) return MOGP_OK;
    int rc;
    // aux: vectors [E][n_vec][n_pad] | results [E][n_vec][m]
    const size_t vbytes = sizeof(double) * (size_t)E * n_vec * np, rbytes = sizeof(double) * (size_t)E * n_vec * m;
    if ((rc = grow(&h->aux, &h->aux_cap, vbytes + rbytes, h->device))) return rc;
    double* V = h->aux;
    double* R = h->aux + (size_t)E * n_vec * np;
    API_CUDA(cudaMemsetAsync(V, 0, vbytes, h->main));
    API_CUDA(cudaMemcpy2DAsync(V, sizeof(double) * np, vecs, sizeof(double) * n, sizeof(double) * n, (size_t)E * n_vec,
                               cudaMemcpyHostToDevice, h->main));
    if ((rc = grow(&h->XsT, &h->XsT_cap, sizeof(double) * d * w_stride, h->device))) return rc;
    if ((rc = grow(&h->h_XsT, &h->h_XsT_cap, sizeof(double) * d * w_stride, -1))) return rc;
    if ((rc = grow(&h->part, &h->part_cap, sizeof(double) * (size_t)std::min<size_t>(fit_idx.size(), MAXG) * n_tiles * w_stride,
                   h->device)))
        return rc;
    memset(h->h_XsT, 0, sizeof(double) * d * w_stride);
    for (int64_t i = 0; i < m; i++)
        for (int k = 0; k < d; k++) h->h_XsT[(size_t)k * w_stride + i] = Xs[i * d + k];
    API_CUDA(cudaMemcpyAsync(h->XsT, h->h_XsT, sizeof(double) * d * w_stride, cudaMemcpyHostToDevice, h->main));
    CUtensorMap tmXsT;
    if (make_2d_tmap(&tmXsT, h->XsT, d, w_stride, w_stride, kmat_dbox(d), 128)) {
        set_error("tensor map (XsT) failed");
        return MOGP_ERR_CUDA;
    }
    for (size_t g0 = 0; g0 < fit_idx.size(); g0 += MAXG) {
        const int cnt = (int)std::min<size_t>(MAXG, fit_idx.size() - g0);
        for (int q = 0; q < n_vec; q++) {
            // K*^T v_q without storing K*: the fused partial dots of the kernel-matrix kernel, vector q of every output
            if (kmat_cross(tmXsT, h->tmXT, h->kernel, n, np, w_stride, d, fit_idx.data() + g0, cnt, h->hyper, nullptr, w_stride, 0,
                           V + (size_t)q * np, (int64_t)n_vec * np, h->part, h->main) ||
                mean_reduce(h->part, fit_idx.data() + g0, cnt, n_tiles, w_stride, m, R + (size_t)q * m, (int64_t)n_vec * m, h->main)) {
                set_error("kstar_dot launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MOGP_ERR_CUDA;
            }
            h->timings[T_NLAUNCH] += 2;
        }
    }
    API_CUDA(cudaStreamSynchronize(h->main));
    for (int o : fit_idx)
        API_CUDA(cudaMemcpy(out + (size_t)o * n_vec * m, R + (size_t)o * n_vec * m, sizeof(double) * n_vec * m, cudaMemcpyDeviceToHost));
    return MOGP_OK;
}

int mogp_get(mogp_handle* h, int32_t idx, int32_t which, double* out) {
    if (!h || idx < 0 || idx >= h->E || !out) return MOGP_ERR_ARG;
    if (!h->fitted[idx]) {
        set_error("mogp_get: output %d has not been fit", idx);
        return MOGP_ERR_NOT_FIT;
    }
    API_CUDA(cudaSetDevice(h->device));
    const int64_t n = h->n, np = h->n_pad;
    if (which == MOGP_GET_ALPHA) {
        API_CUDA(cudaMemcpy(out, h->alpha + (size_t)idx * np, sizeof(double) * n, cudaMemcpyDeviceToHost));
        return MOGP_OK;
    }
    if (which == MOGP_GET_L) {
        API_CUDA(cudaMemcpy2D(out, sizeof(double) * n, h->A + (size_t)idx * np * np, sizeof(double) * np, sizeof(double) * n, n,
                              cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = i + 1; j < n; j++) out[i * n + j] = 0.0;
        return MOGP_OK;
    }
    if (which == MOGP_GET_K) {
        // recomputed (the factor overwrote it): sigma2*k(X,X) without the nugget (GaussianProcess.get_K_matrix)
        double* tmp = (double*)pool_alloc(sizeof(double) * np * np, h->device);
        if (!tmp) {
            set_error("get K: allocation failed");
            return MOGP_ERR_NOMEM;
        }
        const int one[1] = {idx};
        int rc = kmat_sym(h->tmXT, h->kernel, n, np, h->d, h->hyper, one, 1, 0, tmp, 0, h->main);
        if (rc == 0) {
            cudaError_t e = cudaMemcpy2DAsync(out, sizeof(double) * n, tmp, sizeof(double) * np, sizeof(double) * n, n,
                                              cudaMemcpyDeviceToHost, h->main);
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->main);
            rc = e == cudaSuccess ? 0 : 1;
        }
        pool_free(tmp);
        if (rc) {
            set_error("get K failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MOGP_ERR_CUDA;
        }
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = i + 1; j < n; j++) out[i * n + j] = out[j * n + i];
        return MOGP_OK;
    }
    if (which == MOGP_GET_KINV) {
        // K^-1 = (L^-1)^T L^-1 (DenseGP_GPU::get_invQ, densegp_gpu.hpp:629, the reference's stored explicit inverse): L^-1 by the
        // dataflow TRSM on an identity right-hand side (as for the gradient), then one DMMA SYRK.  Nothing in the library keeps
        // or uses an explicit inverse; this getter exists for the reference's interface.
        const int T = (int)(np / NB);
        int rc;
        if ((rc = grow(&h->G, &h->G_cap, sizeof(double) * 2 * (size_t)np * np, h->device))) return rc;
        TrsmPlan plan = predict_plan_square(np, h->n_sms);
        if ((rc = grow(&h->sync, &h->sync_cap, predict_sync_bytes(plan, 1, T), h->device))) return rc;
        double* Wt = h->G;
        double* C = h->G + (size_t)np * np;
        CUtensorMap tmW, tmW128, tmW64;
        if (make_kblocked_tmap(&tmW, Wt, np, np, plan.nw) || make_kblocked_tmap(&tmW128, Wt, np, np, 128) ||
            make_kblocked_tmap(&tmW64, Wt, np, np, 64)) {
            set_error("tensor map (inverse workspace) failed");
            return MOGP_ERR_CUDA;
        }
        const int one[1] = {idx};
        API_CUDA(cudaMemsetAsync(C, 0, sizeof(double) * (size_t)np * np, h->main));
        if (grad_set_identity(Wt, np, h->main) ||
            predict_trsm(plan, one, 1, h->maps.a128, h->maps.d128, tmW, Wt, np, h->hyper, h->d, 0, np, np, nullptr, 0, 1,
                         (int*)h->sync, nullptr, h->n_sms, h->main) ||
            cov_syrk_sub(tmW128, tmW64, 0, np, np, C, h->main)) {
            set_error("get K^-1 launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MOGP_ERR_CUDA;
        }
        API_CUDA(cudaMemcpy2DAsync(out, sizeof(double) * n, C, sizeof(double) * np, sizeof(double) * n, n, cudaMemcpyDeviceToHost,
                                   h->main));
        API_CUDA(cudaStreamSynchronize(h->main));
        h->timings[T_NLAUNCH] += 3;
        for (int64_t i = 0; i < n; i++)
            for (int64_t j = 0; j <= i; j++) {
                const double v = -out[i * n + j];           // the SYRK subtracts from a zero matrix (lower tiles)
                out[i * n + j] = v;
                out[j * n + i] = v;
            }
        return MOGP_OK;
    }
    set_error("mogp_get: selector %d not available", which);
    return MOGP_ERR_ARG;
}

int mogp_logpost_grad_list(mogp_handle* h, const int32_t* idx, int32_t count, double* grad, int32_t n_params) {
    if (!h || !idx || count < 1 || !grad) return MOGP_ERR_ARG;
    const int d = h->d;
    const int want = d + 1 + (h->nug_type == MOGP_NUG_FIT ? 1 : 0);
    if (n_params != want) {
        set_error("mogp_logpost_grad: expected %d parameters, got %d", want, n_params);
        return MOGP_ERR_ARG;
    }
    for (int i = 0; i < count; i++) {
        if (idx[i] < 0 || idx[i] >= h->E) return MOGP_ERR_ARG;
        if (!h->fitted[idx[i]]) {
            set_error("mogp_logpost_grad: output %d has not been fit", idx[i]);
            return MOGP_ERR_NOT_FIT;
        }
    }
    if (d > grad_max_dims()) {
        set_error("mogp_logpost_grad: at most %d input dimensions are supported", grad_max_dims());
        return MOGP_ERR_ARG;
    }
    API_CUDA(cudaSetDevice(h->device));
    NvtxRange nvtx_range("mogp_logpost_grad");
    const int64_t np = h->n_pad;
    const int T = (int)(np / NB);
    const size_t tiles = (size_t)T * (T + 1);
    // workspace of a group of G outputs: [G] Wt = (L^-1)^T matrices (np x np each, contiguous: one tensor map), then
    // [G] scratch blocks: per-tile partial sums (tiles x (d+2)) | gradient (d+2)
    const size_t scr_per = tiles * (d + 2) + (d + 2) + 6;
    const size_t per_out8 = ((size_t)np * np + scr_per + 7) / 8 * 8;
    // group size: as many outputs as the workspace budget holds (the inverse factors are the big part)
    int G = std::min<int>(count, MAXG);
    if (h->G_cap < sizeof(double) * per_out8 * G) {
        size_t free_b = 0, total_b = 0;
        API_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t budget = (size_t)((free_b + h->G_cap + pool_cached_bytes()) * 0.6);
        while (G > 1 && sizeof(double) * per_out8 * G > budget) G = (G + 1) / 2;
    }
    int rc;
    if ((rc = grow(&h->G, &h->G_cap, sizeof(double) * per_out8 * G, h->device))) return rc;
    if ((rc = grow(&h->h_grad, &h->h_grad_cap, sizeof(double) * (d + 2) * G, -1))) return rc;
    TrsmPlan plan = predict_plan_square(np, h->n_sms);
    if ((rc = grow(&h->sync, &h->sync_cap, predict_sync_bytes(plan, G, T), h->device))) return rc;
    for (int g0 = 0; g0 < count; g0 += G) {
        const int cnt = std::min(G, count - g0);
        API_CUDA(cudaEventRecord(h->ev_a, h->main));
        std::vector<int> outs(idx + g0, idx + g0 + cnt);
        double* Wt0 = h->G;
        double* scratch0 = h->G + (size_t)cnt * np * np;
        // L^-1 of every output of the group with ONE dataflow TRSM launch on identity right-hand sides
        CUtensorMap tmW;
        if (make_kblocked_tmap(&tmW, Wt0, (int64_t)cnt * np, np, plan.nw)) {
            set_error("tensor map (grad workspace) failed");
            return MOGP_ERR_CUDA;
        }
        for (int k = 0; k < cnt; k++)
            if (grad_set_identity(Wt0 + (size_t)k * np * np, np, h->main)) {
                set_error("gradient launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MOGP_ERR_CUDA;
            }
        if (predict_trsm(plan, outs.data(), cnt, h->maps.a128, h->maps.d128, tmW, Wt0, np, h->hyper, d, 0, np, np, nullptr, 0, 1,
                         (int*)h->sync, nullptr, h->n_sms, h->main)) {
            set_error("gradient launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return MOGP_ERR_CUDA;
        }
        for (int k = 0; k < cnt; k++) {
            double* Wt = Wt0 + (size_t)k * np * np;
            double* part = scratch0 + (size_t)k * scr_per;
            double* gdev = part + tiles * (d + 2);
            CUtensorMap tmW128, tmW64;
            if (make_kblocked_tmap(&tmW128, Wt, np, np, 128) || make_kblocked_tmap(&tmW64, Wt, np, np, 64)) {
                set_error("tensor map (grad workspace) failed");
                return MOGP_ERR_CUDA;
            }
            const int o = outs[k];
            if (grad_reduce_tiles(tmW128, tmW64, h->kernel, h->XT, h->n, np, d, h->alpha + (size_t)o * np,
                                  h->hyper + (size_t)o * (d + 2), h->nug_type == MOGP_NUG_FIT, part, gdev,
                                  h->n_u[o] ? h->U + (size_t)o * MAXM * np : nullptr, h->n_u[o], np, h->main)) {
                set_error("gradient launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return MOGP_ERR_CUDA;
            }
            API_CUDA(cudaMemcpyAsync(h->h_grad + (size_t)k * (d + 2), gdev, sizeof(double) * (d + 2), cudaMemcpyDeviceToHost,
                                     h->main));
        }
        API_CUDA(cudaEventRecord(h->ev_b, h->main));
        API_CUDA(cudaStreamSynchronize(h->main));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev_a, h->ev_b);
        h->timings[T_GRAD] += ms;
        h->timings[T_NLAUNCH] += 1 + 3 * cnt;
        for (int k = 0; k < cnt; k++)
            for (int i = 0; i < n_params; i++) grad[(size_t)(g0 + k) * n_params + i] = h->h_grad[(size_t)k * (d + 2) + i];
    }
    return MOGP_OK;
}

int mogp_loo_variance(mogp_handle* h, int32_t idx, double* out) {
    if (!h || idx < 0 || idx >= h->E || !out) return MOGP_ERR_ARG;
    if (!h->fitted[idx]) {
        set_error("mogp_loo_variance: output %d has not been fit", idx);
        return MOGP_ERR_NOT_FIT;
    }
    API_CUDA(cudaSetDevice(h->device));
    const int64_t np = h->n_pad;
    const int T = (int)(np / NB);
    int rc;
    // Wt = (L^-1)^T by the dataflow TRSM on an identity right-hand side (as for the gradient), then 1 / squared row norms
    if ((rc = grow(&h->G, &h->G_cap, sizeof(double) * ((size_t)np * np + np), h->device))) return rc;
    TrsmPlan plan = predict_plan_square(np, h->n_sms);
    if ((rc = grow(&h->sync, &h->sync_cap, predict_sync_bytes(plan, 1, T), h->device))) return rc;
    double* Wt = h->G;
    double* res = h->G + (size_t)np * np;
    CUtensorMap tmW;
    if (make_kblocked_tmap(&tmW, Wt, np, np, plan.nw)) {
        set_error("tensor map (loo workspace) failed");
        return MOGP_ERR_CUDA;
    }
    const int one[1] = {idx};
    if (grad_set_identity(Wt, np, h->main) ||
        predict_trsm(plan, one, 1, h->maps.a128, h->maps.d128, tmW, Wt, np, h->hyper, h->d, 0, np, np, nullptr, 0, 1,
                     (int*)h->sync, nullptr, h->n_sms, h->main) ||
        grad_row_inv_sumsq(Wt, np, h->n, res, h->main)) {
        set_error("loo variance launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return MOGP_ERR_CUDA;
    }
    API_CUDA(cudaMemcpyAsync(out, res, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->main));
    API_CUDA(cudaStreamSynchronize(h->main));
    h->timings[T_NLAUNCH] += 3;
    return MOGP_OK;
}

int mogp_logpost_grad(mogp_handle* h, int32_t idx, double* grad, int32_t n_params) {
    return mogp_logpost_grad_list(h, &idx, 1, grad, n_params);
}

int mogp_chol_schedule(int32_t ticket, int32_t n_block_rows, int32_t n_outputs, int32_t* out5) {
    if (!out5 || n_block_rows < 1 || n_outputs < 1 || ticket < 0 || (int64_t)ticket >= (int64_t)n_outputs * n_block_rows * (n_block_rows + 2))
        return MOGP_ERR_ARG;
    int tmp[5];
    chol_ticket(ticket, n_block_rows, n_outputs, tmp);
    for (int k = 0; k < 5; k++) out5[k] = tmp[k];
    return MOGP_OK;
}

int mogp_trim(void) {
    pool_trim();
    return MOGP_OK;
}

int mogp_timings(mogp_handle* h, double* out, int32_t n, int32_t reset) {
    if (!h || !out) return MOGP_ERR_ARG;
    for (int i = 0; i < n && i < T_COUNT; i++) out[i] = h->timings[i];
    if (reset) memset(h->timings, 0, sizeof(h->timings));
    return MOGP_OK;
}

// ---------------------------------------------------------------------------------------------
// NCCL
// ---------------------------------------------------------------------------------------------
int mogp_comm_unique_id(char* out128) {
    if (!out128) return MOGP_ERR_ARG;
    const NcclApi* api = nccl_api();
    if (!api) {
        set_error("libnccl.so.2 could not be loaded");
        return MOGP_ERR_NCCL;
    }
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) {
        set_error("ncclGetUniqueId failed");
        return MOGP_ERR_NCCL;
    }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out128, &id, 128);
    return MOGP_OK;
}

int mogp_comm_create(const char* uid128, int32_t rank, int32_t world, int32_t device, mogp_comm** out) {
    if (!uid128 || !out || rank < 0 || rank >= world) return MOGP_ERR_ARG;
    const NcclApi* api = nccl_api();
    if (!api) {
        set_error("libnccl.so.2 could not be loaded");
        return MOGP_ERR_NCCL;
    }
    API_CUDA(cudaSetDevice(device));
    mogp_comm* c = new mogp_comm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
        delete c;
        return MOGP_ERR_NCCL;
    }
    API_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    API_CUDA(cudaMalloc(&c->dscalar, sizeof(double) * 2));
    *out = c;
    return MOGP_OK;
}

int mogp_comm_destroy(mogp_comm* c) {
    if (!c) return MOGP_OK;
    const NcclApi* api = nccl_api();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (api && c->comm) api->CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    pool_free(c->sendbuf);
    pool_free(c->recvbuf);
    pool_free(c->h_recv);
    if (c->dscalar) cudaFree(c->dscalar);
    delete c;
    return MOGP_OK;
}

int mogp_comm_allreduce_max(mogp_comm* c, double* value) {
    if (!c || !value) return MOGP_ERR_ARG;
    const NcclApi* api = nccl_api();
    API_CUDA(cudaSetDevice(c->device));
    API_CUDA(cudaMemcpyAsync(c->dscalar, value, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = api->AllReduce(c->dscalar, c->dscalar + 1, 1, ncclDouble, ncclMax, c->comm, c->stream);
    if (r != ncclSuccess) {
        set_error("ncclAllReduce failed: %s", api->GetErrorString(r));
        return MOGP_ERR_NCCL;
    }
    API_CUDA(cudaMemcpyAsync(value, c->dscalar + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    API_CUDA(cudaStreamSynchronize(c->stream));
    return MOGP_OK;
}

int mogp_comm_allgather(mogp_comm* c, const double* send, int64_t count, double* recv) {
    if (!c || !send || !recv || count < 1) {
        set_error("mogp_comm_allgather: bad arguments");
        return MOGP_ERR_ARG;
    }
    const NcclApi* api = nccl_api();
    API_CUDA(cudaSetDevice(c->device));
    int rc;
    const size_t n = (size_t)count;
    if ((rc = grow(&c->sendbuf, &c->send_cap, sizeof(double) * n, c->device))) return rc;
    if ((rc = grow(&c->recvbuf, &c->recv_cap, sizeof(double) * n * c->world, c->device))) return rc;
    API_CUDA(cudaMemcpyAsync(c->sendbuf, send, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    ncclResult_t r = api->AllGather(c->sendbuf, c->recvbuf, n, ncclDouble, c->comm, c->stream);
    if (r != ncclSuccess) {
        set_error("ncclAllGather failed: %s", api->GetErrorString(r));
        return MOGP_ERR_NCCL;
    }
    API_CUDA(cudaMemcpyAsync(recv, c->recvbuf, sizeof(double) * n * c->world, cudaMemcpyDeviceToHost, c->stream));
    API_CUDA(cudaStreamSynchronize(c->stream));
    return MOGP_OK;
}

int mogp_predict_allgather(mogp_handle* h, mogp_comm* comm, const double* Xs, int64_t m, int32_t include_nugget,
                           int32_t e_pad, double* mean_all, double* var_all, int32_t* status_all) {
    if (!h || !comm || !Xs || m < 1 || e_pad < h->E || !mean_all || !var_all) {
        set_error("mogp_predict_allgather: bad arguments");
        return MOGP_ERR_ARG;
    }
    const NcclApi* api = nccl_api();
    API_CUDA(cudaSetDevice(h->device));
    NvtxRange nvtx_range("mogp_predict_allgather");
    TraceClock tc("mogp_predict_allgather");
    int rc = predict_device(h, Xs, m, 1, include_nugget);
    if (rc) return rc;
    tc.mark("predict_device (kernels + sync)");
    const size_t blk = (size_t)e_pad * 2 * m;  // doubles per rank; one extra row-pair block carries the status words
    const size_t send_n = blk + e_pad;
    if ((rc = grow(&comm->sendbuf, &comm->send_cap, sizeof(double) * send_n, comm->device))) return rc;
    if ((rc = grow(&comm->recvbuf, &comm->recv_cap, sizeof(double) * send_n * comm->world, comm->device))) return rc;
    if ((rc = grow(&comm->h_recv, &comm->h_cap, sizeof(double) * send_n * comm->world, -1))) return rc;
    // pack: [e_pad][2][m] results (NaN for padding rows) + e_pad status words (as doubles)
    API_CUDA(cudaMemsetAsync(comm->sendbuf, 0xFF, sizeof(double) * blk, h->main));
    API_CUDA(cudaMemcpyAsync(comm->sendbuf, h->res, sizeof(double) * (size_t)h->E * 2 * m, cudaMemcpyDeviceToDevice, h->main));
    std::vector<double> st(e_pad, (double)MOGP_ERR_ARG);
    for (int o = 0; o < h->E; o++) st[o] = h->fitted[o] ? (double)MOGP_OK : (double)MOGP_ERR_NOT_FIT;
    API_CUDA(cudaMemcpyAsync(comm->sendbuf + blk, st.data(), sizeof(double) * e_pad, cudaMemcpyHostToDevice, h->main));
    // the single collective of the path: all-gather of every rank's packed posterior block
    ncclResult_t r = api->AllGather(comm->sendbuf, comm->recvbuf, send_n, ncclDouble, comm->comm, h->main);
    if (r != ncclSuccess) {
        set_error("ncclAllGather failed: %s", api->GetErrorString(r));
        return MOGP_ERR_NCCL;
    }
    API_CUDA(cudaMemcpyAsync(comm->h_recv, comm->recvbuf, sizeof(double) * send_n * comm->world, cudaMemcpyDeviceToHost, h->main));
    tc.mark("pack + enqueue all-gather + D2H");
    API_CUDA(cudaStreamSynchronize(h->main));
    tc.mark("all-gather + D2H complete");
    {
        std::vector<std::pair<double*, const double*>> rows;
        for (int rk = 0; rk < comm->world; rk++) {
            const double* src = comm->h_recv + (size_t)rk * send_n;
            for (int o = 0; o < e_pad; o++) {
                const size_t row = (size_t)rk * e_pad + o;
                rows.emplace_back(mean_all + row * m, src + (size_t)o * 2 * m);
                rows.emplace_back(var_all + row * m, src + (size_t)o * 2 * m + m);
                if (status_all) status_all[row] = (int32_t)src[blk + o];
            }
        }
        copy_rows(rows, sizeof(double) * m);
    }
    tc.mark("host unpack");
    return MOGP_OK;
}

}  // extern "C"
