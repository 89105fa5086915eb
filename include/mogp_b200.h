/* libmogp_b200 -- C ABI of the B200-native GP fit+predict hot path.
 *
 * Drop-in boundary: these entry points are what the reference's Python front-ends
 * (mogp_emulator/GaussianProcessGPU.py, MultiOutputGP_GPU.py, fitting.py) obtain today from the
 * pybind11 module `libgpgpu` (mogp_gpu/src/bindings.cu).  Each function names the reference
 * interface it replaces.  Plain pointers and sizes only; the caller allocates every output
 * (the reference's Eigen::Ref convention, GaussianProcessGPU.py:476-490, 604-612); the library
 * never keeps a caller pointer after the call returns.  All arrays are C-order float64.
 *
 * All functions return an int status (MOGP_OK == 0).  mogp_last_error() gives the message of the
 * last failure on the calling thread.  A handle is not re-entrant.
 */
#ifndef MOGP_B200_H
#define MOGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
enum {
    MOGP_OK = 0,
    MOGP_ERR_CUDA = 1,     /* CUDA runtime/driver failure (no device, launch error, ...)          */
    MOGP_ERR_ARG = 2,      /* bad argument (reference: std::runtime_error, densegp_gpu.hpp:495)    */
    MOGP_ERR_NOT_PD = 3,   /* factorisation failed (densegp_gpu.hpp:556-570, cholesky.py:219,281)  */
    MOGP_ERR_NOT_FIT = 4,  /* predict/get on an output whose hyperparameters are not set           */
    MOGP_ERR_NCCL = 5,
    MOGP_ERR_NOMEM = 6,
    MOGP_ERR_FPE = 7       /* infinite squared distance in the kernel (Kernel.py:482-483 raises FloatingPointError) */
};

/* mogp_gpu/src/types.hpp:29-33 -- the numeric values are part of the reference's Python contract
 * (MultiOutputGP_GPU.py:74-79). */
enum { MOGP_NUG_ADAPTIVE = 0, MOGP_NUG_FIT = 1, MOGP_NUG_FIXED = 2 };
enum { MOGP_KERNEL_SQEXP = 0, MOGP_KERNEL_MATERN52 = 1 };

/* selectors for mogp_get */
enum { MOGP_GET_K = 0, MOGP_GET_L = 1, MOGP_GET_ALPHA = 2, MOGP_GET_KINV = 3 };

typedef struct mogp_handle mogp_handle; /* one emulator bank: E outputs over shared inputs, on one GPU */
typedef struct mogp_comm mogp_comm;     /* one NCCL communicator rank                                  */

/* replaces libgpgpu.have_compatible_device (mogp_gpu/src/util.cu, bindings.cu:600). */
int mogp_device_count(int32_t* count);

const char* mogp_last_error(void);

/* replaces DenseGP_GPU(inputs, targets, testing_size, meanfunc, kernel, nugget_type, nugsize)
 * (densegp_gpu.hpp:777-802) and MultiOutputGP_GPU(inputs, targets[], ...) (multioutputgp_gpu.hpp:
 * 259-265).  X is (n, d); Y is (n_out, n); zero mean function.  The handle owns device copies.
 * n_streams is ignored (kept for ABI stability: every phase is one batched launch over the handle's outputs). */
int mogp_create(const double* X, int64_t n, int32_t d, const double* Y, int32_t n_out, int32_t kernel,
                int32_t nugget_type, double nugget, int32_t device, int32_t n_streams, mogp_handle** out);
int mogp_destroy(mogp_handle* h);
/* Handles return their device / pinned buffers to a process-wide cache for reuse; this releases the cache. */
int mogp_trim(void);

/* replaces DenseGP_GPU::fit(theta) (densegp_gpu.hpp:493-612) / MultiOutputGP_GPU::fit(thetas),
 * fit_emulator(idx, theta) (multioutputgp_gpu.hpp:150-181); semantics follow the CPU
 * GaussianProcess.fit (GaussianProcess.py:629-685) and cholesky_factor / jit_cholesky
 * (linalg/cholesky.py:168-281).
 * thetas: (count, n_params) raw hyperparameters [theta_corr(d), theta_cov, (theta_nugget)] for
 * outputs first..first+count-1.  Per output it reports
 *   quad   = y^T K^-1 y,   logdet = log det(K + nugget I),   nugget = the nugget actually used,
 *   status = MOGP_OK | MOGP_ERR_NOT_PD (that output is then marked not fit; others proceed).
 * The data part of the reference's current_logpost is 0.5*(quad + logdet + n*log(2 pi)); priors are
 * host-side scalars added by the caller.  Return value is MOGP_OK unless the call itself failed.
 * Arithmetic: IEEE FP64 throughout, except that a call with enough factorisation work (outputs x (n/128)^2 >= 4096) evaluates
 * the O(n^3) history products of the Cholesky as EXACT integer GEMM on 8 signed 7-bit planes per operand (tcgen05 int8 tensor
 * cores; products resolved to 2^-63 of the squared scale -- the backward error of an FP64 blocked factorisation); pivots, info,
 * the adaptive-nugget decisions, triangular solves and the log-determinant are FP64.  MOGP_CHOL_I8=0: FP64 tensor pipe only. */
int mogp_fit(mogp_handle* h, int32_t first, int32_t count, const double* thetas, int32_t n_params,
             double* quad_out, double* logdet_out, double* nugget_out, int32_t* status_out);

/* Same for an arbitrary list of distinct outputs idx[0..count) (thetas, and every result array, in list order): the
 * batched objective evaluation of concurrent MAP fits of a multi-output emulator (fitting.py:189-217, 273-340). */
int mogp_fit_list(mogp_handle* h, const int32_t* idx, int32_t count, const double* thetas, int32_t n_params,
                  double* quad_out, double* logdet_out, double* nugget_out, int32_t* status_out);

/* marks outputs not fit (MultiOutputGP_GPU.reset_fit_status, GaussianProcessGPU theta=None). idx<0: all */
int mogp_reset(mogp_handle* h, int32_t idx);
int mogp_is_fit(mogp_handle* h, int32_t idx, int32_t* out);

/* replaces DenseGP_GPU::predict_batch / predict_variance_batch (densegp_gpu.hpp:300-408) and
 * MultiOutputGP_GPU::predict_batch / predict_variance_batch (multioutputgp_gpu.hpp:183-228);
 * values follow the CPU GaussianProcess.predict (GaussianProcess.py:889-920): variance clipped at 0.
 * Xs: (m, d).  mean, var: (n_out, m) caller-allocated; var may be NULL when want_var == 0; want_var == 2 returns the
 * variance before the clip at 0 (the front-end adds the mean-function term first, GaussianProcess.py:913-920).
 * status: (n_out) -- MOGP_ERR_NOT_FIT rows are filled with NaN (MultiOutputGP.py:476-546). */
int mogp_predict(mogp_handle* h, const double* Xs, int64_t m, int32_t want_var, int32_t include_nugget,
                 double* mean, double* var, int32_t* status);

/* replaces DenseGP_GPU::predict_deriv (densegp_gpu.hpp:411-448) / MultiOutputGP_GPU::predict_deriv
 * (multioutputgp_gpu.hpp:230-257): derivative of the posterior mean with respect to the test inputs.
 * deriv: (n_out, m, d) caller-allocated; rows of unfit outputs are NaN. */
int mogp_predict_deriv(mogp_handle* h, const double* Xs, int64_t m, double* deriv, int32_t* status);

/* Full predictive covariance of ONE output at m test points (the CPU GaussianProcess.predict(full_cov=True),
 * GaussianProcess.py:899-911; the reference GPU class has no equivalent):
 *   cov = sigma2 k(X*, X*) [+ nugget I] - K*^T K^-1 K*,   (m, m), not clipped;   mean: (m).
 * V = L^-1 K* by the dataflow TRSM, then one FP64 tensor-pipe SYRK. */
int mogp_predict_cov(mogp_handle* h, int32_t idx, const double* Xs, int64_t m, int32_t include_nugget, double* mean,
                     double* cov);

/* Sharded multi-output predict: every rank predicts its own outputs, then ONE ncclAllGather of the
 * packed per-rank [e_pad][2][m] block delivers all ranks' means and variances to every rank.
 * mean_all, var_all: (world * e_pad, m); status_all: (world * e_pad) (MOGP_ERR_ARG marks padding rows). */
int mogp_predict_allgather(mogp_handle* h, mogp_comm* comm, const double* Xs, int64_t m, int32_t include_nugget,
                           int32_t e_pad, double* mean_all, double* var_all, int32_t* status_all);

/* replaces DenseGP_GPU::get_K / get_cholesky_lower / get_invQt / get_invQ (densegp_gpu.hpp:478,624-637).
 * K, L, KINV: (n, n) row-major (L lower-triangular, upper zero); ALPHA: (n). */
int mogp_get(mogp_handle* h, int32_t idx, int32_t which, double* out);

/* replaces DenseGP_GPU::logpost_deriv (densegp_gpu.hpp:663-770); formula of the CPU
 * GaussianProcess.logpost_deriv (GaussianProcess.py:711-782) for zero mean, evaluated at the theta of
 * the last mogp_fit of output idx:  grad[i] = 0.5*(tr(K^-1 dK_i) - alpha^T dK_i alpha), i over
 * [corr(d), cov, (nugget)].  Prior terms are added by the caller. */
int mogp_logpost_grad(mogp_handle* h, int32_t idx, double* grad, int32_t n_params);

/* Gradients of several fitted outputs at once: grad is (count, n_params) in list order.  L^-1 of all listed outputs comes
 * from one dataflow launch. */
int mogp_logpost_grad_list(mogp_handle* h, const int32_t* idx, int32_t count, double* grad, int32_t n_params);

/* out[c] (n values) = leave-one-out predictive variance of training point c of a fitted output, 1 / (K^-1)_cc with
 * K = sigma2 k(X,X) + nugget I: the quantity MICEFastGP.fast_predict(index) evaluates one index (and one O(n^3) inverse) at a
 * time (SequentialDesign.py:705-748); here one L^-1 for all n points. */
int mogp_loo_variance(mogp_handle* h, int32_t idx, double* out);

/* ---- analytic mean function (the CPU GaussianProcess with a design matrix H, GaussianProcess.py:657-685, 887-920;
 * linalg_utils.py:5-168 calc_Ainv / calc_mean_params / calc_R).  The host front-end keeps H and the n_mean x n_mean
 * algebra; these are the device primitives it needs.  All index lists are distinct handle-local outputs. ---- */
/* out[i] = K_i^-1 rhs[i] for fitted outputs (rhs, out: (count, n)) -- Kinv.solve(dm) */
int mogp_solve_list(mogp_handle* h, const int32_t* idx, int32_t count, const double* rhs, double* out);
/* replace the stored K^-1 y of the listed outputs (used by the posterior mean, its derivative and the gradient) with
 * alpha[i] (count, n) -- Kinv_t_mean = K^-1 (y - H beta) */
int mogp_set_alpha_list(mogp_handle* h, const int32_t* idx, int32_t count, const double* alpha);
/* vectors u_q (count, n_vec, n), n_vec <= 32, with K^-1 H A^-1 H^T K^-1 = sum_q u_q u_q^T: mogp_logpost_grad_list then
 * differentiates the mean-integrated likelihood.  Every mogp_fit of an output clears its vectors. */
int mogp_set_mean_vectors_list(mogp_handle* h, const int32_t* idx, int32_t count, int32_t n_vec, const double* U);
/* out[o][q][c] = sum_i k_o(x*_c, x_i) vecs[o][q][i] for every fitted output o (vecs: (n_out, n_vec, n); out: (n_out, n_vec,
 * m), NaN rows for unfit outputs) -- H^T K^-1 K* of calc_R, without materialising K*. */
int mogp_kstar_dot(mogp_handle* h, const double* Xs, int64_t m, const double* vecs, int32_t n_vec, double* out);

/* Introspection of the factorisation's tile schedule (no device needed): ticket -> out5 = {kind (0 = DIAG accumulation of a
 * diagonal block half, 1 = D factor + invert the diagonal block, 2 = ROW panel tile), output, block row i, half p, block column j}
 * for a launch over n_outputs matrices of n_block_rows x n_block_rows 128-blocks; tickets run 0 .. n_outputs * T * (T + 2) - 1.
 * The persistent kernel draws tickets in this order and spin-waits on the tiles a tile depends on, so the order must be
 * topological; tests/test_host.py checks that on the CPU. */
int mogp_chol_schedule(int32_t ticket, int32_t n_block_rows, int32_t n_outputs, int32_t* out5);

/* accumulated device time per phase in ms since the last call with reset != 0:
 * out[0..10] = kmat, cholesky, solves, kstar, predict_trsm, grad, n_trsm_launches, n_kernel_launches, fit (device),
 *              predict host wall up to the last kernel's completion, predict result copy-out host wall;
 * out[11..15] = int8 path of the predict TRSM (inside predict_trsm): slicing L into planes (ms), the FP64 solve of the
 *               sampled test points for the a-posteriori accuracy check (ms), the persistent integer TRSM kernel (ms), block
 *               rows solved by it, number of groups the check sent back to the FP64 kernel
 * out[16]     = factorisations whose history products ran on the int8 tensor cores (chol_i8_kernel)
 * out[17..18] = of those, the ones that reported "not positive definite" and were therefore repeated on the FP64 kernel (its
 *               verdict is the one returned), and how many of these the FP64 kernel then factorised successfully */
int mogp_timings(mogp_handle* h, double* out, int32_t n, int32_t reset);

/* NCCL plumbing (no reference equivalent: the reference is single-device, multioutputgp_gpu.hpp:183). */
int mogp_comm_unique_id(char* out128);
int mogp_comm_create(const char* uid128, int32_t rank, int32_t world, int32_t device, mogp_comm** out);
int mogp_comm_destroy(mogp_comm* c);
/* max-reduce of one double over ranks + barrier (bench timing). */
int mogp_comm_allreduce_max(mogp_comm* c, double* value);
/* all-gather of `count` doubles per rank between HOST buffers (recv: world * count doubles, rank-major): one ncclAllGather
 * bracketed by the two copies.  Carries what mogp_predict_allgather does not: fit status after a sharded fit, and the
 * posteriors of sharded emulators with a mean function or predictive derivatives (finished on the host per rank).  A rank
 * that holds no outputs joins mogp_predict_allgather's collective through this entry point with a padding block of the
 * same length. */
int mogp_comm_allgather(mogp_comm* c, const double* send, int64_t count, double* recv);

/* measured FP64 tensor-pipe (DMMA) issue peak of `device` in TFLOP/s: the roofline denominator bench.py uses
 * (MEASURED_PEAKS.json has no FP64 entry). */
int mogp_peak_dmma(int32_t device, int32_t iters, double* tflops);
/* int8 tcgen05 (kind::i8) issue peaks in TOP/s: tops[0] = M128 N256 K32 MMAs (the chip's int8 tensor peak), tops[1] = the
 * M128 N64 K32 shape of the predict kernel (csrc/trsm_i8.cu) */
int mogp_peak_i8(int32_t device, int32_t iters, double* tops);

int mogp_version(int32_t* major, int32_t* minor);

#ifdef __cplusplus
}
#endif
#endif /* MOGP_B200_H */
