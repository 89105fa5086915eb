// Internal host-side interface between the translation units of libmogp_b200.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace mogp {

constexpr int MAXG = 64;  // outputs handled by one batched predict launch

struct CholMaps {
    CUtensorMap a128;  // K-blocked map over the matrix slab [E*n_pad][n_pad], box 128 rows
    CUtensorMap a64;   // same slab, box 64 rows
    CUtensorMap d128;  // K-blocked map over the Dinv slab [E*n_pad][128], box 128 rows
    CUtensorMap d8;    // same slab, box 128 rows x one 8-column K slab
};

// per-output hyperparameters in device memory: [w_0 .. w_{d-1}, sigma2, nugget]
//   w_i = exp(theta_i), sigma2 = exp(theta_d)   (GPParams.py:3-161)

// ---- chol.cu ----
int chol_init();
int chol_make_maps(CholMaps* maps, double* A_slab, double* Dinv_slab, int64_t total_rows, int64_t n_pad);
size_t chol_sync_bytes(int count, int T);
// factorise `count` outputs (slab indices outs[]) in place as one dataflow launch on `st`; info[o] / scal[2*o]
// must be zero.  Returns #launches or -1.
int chol_factor_batch(const CholMaps& maps, double* A_slab, double* Dinv_slab, const int* outs, int count,
                      int64_t n_pad, int* info, double* scal, int* sync, int n_sms, cudaStream_t st);

// the same with the history products on the int8 tensor cores (tcgen05, 8 planes per operand); also writes the planes of the
// strictly lower blocks of L to Lq (layout of i8_slice_L; exps[k] = i8_scale_exponent of outs[k])
int chol_i8_factor_batch(const CholMaps& maps, double* A_slab, double* Dinv_slab, const int* outs, const int* exps, int count,
                         int64_t n_pad, int8_t* Lq, int64_t lq_stride, int* info, double* scal, int* sync, int n_sms,
                         cudaStream_t st);
int chol_i8_planes();

// ticket t of a launch over `count` outputs with T block rows -> {kind (0 DIAG, 1 D, 2 ROW), output, block row i, half p, column j}
void chol_ticket(int t, int T, int count, int out[5]);

// ---- kmat.cu ----
int kmat_init();
int kmat_dbox(int d);
int kmat_sym(const CUtensorMap& tmXT, int kernel, int64_t n, int64_t n_pad, int d, const double* hyper, const int* outs,
             int count, int add_nugget, double* A_slab, int64_t slab_rows_per_output, cudaStream_t st, int* inf_flag = nullptr);
int kmat_cross(const CUtensorMap& tmXsT, const CUtensorMap& tmXT, int kernel, int64_t n, int64_t n_pad, int64_t m_pad,
               int d, const int* outs, int count, const double* hyper, double* W_slab, int64_t w_stride, int store,
               const double* alpha, int64_t alpha_stride, double* part, cudaStream_t st, int* inf_flag = nullptr);
// inf_flag (device, optional): set to 1 when any squared distance is +inf -- the reference raises FloatingPointError there
int mean_reduce(const double* part, const int* outs, int count, int n_tiles, int64_t m_pad, int64_t m, double* mean,
                int64_t mean_stride, cudaStream_t st);

int kderiv_max_dims();
// d mean / d x* of `count` outputs at m test points: out[k][c][q], k-th listed output (XsT: [d][xs_stride] on the device)
int kmat_deriv(int kernel, const double* XsT, int64_t xs_stride, const double* XT, int64_t n, int64_t n_pad, int64_t m,
               int d, const int* outs, int count, const double* hyper, const double* alpha, double* out, cudaStream_t st);

// ---- solve.cu ----
int solve_init();
int solve_alpha(const double* A_slab, int64_t n_pad, const double* Dinv_slab, const double* Y, double* z, double* alpha,
                double* scal, const int* info, const int* outs, int count, int n_sms, cudaStream_t st);

// ---- predict.cu ----
int predict_init();
struct TrsmPlan {
    int nw;      // test points per tile (32 or 64)
    int panels;  // panels per output
};
TrsmPlan predict_plan(int64_t m, int n_outputs, int n_pad, int n_sms);
TrsmPlan predict_plan_square(int64_t n_pad, int n_sms);
// bytes of the ticket/flag workspace one launch needs (zeroed by predict_trsm itself)
size_t predict_sync_bytes(const TrsmPlan& plan, int count, int T);
int predict_trsm(const TrsmPlan& plan, const int* outs, int count, const CUtensorMap& tmL, const CUtensorMap& tmD,
                 const CUtensorMap& tmW, double* W, int64_t w_stride, const double* hyper, int d, int include_nugget,
                 int64_t n_pad, int64_t m, double* var, int64_t var_stride, int tri_rhs, int* sync, double* normacc,
                 int n_sms, cudaStream_t st, int keep_v = 0, int no_clip = 0, int diag_only = 0);

// ---- trsm_i8.cu: the predict TRSM from int8 slices on tcgen05 (many right-hand sides) ----
int i8_init();
// S = planes per operand: 6 (pairs t + u <= 7) or 7 (t + u <= 8, 128 x finer)
size_t i8_lq_bytes(int T, int S);                           // planes of L per output (strictly lower blocks)
size_t i8_vq_bytes(int count, int panels, int T, int S);    // planes of V for one batched call
size_t i8_sync_bytes(int count, int panels);                // ticket counter + per-panel progress words
int i8_panel_width();
// exponent e of the common power-of-two scale of L and V of one output: sqrt(sigma2 + nugget) <= 0.99 2^e
int i8_scale_exponent(double sigma2, double nugget);
// planes of the strictly lower 128 x 128 blocks of L of the listed outputs (exps[k] = i8_scale_exponent of outs[k])
int i8_slice_L(int S, const double* A_slab, int64_t n_pad, const int* outs, const int* exps, int count, int8_t* Lq,
               int64_t lq_stride, cudaStream_t st);
// forward substitution + variances, one persistent launch; W holds K* (test-major) and is only read (tmW: its K-blocked map,
// box 64 rows); tmD: K-blocked map over the Dinv slab (box 128 rows); sync: i8_sync_bytes(count, panels) bytes (zeroed by the call)
int i8_trsm(int S, const int* outs, const int* exps, int count, int panels, const int8_t* Lq, int64_t lq_stride, int8_t* Vq,
            const CUtensorMap& tmD, const CUtensorMap& tmW, const double* W, int64_t w_stride, const double* hyper, int d,
            int include_nugget, int no_clip, int64_t n_pad, int64_t m, double* var, int64_t var_stride, double* normacc, int* sync,
            int n_sms, cudaStream_t st);

// a-posteriori accuracy check of the int8 path: i8_check_points() test points per output (spread evenly over the m of the
// call) are gathered from K* into Wc [count][points][n_pad] for an FP64 solve (predict_trsm), then compared:
// ratio[k] = max over the samples of |var - var_ref| / (1 % of the parity bar + FP64 rounding floor); > 1 fails
int i8_check_points();
int i8_check_gather(const int* outs, int count, const double* W, int64_t w_stride, int64_t n_pad, int64_t m, double* Wc,
                    cudaStream_t st);
int i8_check_compare(const int* outs, int count, int64_t m, const double* var, int64_t var_stride, const double* var_ref,
                     const double* hyper, int d, double* ratio, cudaStream_t st);

// ---- grad.cu ----
int grad_init();
int grad_max_dims();
int grad_set_identity(double* W, int64_t n_pad, cudaStream_t st);
int grad_row_inv_sumsq(const double* W, int64_t n_pad, int64_t n, double* out, cudaStream_t st);
int grad_reduce_tiles(const CUtensorMap& tmW128, const CUtensorMap& tmW64, int kernel, const double* XT, int64_t n,
                      int64_t n_pad, int d, const double* alpha, const double* hyper, int fit_nugget, double* partial,
                      double* grad, const double* U, int n_u, int64_t u_stride, cudaStream_t st);
int grad_max_mean();

int cov_syrk_sub(const CUtensorMap& tmW128, const CUtensorMap& tmW64, int row_base, int64_t m_pad, int64_t k_len, double* C,
                 cudaStream_t st);

}  // namespace mogp
