// Shared declarations of libklnmf (B200 / sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/klnmf.h"

#define KL_EPS      1.0e-8     // nmf.py:232,297,325 literal default
#define KL_NORM_EPS 1.0e-16    // array_utils.py:19

namespace klnmf {

void set_error(const char *fmt, ...);

#define KL_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      klnmf::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return KLNMF_ECUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define KL_CHECK(cond, code, ...)     \
  do {                                \
    if (!(cond)) {                    \
      klnmf::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

#define KL_TRY(expr)              \
  do {                            \
    int _r = (expr);              \
    if (_r != KLNMF_OK) return _r; \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// device scalars (double) kept in ctx->dscal
enum { DS_KL = 0, DS_PREV = 1, DS_WHSUM = 2, DS_SUMX = 3, DS_TOL = 4, DS_COUNT = 8 };
// device flags (int) kept in ctx->flags
enum { FL_STOP = 0, FL_NERR = 1, FL_NEG = 2, FL_NONFINITE = 3, FL_COUNT = 8 };
// profile phases
enum { PH_RATIO = 0, PH_COEF = 1, PH_NUM = 2, PH_DICT = 3, PH_COMM = 4, PH_TOTAL = 5 };

}  // namespace klnmf

struct klnmf_ctx {
  int device = 0;
  int64_t n = 0, f = 0, k = 0;
  int mode = 0;
  int es = 4;                  // element size of device state (4: TF32 modes, 8: FP64)
  bool debug_simt = false;     // KLNMF_DEBUG_ENGINE=simt: FP32 FMA bring-up engine (never default)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;

  // ---- data -------------------------------------------------------------------------
  bool have_x = false, sparse = false;
  void *X = nullptr;           // dense n x f
  int64_t ldx = 0;
  bool x_owned = false;
  int64_t nnz = 0;
  int64_t *indptr = nullptr;   // n+1
  int32_t *indices = nullptr;  // nnz
  void *vals = nullptr;        // nnz
  bool csr_owned = false;
  void *qnz = nullptr;         // nnz ratio values (sparse path)
  void *bcsc = nullptr;        // blocked-CSC copy of the pattern for the numerator pass (sparse.cu), built lazily
  void *hyb = nullptr;         // hybrid stack: the dense block's side (api.cu: HybridSide); ctx->f then counts the CSR columns
  int64_t hybrid_min_cols = 1024;   // dense columns from which klnmf_set_stacked_blocks_host keeps the dense blocks dense

  // ---- state -------------------------------------------------------------------------
  bool have_h = false, have_w = false;
  int cur = 0;                 // index of the current W in the ping-pong buffers
  int hcur = 0;                // index of the current dictionary (flips only while fitting)
  bool split = false;          // TF32X3: arrays hold the tf32-exact high part, *lo the residual
  void *W[2] = {nullptr, nullptr};
  int64_t ldw = 0;             // n x k
  void *H[2] = {nullptr, nullptr};
  int64_t ldh = 0;             // dense: k x f (ld >= f);  sparse: f x k transposed (ld >= k)
  void *Wlo[2] = {nullptr, nullptr};   // TF32X3 low parts (same layout)
  void *Hlo[2] = {nullptr, nullptr};
  void *num = nullptr;         // numerator accumulator, same layout as H -- or, dense path on several ranks, in f-chunks:
  int64_t num_bytes = 0;       //   [chunk][k][num_cc], every chunk contiguous, so that a finished chunk is all-reduced
  int num_chunks = 1;          //   on a side stream while the next one is still being contracted (api.cu)
  int64_t num_cc = 0;          //   columns per chunk (multiple of 32); 0 = plain k x ldh layout
  bool num_reduced = false;    // this iteration's chunks were all-reduced as they finished
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t comm_ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t comm_done = nullptr;
  double *rowsumH = nullptr;   // k   (sum over f of H, used by the sparse objective)
  double *colsumW = nullptr;   // k   (dense, centered ratio: column sums of the new coefficients, all ranks)
  float *rsh32 = nullptr;      // ldw + 32 floats: rowsumH as FP32, zero beyond k (bias of the centered coefficient update)
  bool centered = false;       // this iteration's ratio panel holds Q - 1 (see dense_iteration)
  bool single_pass = false;    // TF32R: (hi, lo) state like TF32X3, but the contractions of the loop multiply the hi parts only
  double *hsum = nullptr;      // k   scratch of the normaliser
  double *dred = nullptr;      // [kl, sum(X.data), colsum(W)[0..ldw)] : the doubles that are all-reduced
  int64_t dred_len = 0;
  void *stage = nullptr;       // device staging for host<->device conversions
  int64_t stage_bytes = 0;
  void *pin_stage[2] = {nullptr, nullptr};   // borrowed process-wide pinned staging of large downloads (api.cu: download_large)

  // ---- dense scratch -------------------------------------------------------------------
  void *Q = nullptr, *Qlo = nullptr;   // panel_rows x f ratio panel
  int64_t ldq = 0, panel_rows = 0;
  int64_t scratch_limit = (int64_t)16 << 30;

  // ---- control ---------------------------------------------------------------------------
  double *dscal = nullptr;     // DS_COUNT doubles
  int *flags = nullptr;        // FL_COUNT ints
  double *errors_dev = nullptr;
  int errors_cap = 0;
  double *pinned = nullptr;    // small pinned host buffer for read-backs

  // ---- multi GPU ----------------------------------------------------------------------------
  void *comm = nullptr;        // ncclComm_t
  bool comm_owned = false;     // created by klnmf_comm_init (destroyed with the context) vs attached (klnmf_comm_attach)
  int rank = 0, world = 1;

  // ---- accounting -----------------------------------------------------------------------------
  int64_t n_launch = 0, n_nccl = 0, bytes_h2d = 0, bytes_d2h = 0;
  double prof_ms[6] = {0, 0, 0, 0, 0, 0};
  int64_t prof_cnt[5] = {0, 0, 0, 0, 0};
  bool profile = false;        // per-phase events (adds syncs at the end only)
  void *tc = nullptr;          // tcgen05 engine private state (tensor maps, ...)
  void *fused = nullptr;       // fused half-step kernel private state
  void *Ht = nullptr;          // dense TF32 mode, k <= 128: transposed dictionary f x ldht for the fused kernel
  int64_t ldht = 0;
  int ht_of = -1;              // which H buffer Ht mirrors (-1: none)
  bool ht_stale = true;        // the dictionary changed since Ht was written
};

namespace klnmf {

// ---- generic (DMMA fp64 / FMA fp32 bring-up) contraction engine: dense_generic.cu -----------
// C = op(A) . op(B) with arbitrary strides and a fused epilogue.
enum Epilogue {
  EPI_STORE = 0,   // out = C
  EPI_RATIO = 1,   // out = (X+eps)/(C+eps), kl += X*log(out) - X + C        (nmf.py:325-336, metrics.py:18-20)
  EPI_MULW = 2,    // out = aux * C                                           (nmf.py:338-343)
  EPI_ACC = 3      // out += C  (atomic, split over the contraction)          (nmf.py:345-349)
};
struct GemmDesc {
  int64_t M, N, K;
  const void *A; int64_t a_sm, a_sk;     // A(m,kk) = A[m*a_sm + kk*a_sk]
  const void *B; int64_t b_sk, b_sn;     // B(kk,n) = B[kk*b_sk + n*b_sn]
  void *out; int64_t ldo;                // row-major M x N
  const void *aux; int64_t ldaux;        // X (EPI_RATIO) or W (EPI_MULW), row-major M x N
  double *kl;                            // objective accumulator (EPI_RATIO) or nullptr
  const int *stop;                       // device stop flag: kernel exits when *stop != 0
  int splitk;                            // EPI_ACC only
  int only_kl;                           // EPI_RATIO: do not write out (klnmf_error)
  float qshift;                          // EPI_RATIO: the stored ratio is q - qshift (1 = centered, 0 = plain)
  int single_pass;                       // multiply the hi parts only (A_lo, B_lo ignored); out_lo / aux_lo still honoured
  int round_out;                         // EPI_RATIO without out_lo: store the ratio rounded to nearest TF32
  const float *colbias;                  // EPI_MULW: out = aux * (C + colbias[col]) (nullptr: no bias)
  // split-TF32 residual arrays (same layout as their high parts); nullptr unless ctx->split
  const void *A_lo; const void *B_lo; void *out_lo; const void *aux_lo;
};
int generic_gemm(klnmf_ctx *ctx, int es, int epi, const GemmDesc &d);

// ---- tcgen05 engine: dense_tc.cu ---------------------------------------------------------------
int tc_gemm(klnmf_ctx *ctx, int epi, const GemmDesc &d);
int tc_selftest(int *n_fail, char *report, int report_len);
void tc_release(klnmf_ctx *ctx);

// ---- fused coefficient half-step (k <= 128, TF32): dense_fused.cu --------------------------------
struct FusedDesc {
  int64_t M, F, K;
  const void *W; int64_t ldw;        // current coefficients M x K (zero padded to ldw)
  const void *H; int64_t ldh;        // dictionary K x F
  const void *Ht; int64_t ldht;      // its transpose F x K (zero padded to ldht)
  const void *X; int64_t ldx;        // data M x F
  void *Wout; int64_t ldwo;          // updated coefficients
  void *Q; int64_t ldq;              // fit only: the ratio panel M x F is also written (nullptr: transform)
  double *kl;                        // objective accumulator or nullptr
  const int *stop;
  int only_kl;
  float qshift;                      // the ratio tile holds q - qshift
  const float *colbias;              // W' = W (.) (G + colbias[component]) (nullptr: no bias)
  // TF32R: W is a (hi, lo) pair, hi the round-to-nearest TF32 value the contractions multiply; the ratio tile is
  // rounded to nearest before the second contraction reads it, the objective takes its cancellation-free form
  const void *Wlo;                   // low parts of W (nullptr: W is plain FP32)
  void *Wout_lo;                     // low parts of the updated coefficients (nullptr: plain FP32 output)
  int accurate;                      // cancellation-free objective + rounded ratio tile
};
bool fused_supported(const klnmf_ctx *ctx, int fit);
int fused_coef_step256(klnmf_ctx *ctx, const FusedDesc &d, int *err_dev);   // 128 < k <= 256, transform, CTA pairs
int fused_coef_step(klnmf_ctx *ctx, const FusedDesc &d);
void fused_release(klnmf_ctx *ctx);

// ---- elementwise / reductions: elementwise.cu ----------------------------------------------------
int launch_dict_update(klnmf_ctx *ctx, const void *H_old, void *H_new, void *Hlo_new, const double *rowadd = nullptr);
int launch_scale_block(klnmf_ctx *ctx, void *X, int64_t ld, int64_t rows, int64_t cols, double scale, int f32 = 0);   // X[:, :cols] *= scale
int launch_rsh32(klnmf_ctx *ctx);           // rsh32 <- rowsumH (FP32, zero padded)
int launch_colsum_w(klnmf_ctx *ctx, const void *W, const void *Wlo, double *out);   // out[a] += sum_i W[i,a]
int launch_dict_update_t(klnmf_ctx *ctx, const void *Ht_old, void *Ht_new);     // sparse: f x k layout
int launch_decide(klnmf_ctx *ctx, int iter_index);
// hybrid stacks (api.cu: HybridSide): Q <- 0 where X == 0; the joint dictionary update of the dense and the CSR block
int launch_mask_ratio(klnmf_ctx *ctx, void *Q, int64_t ldq, const void *X, int64_t ldx, int64_t rows, int64_t cols, const int *stop,
                      int round_q);
int launch_dict_update_hybrid(klnmf_ctx *ctx, const void *Ht_old, void *Ht_new, const void *Hd_old, void *Hd_new, const void *Nd,
                              int64_t fd, int64_t ldhd, double *total);
int launch_split(klnmf_ctx *ctx, const float *src, float *hi, float *lo, int64_t rows, int64_t cols, int64_t ld);
// dst = (src [+ src_lo]) converted; optional transpose
int launch_convert(klnmf_ctx *ctx, const void *src, const void *src_lo, int src_dtype, int64_t src_ld, void *dst,
                   int dst_es, int64_t dst_ld, int64_t rows, int64_t cols, bool transpose, double scale = 1.0, int scale_f32 = 0);
int launch_zero(klnmf_ctx *ctx, void *p, int64_t bytes);
int launch_fill_uniform(klnmf_ctx *ctx, void *p, int es, int64_t rows, int64_t cols, int64_t ld, uint64_t seed);
int launch_check(klnmf_ctx *ctx, const void *p, int es, int64_t rows, int64_t cols, int64_t ld);
int launch_rowsum_h(klnmf_ctx *ctx);        // rowsumH from the current dictionary
int launch_sum_vals(klnmf_ctx *ctx);        // DS_SUMX = sum of CSR values
int l2_read_bench(int device, int64_t bytes, int iters, double *gbps);   // measurement support

// ---- sparse path: sparse.cu ------------------------------------------------------------------------
// mode: 0 full pass, 1 objective only, 3 SDDMM only.  q_order (mode 0): 0 the ratio is not kept (transform), 1 kept in
// CSR order (the _Q hook), 2 kept in blocked-CSC order for the numerator pass (fit)
// g0: n x ldw, what G starts from; [r0, r0 + rows): a row range (rows < 0: all)
int sparse_rows(klnmf_ctx *ctx, int mode, int q_order = 0, const void *g0 = nullptr, int64_t r0 = 0, int64_t rows = -1);
int sparse_scatter(klnmf_ctx *ctx, bool use_current_w);
int sparse_init_w(klnmf_ctx *ctx, const void *g0 = nullptr, int64_t r0 = 0, int64_t rows = -1);
int sparse_fill_synthetic(klnmf_ctx *ctx, int64_t nnz_per_row, uint64_t seed);
void sparse_release_pattern(klnmf_ctx *ctx);   // drop the blocked-CSC copy (the data changed)

// ---- device-side stacking of mixed dense / CSR modality blocks: stack.cu ------------------------------
int stack_blocks_to_csr(klnmf_ctx *ctx, int n_blocks, const klnmf_block *blocks);   // fills ctx->indptr / indices / vals / nnz

// ---- NCCL through dlopen: nccl_dyn.cu ----------------------------------------------------------------
int nccl_load(const char *path);
int nccl_unique_id(void *id128);
int nccl_comm_init(klnmf_ctx *ctx, const void *id128, int rank, int world);
int nccl_comm_create(void **out, int device, const void *id128, int rank, int world);
int nccl_comm_free(void *comm);
int nccl_group_start();
int nccl_group_end();
int nccl_allreduce_sum(klnmf_ctx *ctx, void *buf, int64_t count, int es);
int nccl_allreduce_sum_f64(klnmf_ctx *ctx, double *buf, int64_t count);
int nccl_allreduce_sum_on(klnmf_ctx *ctx, void *buf, int64_t count, int es, cudaStream_t stream);
void nccl_comm_destroy(klnmf_ctx *ctx);

}  // namespace klnmf
