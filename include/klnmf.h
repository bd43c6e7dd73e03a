/*
 * klnmf.h -- C ABI of the B200-native KL-divergence NMF engine (libklnmf.so).
 *
 * This is the drop-in boundary for ONE hot path of omangin/multimodal: the
 * multiplicative-update loop of `multimodal/lib/nmf.py` (KLdivNMF) as driven by
 * `multimodal/learner.py`.  The reference has no FFI of its own (it is pure
 * Python on numpy/scipy); each entry point below names the reference code it
 * replaces so that a maintainer can bind it with ctypes.
 *
 * Letters follow the reference: X is n x f (samples x features), W is n x k
 * (coefficients, "internal"), H is k x f (dictionary, `components_`).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types;
 *   - every function returns 0 on success, a negative KLNMF_E* code otherwise,
 *     and klnmf_last_error() then holds a human-readable message (thread local);
 *   - host matrices are row-major with a leading dimension given in ELEMENTS;
 *   - "*_host" entry points take host pointers and do the host<->device copies
 *     themselves; "*_device" entry points take device pointers (e.g. the
 *     data_ptr() of a torch CUDA tensor) which are BORROWED, not copied;
 *   - all work is enqueued on the context's stream (klnmf_set_stream);
 *   - there is no CPU fallback: without a CUDA device klnmf_create fails.
 */
#ifndef KLNMF_H_
#define KLNMF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KLNMF_ABI_VERSION 3   /* bumped whenever a symbol or an argument list changes */

/* error codes */
#define KLNMF_OK            0
#define KLNMF_EINVAL       -1   /* bad argument / shape mismatch (reference: AssertionError/ValueError) */
#define KLNMF_ECUDA        -2   /* CUDA runtime / driver error                                          */
#define KLNMF_ENODEVICE    -3   /* no CUDA device: the engine has no CPU path                           */
#define KLNMF_ENCCL        -4   /* NCCL could not be loaded or returned an error                        */
#define KLNMF_ESTATE       -5   /* call sequence error (e.g. iterate before data/dictionary are set)    */
#define KLNMF_ENOMEM       -6

/* arithmetic modes of the dense contractions (the sparse path uses FP32 FMA for
 * the two TF32 modes and FP64 FMA for KLNMF_FP64).  The KL objective is always
 * accumulated in FP64. */
#define KLNMF_MODE_TF32     0   /* tcgen05 kind::tf32, one pass                      */
#define KLNMF_MODE_TF32X3   1   /* tcgen05 split-TF32 (hi*hi + hi*lo + lo*hi)        */
#define KLNMF_MODE_FP64     2   /* DMMA (mma.sync m8n8k4.f64)                        */
#define KLNMF_MODE_TF32R    3   /* one pass on round-to-nearest TF32 copies of W, H, Q-1; FP32 state (hi, lo) */

/* element types of user buffers */
#define KLNMF_F32 0
#define KLNMF_F64 1

typedef struct klnmf_ctx klnmf_ctx;

/* ---- library ------------------------------------------------------------------------ */
int         klnmf_abi_version(void);
const char *klnmf_last_error(void);
int         klnmf_device_count(void);            /* 0 when no usable CUDA device */

/* ---- context ------------------------------------------------------------------------ */
/* One context = one problem shard on one GPU: n_local rows of X and W, the full k x f
 * dictionary.  Replaces the per-call state of KLdivNMF.fit_transform (nmf.py:159-230). */
int klnmf_create(klnmf_ctx **out, int device, int64_t n_local, int64_t f, int64_t k, int mode);
int klnmf_destroy(klnmf_ctx *ctx);
int klnmf_set_stream(klnmf_ctx *ctx, void *cuda_stream);      /* cudaStream_t; NULL = own stream */
/* upper bound (bytes) for the ratio-matrix scratch of the dense path; rows are processed in
 * panels that fit (default 16 GiB). */
int klnmf_set_scratch_limit(klnmf_ctx *ctx, int64_t bytes);

/* ---- data: replaces atleast2d_or_csr + check_non_negative (nmf.py:193-194) -------------
 * Validation (finite, non-negative) is the caller's job on the host side, exactly where the
 * reference does it; klnmf_check_input offers the same test on the device copy. */
int klnmf_set_dense_host(klnmf_ctx *ctx, const void *X, int dtype, int64_t ld);
int klnmf_set_dense_device(klnmf_ctx *ctx, const void *X_dev, int dtype, int64_t ld);
/* The scaled concatenation of modalities, safe_hstack([coef_m * X_m]) of MultimodalLearner.stack_data
 * (learner.py:53-56, array_utils.py:5-9), formed on the device: block b (n x cols[b], row pitch lds[b], host memory)
 * lands in columns [sum(cols[:b]), +cols[b]) of X multiplied by scales[b].  The product is formed in double and
 * rounded once, like numpy's float64 product -- or in float where product_f32[b] != 0, which is what numpy does for a
 * float32 block with a float32 (or Python float) coefficient; product_f32 may be NULL.  sum(cols) must equal f.
 * No stacked copy is ever made on the host. */
int klnmf_set_dense_blocks_host(klnmf_ctx *ctx, int n_blocks, const void *const *X, const int *dtypes,
                                const int64_t *lds, const int64_t *cols, const double *scales, const int *product_f32);
/* The same stack when at least one modality is sparse: the reference then makes the WHOLE stack sparse
 * (scipy.sparse.hstack, array_utils.py:5-9), sparsifying, scaling and concatenating on the host.  Here every block is
 * uploaded as it is and the scaled, stacked CSR matrix is built on the device (count -> scan -> fill); zeros -- explicit
 * ones of a CSR block, any of a dense block -- are dropped, as the reference's eliminate_zeros() drops them from the
 * stack before its first use (nmf.py:66).  CSR blocks: canonical (sorted indices, no duplicates), int64 indptr. */
typedef struct klnmf_block {
  int kind;                 /* 0: dense row-major host array, 1: CSR host arrays                               */
  int dtype;                /* KLNMF_F32 / KLNMF_F64 of the dense array or of the CSR values                    */
  int64_t cols;             /* width of the modality                                                            */
  double scale;             /* its coefficient                                                                  */
  int product_f32;          /* form scale * x in float (numpy: float32 data with a float32 / Python-float coef) */
  const void *dense;        /* kind 0: n x cols, row pitch ld elements                                          */
  int64_t ld;
  const int64_t *indptr;    /* kind 1: n + 1                                                                    */
  const int32_t *indices;   /*         nnz                                                                      */
  const void *values;       /*         nnz                                                                      */
  int64_t nnz;
} klnmf_block;
int klnmf_set_stacked_blocks_host(klnmf_ctx *ctx, int n_blocks, const klnmf_block *blocks);
/* Hybrid stacks (learner.py:53-56 + array_utils.py:5-9): when the dense blocks of a mixed stack have at
 * least `cols` columns together (default 1024; 0 = never), klnmf_set_stacked_blocks_host keeps them DENSE -- they go
 * through the contraction engine (tcgen05 / DMMA), the CSR blocks through the sparse passes, one coefficient matrix and
 * one row normaliser of the dictionary over both.  The results are the reference's for its all-sparse stack: zeros of
 * a dense block carry no ratio.  One-pass arithmetic for the dense block (TF32, TF32R, FP64 modes; TF32X3 keeps the
 * all-CSR stack); anything else builds the CSR stack.  The shards of a
 * multi-GPU fit must all take the same form (the caller decides once: distributed.DeviceGroup).
 * Call before the data.  KLNMF_HYBRID=0 / 1 in the environment overrides: never / whenever possible. */
int klnmf_set_hybrid_min_cols(klnmf_ctx *ctx, int64_t cols);
/* 1 if the context holds a hybrid stack (its dense blocks kept dense), 0 otherwise */
int klnmf_is_hybrid(klnmf_ctx *ctx);
/* CSR with sorted-or-not column indices, no duplicate entries, explicit zeros already
 * removed (the reference calls eliminate_zeros() on the caller's matrix, nmf.py:66). */
int klnmf_set_csr_host(klnmf_ctx *ctx, const int64_t *indptr, const int32_t *indices,
                       const void *values, int dtype, int64_t nnz);
int klnmf_set_csr_device(klnmf_ctx *ctx, const int64_t *indptr_dev, const int32_t *indices_dev,
                         const void *values_dev, int dtype, int64_t nnz);
/* A NEW context over a column subset of a dense parent's data: columns [starts[r], starts[r] + widths[r]) for
 * r < n_ranges, concatenated in that order and gathered ON THE DEVICE (no host copy, no upload); same device, arithmetic
 * mode and rows as the parent, k components.  This is what the evaluation of a multimodal dictionary needs: the
 * coefficients of the SAME test samples observed through every subset of the modalities (experiment.py:350-369 calls
 * learner.reconstruct_internal_multi once per subset, learner.py:71-78, each time re-stacking and re-uploading the data);
 * here the scaled stack of all modalities is uploaded once (klnmf_set_dense_blocks_host) and every subset is a view.
 * The view owns its copy: it stays valid after the parent is destroyed. */
int klnmf_create_column_view(klnmf_ctx *parent, int n_ranges, const int64_t *starts, const int64_t *widths, int64_t k,
                             klnmf_ctx **out);
/* out[0] = 1 if any stored value is negative, out[1] = 1 if any is NaN/Inf */
int klnmf_check_input(klnmf_ctx *ctx, int32_t out[2]);

/* ---- dictionary / coefficients -------------------------------------------------------- */
/* H0 is drawn by the caller from the host numpy RNG so that it is bit-identical with
 * the reference (nmf.py:150-151); float64 k x f row-major. */
int klnmf_set_dictionary_host(klnmf_ctx *ctx, const double *H, int64_t ld);
int klnmf_get_dictionary_host(klnmf_ctx *ctx, double *H, int64_t ld);
/* W0 = X . H0^T (nmf.py:156) */
int klnmf_init_coefficients(klnmf_ctx *ctx);
int klnmf_set_coefficients_host(klnmf_ctx *ctx, const double *W, int64_t ld);
int klnmf_get_coefficients_host(klnmf_ctx *ctx, double *W, int64_t ld);
/* device views of the current state (float32 for the TF32 modes, float64 for FP64);
 * valid until the next klnmf_run / klnmf_destroy.  ld in elements. */
int klnmf_coefficients_device(klnmf_ctx *ctx, void **ptr, int64_t *ld, int *dtype);
int klnmf_dictionary_device(klnmf_ctx *ctx, void **ptr, int64_t *ld, int *dtype);

/* ---- the loop: replaces fit_transform's for-loop (nmf.py:212-222) ---------------------
 * Runs up to max_iter iterations of { e_t = KL(X || W H); stop if prev - e_t < tol_abs;
 * Q = (X+eps)/(WH+eps); W <- W (.) Q H^T; if fit: H <- rownorm(H (.) W_new^T Q) }.
 * eps = 1e-8 (nmf.py:232), row-norm eps = 1e-16 (array_utils.py:19).
 * errors_out (host, max_iter doubles, may be NULL) receives e_1.. ; *n_errors the number
 * recorded (== number of updates applied); *n_iter the reference's n_iter on exit.
 * On a stop the PRE-update W, H are kept, as in the reference.  One host sync at the end
 * (plus one every 32 iterations when tol_abs > 0). */
int klnmf_run(klnmf_ctx *ctx, int max_iter, double tol_abs, int fit,
              double *errors_out, int *n_errors, int *n_iter);
/* The same loop continued: `prev_objective` is the last objective an earlier klnmf_run on this context recorded
 * (+inf for a fresh start, which is what klnmf_run passes), so that a fit split into several calls -- e.g. to write a
 * dictionary checkpoint every so many iterations -- applies the reference's stop test (nmf.py:215) across the call
 * boundary exactly as one long call would. */
int klnmf_run_resume(klnmf_ctx *ctx, int max_iter, double tol_abs, int fit, double prev_objective,
                     double *errors_out, int *n_errors, int *n_iter);
/* KL(X || W H) of the current state: KLdivNMF.error (nmf.py:297-310) */
int klnmf_error(klnmf_ctx *ctx, double *out);

/* ---- the reference's building blocks, kept for API parity (tests/test_nmf_kl.py) ---------- */
/* KLdivNMF._updated_H with Q=None (nmf.py:345-351): dictionary step with the CURRENT W */
int klnmf_dictionary_step(klnmf_ctx *ctx);
/* KLdivNMF._Q (nmf.py:325-336): dense -> n x f (ld in elements); CSR -> nnz values */
int klnmf_ratio_host(klnmf_ctx *ctx, void *out, int dtype, int64_t ld);
/* _special_sparse_dot (nmf.py:52-70): (W.H) at the stored entries of the CSR data */
int klnmf_sddmm_host(klnmf_ctx *ctx, double *out_vals);

/* ---- reconstruction: internal.dot(dico_dest) (learner.py:80-84) ------------------------ */
int klnmf_reconstruct_host(klnmf_ctx *ctx, const double *H_dest, int64_t f_dest, int64_t ld_h,
                           double *out, int64_t ld_out);

/* ---- multi-GPU: rows sharded, one all-reduce of the k x f numerator ---------- */
int klnmf_nccl_load(const char *libnccl_path);              /* dlopen; NULL = default search   */
int klnmf_nccl_unique_id(void *id128);                      /* 128-byte ncclUniqueId           */
int klnmf_comm_init(klnmf_ctx *ctx, const void *id128, int rank, int world);   /* communicator owned by the context */
/* A communicator that outlives contexts: ncclCommInitRank costs 0.2-1 s, a context is created per fit / transform
 * call (nmf.py:159-230 keeps no state between calls either).  klnmf_comm_create is collective over the `world` ranks
 * (processes, or threads of one process with one device each); klnmf_comm_attach lends it to a context, which then
 * all-reduces its numerator and objective partials over it and does NOT destroy it. */
int klnmf_comm_create(void **comm, int device, const void *id128, int rank, int world);
int klnmf_comm_attach(klnmf_ctx *ctx, void *comm, int rank, int world);
int klnmf_comm_destroy(void *comm);

/* ---- measurement support ------------------------------------------------------------------ */
/* synthetic non-negative data generated in place on the device (counter-based hash,
 * uniform in (0,1]); used by bench.py so that 100+ GB inputs need no host copy. */
int klnmf_fill_dense_synthetic(klnmf_ctx *ctx, uint64_t seed);
int klnmf_fill_csr_synthetic(klnmf_ctx *ctx, int64_t nnz_per_row, uint64_t seed);
/* copy the device-resident dense X to / from a host buffer (bench e2e leg) */
int klnmf_get_dense_host(klnmf_ctx *ctx, void *X, int dtype, int64_t ld);
/* counters since create: [0] kernels launched by this library, [1] NCCL calls,
 * [2] bytes copied H2D, [3] bytes copied D2H */
int klnmf_counters(klnmf_ctx *ctx, int64_t out[4]);
/* device time (ms, CUDA events on the context stream) spent in each phase during the last
 * klnmf_run: [0] ratio+objective GEMM, [1] coefficient GEMM, [2] numerator GEMM,
 * [3] dictionary update + normalise, [4] all-reduce, [5] whole run; launches per phase in
 * counts[0..4]. */
int klnmf_last_run_profile(klnmf_ctx *ctx, double ms[6], int64_t counts[5]);
/* ---- evaluation: nearest-example classification on the coefficients ------------------------
 * Replaces evaluation.all_distances / classify_NN / dists_to_found_labels (evaluation.py:68-116) with the
 * measures of metrics.py:58-86.  A is n_test x d, B is n_ex x d (host, row-major float64).  Any of `dists`
 * (n_test x n_ex, ld ldd), `argmin` (n_test; np.argmin semantics: first minimum) and `minval` may be NULL;
 * with dists == NULL the distance matrix is never materialised. */
#define KLNMF_MEASURE_KL           0   /* metrics.kl_div(a, b)       */
#define KLNMF_MEASURE_REV_KL       1   /* metrics.rev_kl_div(a, b)   */
#define KLNMF_MEASURE_SYM_KL       2   /* metrics.sym_kl_div(a, b)   */
#define KLNMF_MEASURE_FROBENIUS    3   /* metrics.frobenius(a, b)    */
#define KLNMF_MEASURE_COSINE_DIFF  4   /* metrics.cosine_diff(a, b)  */
int klnmf_pairwise_host(int device, int measure, int64_t n_test, int64_t n_ex, int64_t d, const double *A,
                        int64_t lda, const double *B, int64_t ldb, double *dists, int64_t ldd, int32_t *argmin,
                        double *minval);

/* diagnostic: one contraction out(M x N) = op(A).op(B) on host float64 buffers through the
 * dense engine of `mode` (the kernels klnmf_run uses).  a_trans=0: A is M x K row-major,
 * a_trans=1: A is stored K x M; b_trans=0: B is K x N row-major, b_trans=1: B is stored N x K. */
int klnmf_contract_host(int device, int mode, int64_t M, int64_t N, int64_t K, const double *A, int a_trans,
                        const double *B, int b_trans, double *out);
/* diagnostic: average device milliseconds of one such contraction on synthetic device-resident operands */
int klnmf_contract_bench(int device, int mode, int64_t M, int64_t N, int64_t K, int a_trans, int b_trans, int iters,
                         double *ms_out);
/* diagnostic: sustained L2 -> SM read bandwidth (GB/s) on a `bytes`-sized buffer (choose it below the L2 capacity) read
 * `iters` times by every SM: the measured denominator of the sparse path's L2 roofline, whose dictionary-column and
 * coefficient-row gathers are served by the L2, not by HBM */
int klnmf_l2_read_bench(int device, int64_t bytes, int iters, double *gbps_out);
/* name of the kernel family that serves the dense contractions in this context
 * ("tcgen05_tf32", "tcgen05_tf32x3", "dmma_f64") -- lets tests assert the native path. */
const char *klnmf_engine_name(klnmf_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* KLNMF_H_ */
