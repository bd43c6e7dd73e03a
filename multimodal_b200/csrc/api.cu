// C ABI of libklnmf (include/klnmf.h): context, data movement, and the device-resident
// iteration loop that replaces KLdivNMF.fit_transform's for-loop (nmf.py:212-222).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace klnmf {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

struct PhaseEvents {
  int phase;
  cudaEvent_t a, b;
};
struct Profiler {
  std::vector<PhaseEvents> ev;
};

inline bool dense_tc(const klnmf_ctx *c) { return c->es == 4 && !c->debug_simt; }

int dense_gemm(klnmf_ctx *ctx, int epi, const GemmDesc &d) {
  if (dense_tc(ctx)) return tc_gemm(ctx, epi, d);
  return generic_gemm(ctx, ctx->es, epi, d);
}

int dmalloc(void **p, int64_t bytes) {
  *p = nullptr;
  if (bytes <= 0) bytes = 16;
  cudaError_t e = cudaMalloc(p, (size_t)bytes);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%lld bytes) failed: %s", (long long)bytes, cudaGetErrorString(e));
    cudaGetLastError();
    return KLNMF_ENOMEM;
  }
  return KLNMF_OK;
}

int ensure_stage(klnmf_ctx *ctx, int64_t bytes) {
  if (ctx->stage_bytes >= bytes) return KLNMF_OK;
  if (ctx->stage) {
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
  }
  KL_TRY(dmalloc(&ctx->stage, bytes));
  ctx->stage_bytes = bytes;
  return KLNMF_OK;
}

constexpr int64_t kStageBytes = (int64_t)256 << 20;

// Allocate W/H/numerator/control buffers once the data kind (dense / CSR) is known.
int ensure_state(klnmf_ctx *ctx) {
  if (ctx->W[0]) return KLNMF_OK;
  const int64_t es = ctx->es;
  ctx->ldw = round_up(ctx->k, 32);
  int64_t h_rows, h_cols;
  if (ctx->sparse) { ctx->ldh = round_up(ctx->k, 32); h_rows = ctx->f; h_cols = ctx->ldh; }
  else { ctx->ldh = round_up(ctx->f, 32); h_rows = ctx->k; h_cols = ctx->ldh; }
  const int64_t wbytes = (ctx->n > 0 ? ctx->n : 1) * ctx->ldw * es, hbytes = h_rows * h_cols * es;
  for (int i = 0; i < 2; i++) {
    KL_TRY(dmalloc(&ctx->W[i], wbytes));
    KL_CUDA(cudaMemsetAsync(ctx->W[i], 0, wbytes, ctx->stream));
    KL_TRY(dmalloc(&ctx->H[i], hbytes));
    KL_CUDA(cudaMemsetAsync(ctx->H[i], 0, hbytes, ctx->stream));
    if (ctx->split) {
      KL_TRY(dmalloc(&ctx->Wlo[i], wbytes));
      KL_CUDA(cudaMemsetAsync(ctx->Wlo[i], 0, wbytes, ctx->stream));
      KL_TRY(dmalloc(&ctx->Hlo[i], hbytes));
      KL_CUDA(cudaMemsetAsync(ctx->Hlo[i], 0, hbytes, ctx->stream));
    }
  }
  ctx->num_bytes = hbytes + (ctx->sparse ? 0 : h_rows * 8 * 32 * es);     // room for the chunked layout's padding (<= 8 chunks)
  KL_TRY(dmalloc(&ctx->num, ctx->num_bytes));
  KL_CUDA(cudaMemsetAsync(ctx->num, 0, ctx->num_bytes, ctx->stream));
  ctx->dred_len = 2 + ctx->ldw + 128;
  KL_TRY(dmalloc((void **)&ctx->dred, ctx->dred_len * 8));
  KL_CUDA(cudaMemsetAsync(ctx->dred, 0, ctx->dred_len * 8, ctx->stream));
  KL_TRY(dmalloc((void **)&ctx->rowsumH, (ctx->k + 1) * 8));
  KL_TRY(dmalloc((void **)&ctx->hsum, (ctx->k + 1) * 8));
  KL_TRY(dmalloc((void **)&ctx->colsumW, (ctx->ldw + 32) * 8));
  KL_CUDA(cudaMemsetAsync(ctx->colsumW, 0, (ctx->ldw + 32) * 8, ctx->stream));
  KL_TRY(dmalloc((void **)&ctx->rsh32, (ctx->ldw + 32) * 4));
  KL_CUDA(cudaMemsetAsync(ctx->rsh32, 0, (ctx->ldw + 32) * 4, ctx->stream));
  KL_CUDA(cudaMemsetAsync(ctx->rowsumH, 0, (ctx->k + 1) * 8, ctx->stream));
  if (!ctx->sparse) {
    // TF32R contracts the rounded ratio only: its low-part panel exists for the parity hooks alone (ensure_qlo)
    const int64_t per_row = round_up(ctx->f, 32) * es * (ctx->split && !ctx->single_pass ? 2 : 1);
    int64_t rows = ctx->scratch_limit / per_row;
    rows = rows / 128 * 128;
    if (rows < 128) rows = 128;
    if (rows > round_up(ctx->n, 128)) rows = round_up(ctx->n > 0 ? ctx->n : 1, 128);
    ctx->panel_rows = rows;
    ctx->ldq = round_up(ctx->f, 32);
    KL_TRY(dmalloc(&ctx->Q, rows * ctx->ldq * es));
    if (ctx->split && !ctx->single_pass) KL_TRY(dmalloc(&ctx->Qlo, rows * ctx->ldq * es));
  }
  return KLNMF_OK;
}

// low parts of the ratio panel, on demand (TF32R: only klnmf_ratio_host, the _Q parity hook, writes them)
int ensure_qlo(klnmf_ctx *ctx) {
  if (ctx->Qlo || !ctx->split || ctx->sparse) return KLNMF_OK;
  return dmalloc(&ctx->Qlo, ctx->panel_rows * ctx->ldq * (int64_t)ctx->es);
}

void release_hybrid(klnmf_ctx *ctx);

void release_data(klnmf_ctx *ctx) {
  if (ctx->x_owned && ctx->X) cudaFree(ctx->X);
  if (ctx->csr_owned) {
    if (ctx->indptr) cudaFree(ctx->indptr);
    if (ctx->indices) cudaFree(ctx->indices);
    if (ctx->vals) cudaFree(ctx->vals);
  }
  if (ctx->qnz) cudaFree(ctx->qnz);
  sparse_release_pattern(ctx);
  ctx->X = nullptr; ctx->indptr = nullptr; ctx->indices = nullptr; ctx->vals = nullptr; ctx->qnz = nullptr;
  ctx->x_owned = ctx->csr_owned = false;
  ctx->have_x = false;
}

int kind_guard(klnmf_ctx *ctx, bool sparse) {
  KL_CHECK(!ctx->hyb, KLNMF_ESTATE, "a context that holds a hybrid (dense + CSR) stack keeps it for life; create a new one");
  KL_CHECK(!ctx->W[0] || ctx->sparse == sparse, KLNMF_ESTATE,
           "a context serves either dense or CSR data for its whole life; create a new one");
  ctx->sparse = sparse;
  if (sparse) ctx->split = false;   // the CSR path is FP32 FMA: no (hi, lo) operand pairs
  return KLNMF_OK;
}

// host (row-major, ld elements, dtype) -> device (es, dst_ld), optional transpose
int upload_matrix(klnmf_ctx *ctx, const void *src, int dtype, int64_t ld, void *dst, int64_t dst_ld, int64_t rows,
                  int64_t cols, bool transpose, double scale = 1.0, int scale_f32 = 0) {
  if (rows <= 0 || cols <= 0) return KLNMF_OK;
  const int64_t ses = dtype == KLNMF_F64 ? 8 : 4;
  if (!transpose && ses == ctx->es) {
    KL_CUDA(cudaMemcpy2DAsync(dst, dst_ld * ses, src, ld * ses, cols * ses, rows, cudaMemcpyHostToDevice, ctx->stream));
    ctx->bytes_h2d += rows * cols * ses;
    return launch_scale_block(ctx, dst, dst_ld, rows, cols, scale, scale_f32);     // in place
  }
  int64_t chunk = kStageBytes / (cols * ses);
  if (chunk < 1) chunk = 1;
  KL_TRY(ensure_stage(ctx, (chunk < rows ? chunk : rows) * cols * ses));
  for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
    const int64_t r = rows - r0 < chunk ? rows - r0 : chunk;
    KL_CUDA(cudaMemcpy2DAsync(ctx->stage, cols * ses, (const char *)src + r0 * ld * ses, ld * ses, cols * ses, r,
                              cudaMemcpyHostToDevice, ctx->stream));
    void *d = transpose ? (void *)((char *)dst + r0 * ctx->es) : (void *)((char *)dst + r0 * dst_ld * ctx->es);
    KL_TRY(launch_convert(ctx, ctx->stage, nullptr, dtype, cols, d, ctx->es, dst_ld, r, cols, transpose, scale, scale_f32));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));   // staging buffer is reused
    ctx->bytes_h2d += r * cols * ses;
  }
  return KLNMF_OK;
}

// Large device -> pageable-host downloads (the coefficients of a big shard: GBs).  A plain cudaMemcpy into
// pageable memory moved 4 GB/s (measured, tools/e2e_breakdown.py: 52 % of an end-to-end fit_transform call);
// here the panel is converted on the device, lands in one of two pinned staging buffers at PCIe speed, and a
// few host threads copy it into the caller's array (first-touch page faults included) while the next panel is
// already on its way.
int download_large(klnmf_ctx *ctx, const void *src, const void *src_lo, int64_t src_ld, void *dst, int dtype, int64_t ld,
                   int64_t rows, int64_t cols) {
  const int64_t des = dtype == KLNMF_F64 ? 8 : 4;
  const int src_dtype = ctx->es == 8 ? KLNMF_F64 : KLNMF_F32;
  constexpr int64_t kPin = (int64_t)64 << 20;
  // the two pinned buffers are process-wide (pinning / unpinning 128 MB costs ~0.3 s per context otherwise) and
  // held for the duration of one download
  // (one pair per DEVICE: the shards of a multi-GPU call download concurrently, one host thread each)
  static std::mutex pin_mutex[64];
  static void *g_pin[64][2] = {};
  const int slot = ctx->device & 63;
  std::lock_guard<std::mutex> pin_lock(pin_mutex[slot]);
  for (int b = 0; b < 2; b++) {
    if (!g_pin[slot][b]) KL_CUDA(cudaHostAlloc(&g_pin[slot][b], kPin, cudaHostAllocPortable));
    ctx->pin_stage[b] = g_pin[slot][b];
  }
  int64_t chunk = kPin / (cols * des);
  if (chunk < 1) chunk = 1;
  KL_CHECK(chunk * cols * des <= kPin, KLNMF_EINVAL, "download: one row exceeds the staging buffer");
  KL_TRY(ensure_stage(ctx, 2 * chunk * cols * des));
  cudaEvent_t ev[2];
  KL_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  KL_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
  auto host_copy = [&](int b, int64_t r0, int64_t r) {
    const char *ps = (const char *)ctx->pin_stage[b];
    std::vector<std::thread> th;
    const int64_t per = (r + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
      const int64_t a = t * per, e = a + per < r ? a + per : r;
      if (a >= e) break;
      th.emplace_back([=]() {
        if (ld == cols) memcpy((char *)dst + (r0 + a) * ld * des, ps + a * cols * des, (size_t)((e - a) * cols * des));
        else
          for (int64_t i = a; i < e; i++) memcpy((char *)dst + (r0 + i) * ld * des, ps + i * cols * des, (size_t)(cols * des));
      });
    }
    for (auto &x : th) x.join();
  };
  int rc = KLNMF_OK;
  int64_t prev_r0 = -1, prev_r = 0;
  int b = 0;
  for (int64_t r0 = 0; r0 < rows && rc == KLNMF_OK; r0 += chunk, b ^= 1) {
    const int64_t r = rows - r0 < chunk ? rows - r0 : chunk;
    const void *s = (const char *)src + r0 * src_ld * ctx->es;
    const void *sl = src_lo ? (const void *)((const char *)src_lo + r0 * src_ld * ctx->es) : nullptr;
    void *dstage = (char *)ctx->stage + b * chunk * cols * des;
    rc = launch_convert(ctx, s, sl, src_dtype, src_ld, dstage, (int)des, cols, r, cols, false);
    if (rc == KLNMF_OK && cudaMemcpyAsync(ctx->pin_stage[b], dstage, (size_t)(r * cols * des), cudaMemcpyDeviceToHost,
                                          ctx->stream) != cudaSuccess) {
      set_error("download: cudaMemcpyAsync failed");
      rc = KLNMF_ECUDA;
    }
    cudaEventRecord(ev[b], ctx->stream);
    if (prev_r0 >= 0) {                         // the previous panel is (being) copied out while this one is in flight
      cudaEventSynchronize(ev[b ^ 1]);
      host_copy(b ^ 1, prev_r0, prev_r);
    }
    prev_r0 = r0; prev_r = r;
    ctx->bytes_d2h += r * cols * des;
  }
  if (rc == KLNMF_OK && prev_r0 >= 0) {
    if (cudaEventSynchronize(ev[b ^ 1]) != cudaSuccess) { set_error("download: device error"); rc = KLNMF_ECUDA; }
    else host_copy(b ^ 1, prev_r0, prev_r);
  }
  cudaStreamSynchronize(ctx->stream);
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  return rc;
}

// device (es [+lo], src_ld) -> host double/float (ld), optional transpose (src is cols x rows then)
int download_matrix(klnmf_ctx *ctx, const void *src, const void *src_lo, int64_t src_ld, void *dst, int dtype,
                    int64_t ld, int64_t rows, int64_t cols, bool transposed_src) {
  if (rows <= 0 || cols <= 0) return KLNMF_OK;
  const int64_t des = dtype == KLNMF_F64 ? 8 : 4;
  const int src_dtype = ctx->es == 8 ? KLNMF_F64 : KLNMF_F32;
  if (!transposed_src && !src_lo && des == ctx->es) {
    KL_CUDA(cudaMemcpy2DAsync(dst, ld * des, src, src_ld * des, cols * des, rows, cudaMemcpyDeviceToHost, ctx->stream));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->bytes_d2h += rows * cols * des;
    return KLNMF_OK;
  }
  if (transposed_src) {
    // src is (cols x rows) with leading dimension src_ld; produce rows x cols on the host
    KL_TRY(ensure_stage(ctx, rows * cols * des));
    KL_TRY(launch_convert(ctx, src, src_lo, src_dtype, src_ld, ctx->stage, (int)des, cols, cols, rows, true));
    KL_CUDA(cudaMemcpy2DAsync(dst, ld * des, ctx->stage, cols * des, cols * des, rows, cudaMemcpyDeviceToHost, ctx->stream));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->bytes_d2h += rows * cols * des;
    return KLNMF_OK;
  }
  if (rows * cols * des >= ((int64_t)32 << 20)) return download_large(ctx, src, src_lo, src_ld, dst, dtype, ld, rows, cols);
  int64_t chunk = kStageBytes / (cols * des);
  if (chunk < 1) chunk = 1;
  KL_TRY(ensure_stage(ctx, (chunk < rows ? chunk : rows) * cols * des));
  for (int64_t r0 = 0; r0 < rows; r0 += chunk) {
    const int64_t r = rows - r0 < chunk ? rows - r0 : chunk;
    const void *s = (const char *)src + r0 * src_ld * ctx->es;
    const void *sl = src_lo ? (const void *)((const char *)src_lo + r0 * src_ld * ctx->es) : nullptr;
    KL_TRY(launch_convert(ctx, s, sl, src_dtype, src_ld, ctx->stage, (int)des, cols, r, cols, false));
    KL_CUDA(cudaMemcpy2DAsync((char *)dst + r0 * ld * des, ld * des, ctx->stage, cols * des, cols * des, r,
                              cudaMemcpyDeviceToHost, ctx->stream));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->bytes_d2h += r * cols * des;
  }
  return KLNMF_OK;
}

struct PhaseTimer {
  klnmf_ctx *ctx;
  Profiler *prof;
  int phase;
  cudaEvent_t a = nullptr, b = nullptr;
  PhaseTimer(klnmf_ctx *c, Profiler *p, int ph) : ctx(c), prof(p), phase(ph) {
    if (prof) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, ctx->stream);
    }
  }
  ~PhaseTimer() {
    if (prof) {
      cudaEventRecord(b, ctx->stream);
      prof->ev.push_back({phase, a, b});
    }
  }
};

// elements of the numerator buffer that hold data (what a whole-buffer all-reduce covers)
int64_t num_elems(const klnmf_ctx *ctx) {
  if (ctx->sparse) return ctx->f * ctx->ldh;
  return ctx->num_chunks > 1 ? (int64_t)ctx->num_chunks * ctx->k * ctx->num_cc : ctx->k * ctx->ldh;
}

// Dense fit on several ranks: keep the numerator in contiguous f-chunks so that finished chunks can be all-reduced while
// later ones are contracted (dense_iteration).  KLNMF_AR_CHUNKS (1..8, default 4; 1 = one all-reduce after the
// numerator) and KLNMF_AR_CHUNK_MIN_F (default 2048: narrower dictionaries reduce in one piece) tune it.
int setup_num_chunks(klnmf_ctx *ctx, int fit) {
  int nch = 1;
  if (fit && !ctx->sparse && ctx->world > 1) {
    const char *e = getenv("KLNMF_AR_CHUNKS");
    nch = e ? atoi(e) : 4;
    if (nch < 1) nch = 1;
    if (nch > 8) nch = 8;
    const char *m = getenv("KLNMF_AR_CHUNK_MIN_F");
    const int64_t min_f = m ? atoll(m) : 2048;
    if (ctx->f < min_f || ctx->f < 32 * nch) nch = 1;
  }
  int64_t cc = 0;
  if (nch > 1) {
    cc = round_up(ceil_div(ctx->ldh, nch), 32);
    if ((int64_t)nch * ctx->k * cc * ctx->es > ctx->num_bytes) { nch = 1; cc = 0; }
  }
  if (nch > 1 && !ctx->comm_stream) {
    KL_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (int c = 0; c < 8; c++) KL_CUDA(cudaEventCreateWithFlags(&ctx->comm_ev[c], cudaEventDisableTiming));
    KL_CUDA(cudaEventCreateWithFlags(&ctx->comm_done, cudaEventDisableTiming));
  }
  ctx->num_chunks = nch;
  ctx->num_cc = cc;
  return KLNMF_OK;
}

// Local (per-rank) work of one iteration on the dense path: three contractions per row panel.
int dense_iteration(klnmf_ctx *ctx, int fit, bool only_error, Profiler *prof, bool dict_only = false,
                    void *ratio_host = nullptr, int ratio_dtype = KLNMF_F64, int64_t ratio_ld = 0) {
  const int64_t es = ctx->es;
  const int cur = ctx->cur, hc = ctx->hcur;
  const int *stop = only_error ? nullptr : ctx->flags + FL_STOP;
  const bool fused = !only_error && !dict_only && !ratio_host && fused_supported(ctx, fit) && ctx->n > 0;
  // Centered ratio (split-TF32 mode, the iterations of klnmf_run): the panel holds Q - 1 instead of Q, so that the long
  // contractions sum terms of both signs around zero instead of non-negative terms -- the tensor core accumulates FP32
  // with truncation, which otherwise costs ~1e-8 f of relative error on W (8e-5 at f = 8192).  The
  // missing parts are exact and cheap: G = (Q-1).H^T + rowsum(H) in the coefficient epilogue, N = W'^T.(Q-1) +
  // colsum(W') in the dictionary update.  The parity hooks (_Q, _updated_H) and klnmf_error keep the plain form, and so
  // does the one-pass tf32 mode: there the tensor core also truncates the operands themselves, a bias that cancels
  // between S = W.H and G = Q.H^T in the plain form and would not in the centered one (measured 3e-5 -> 7e-4 on W).
  const bool centered = dense_tc(ctx) && ctx->split && !only_error && !dict_only && !ratio_host &&
                        !(getenv("KLNMF_CENTER") && atoi(getenv("KLNMF_CENTER")) == 0);
  ctx->centered = centered;
  const float qshift = centered ? 1.f : 0.f;
  const float *colbias = centered ? ctx->rsh32 : nullptr;
  // TF32R: one MMA per step on the round-to-nearest TF32 halves of W, H and Q - 1 (the _Q parity hook and the
  // stand-alone objective, KLdivNMF.error, keep the split form: they are not on the loop)
  const int single = ctx->single_pass && !ratio_host && !only_error ? 1 : 0;
  if (ratio_host) KL_TRY(ensure_qlo(ctx));
  if (centered) KL_TRY(launch_rsh32(ctx));
  if (fused) {
    // k <= 128 (fit, transform) / k <= 256 (transform): the coefficient half-step is one fused kernel, the ratio
    // never leaves the SM (dense_fused.cu, dense_fused256.cu)
    if (!ctx->Ht) {
      ctx->ldht = ctx->ldw;
      KL_TRY(dmalloc(&ctx->Ht, ctx->f * ctx->ldht * es));
      KL_CUDA(cudaMemsetAsync(ctx->Ht, 0, ctx->f * ctx->ldht * es, ctx->stream));
      ctx->ht_of = -1;
    }
    if (ctx->ht_of != hc || ctx->ht_stale) {
      KL_TRY(launch_convert(ctx, ctx->H[hc], nullptr, KLNMF_F32, ctx->ldh, ctx->Ht, 4, ctx->ldht, ctx->k, ctx->f, true));
      ctx->ht_of = hc;
      ctx->ht_stale = false;
    }
  }
  if (fused && !fit) {
    PhaseTimer t(ctx, prof, PH_RATIO);
    FusedDesc d{};
    d.M = ctx->n; d.F = ctx->f; d.K = ctx->k;
    d.W = ctx->W[cur]; d.ldw = ctx->ldw;
    d.H = ctx->H[hc]; d.ldh = ctx->ldh;
    d.Ht = ctx->Ht; d.ldht = ctx->ldht;
    d.X = ctx->X; d.ldx = ctx->ldx;
    d.Wout = ctx->W[cur ^ 1]; d.ldwo = ctx->ldw;
    d.kl = ctx->dred; d.stop = stop;
    d.qshift = qshift; d.colbias = colbias;
    d.Wlo = ctx->split ? ctx->Wlo[cur] : nullptr; d.Wout_lo = ctx->split ? ctx->Wlo[cur ^ 1] : nullptr;
    d.accurate = ctx->single_pass ? 1 : 0;
    return fused_coef_step(ctx, d);
  }
  for (int64_t r0 = 0; r0 < ctx->n; r0 += ctx->panel_rows) {
    const int64_t rows = ctx->n - r0 < ctx->panel_rows ? ctx->n - r0 : ctx->panel_rows;
    const char *Wc = (const char *)ctx->W[cur] + r0 * ctx->ldw * es;
    const char *Wclo = ctx->split ? (const char *)ctx->Wlo[cur] + r0 * ctx->ldw * es : nullptr;
    char *Wn = (char *)ctx->W[cur ^ 1] + r0 * ctx->ldw * es;
    char *Wnlo = ctx->split ? (char *)ctx->Wlo[cur ^ 1] + r0 * ctx->ldw * es : nullptr;
    if (fused) {   // fit, k <= 128: ratio + objective + coefficient update in one kernel, Q written once for the numerator
      PhaseTimer t(ctx, prof, PH_RATIO);
      FusedDesc d{};
      d.M = rows; d.F = ctx->f; d.K = ctx->k;
      d.W = Wc; d.ldw = ctx->ldw;
      d.H = ctx->H[hc]; d.ldh = ctx->ldh;
      d.Ht = ctx->Ht; d.ldht = ctx->ldht;
      d.X = (const char *)ctx->X + r0 * ctx->ldx * es; d.ldx = ctx->ldx;
      d.Wout = Wn; d.ldwo = ctx->ldw;
      d.Q = ctx->Q; d.ldq = ctx->ldq;
      d.kl = ctx->dred; d.stop = stop;
      d.qshift = qshift; d.colbias = colbias;
      d.Wlo = Wclo; d.Wout_lo = Wnlo;
      d.accurate = ctx->single_pass ? 1 : 0;
      KL_TRY(fused_coef_step(ctx, d));
    } else {  // ratio + objective: Q = (X+eps)/(W.H+eps)   (nmf.py:325-336, metrics.py:18-20)
      PhaseTimer t(ctx, prof, PH_RATIO);
      GemmDesc d{};
      d.M = rows; d.N = ctx->f; d.K = ctx->k;
      d.A = Wc; d.a_sm = ctx->ldw; d.a_sk = 1; d.A_lo = Wclo;
      d.B = ctx->H[hc]; d.b_sk = ctx->ldh; d.b_sn = 1; d.B_lo = ctx->split ? ctx->Hlo[hc] : nullptr;
      d.out = ctx->Q; d.ldo = ctx->ldq; d.out_lo = ctx->Qlo;
      d.aux = (const char *)ctx->X + r0 * ctx->ldx * es; d.ldaux = ctx->ldx;
      d.kl = ctx->dred; d.stop = stop; d.only_kl = only_error ? 1 : 0;
      d.qshift = qshift;
      d.single_pass = single; d.round_out = single;
      KL_TRY(dense_gemm(ctx, EPI_RATIO, d));
    }
    if (ratio_host) {   // _Q parity hook: hand the ratio panel back to the host
      KL_TRY(download_matrix(ctx, ctx->Q, ctx->Qlo, ctx->ldq,
                             (char *)ratio_host + r0 * ratio_ld * (ratio_dtype == KLNMF_F64 ? 8 : 4), ratio_dtype,
                             ratio_ld, rows, ctx->f, false));
      continue;
    }
    if (only_error) continue;
    if (!dict_only && !fused) {  // coefficients: W' = W (.) (Q.H^T)           (nmf.py:338-343)
      PhaseTimer t(ctx, prof, PH_COEF);
      GemmDesc d{};
      d.M = rows; d.N = ctx->k; d.K = ctx->f;
      d.A = ctx->Q; d.a_sm = ctx->ldq; d.a_sk = 1; d.A_lo = ctx->Qlo;
      d.B = ctx->H[hc]; d.b_sk = 1; d.b_sn = ctx->ldh; d.B_lo = ctx->split ? ctx->Hlo[hc] : nullptr;
      d.out = Wn; d.ldo = ctx->ldw; d.out_lo = Wnlo;
      d.aux = Wc; d.ldaux = ctx->ldw; d.aux_lo = Wclo;
      d.stop = stop;
      d.colbias = colbias;
      d.single_pass = single;
      KL_TRY(dense_gemm(ctx, EPI_MULW, d));
    }
    if (fit) {  // dictionary numerator: N += W'^T.Q  (stale Q, new W: nmf.py:345-349)
      PhaseTimer t(ctx, prof, PH_NUM);
      // Several ranks: the numerator is contracted f-chunk by f-chunk (contiguous chunk buffers), and on the LAST row
      // panel every finished chunk is all-reduced on a side stream while the next one is still being contracted
      // (SURVEY 8e "overlap"): only the last chunk's reduction is exposed.
      const bool last_panel = r0 + ctx->panel_rows >= ctx->n;
      const int nch = ctx->num_chunks;
      const int64_t cc = nch > 1 ? ctx->num_cc : ctx->ldh;
      for (int c = 0; c < nch; c++) {
        const int64_t j0 = (int64_t)c * cc;
        const int64_t ncols = nch > 1 ? (ctx->f - j0 < cc ? ctx->f - j0 : cc) : ctx->f;
        char *out_c = (char *)ctx->num + (nch > 1 ? (int64_t)c * ctx->k * cc * es : 0);
        if (ncols > 0) {
          GemmDesc d{};
          d.M = ctx->k; d.N = ncols; d.K = rows;
          d.A = dict_only ? (const void *)Wc : (const void *)Wn; d.a_sm = 1; d.a_sk = ctx->ldw;
          d.A_lo = dict_only ? (const void *)Wclo : (const void *)Wnlo;
          d.B = (const char *)ctx->Q + j0 * es; d.b_sk = ctx->ldq; d.b_sn = 1;
          d.B_lo = ctx->Qlo ? (const char *)ctx->Qlo + j0 * es : nullptr;
          d.out = out_c; d.ldo = cc;
          d.stop = stop;
          d.single_pass = single;
          KL_TRY(dense_gemm(ctx, EPI_ACC, d));
        }
        if (nch > 1 && last_panel && ctx->world > 1) {
          KL_CUDA(cudaEventRecord(ctx->comm_ev[c], ctx->stream));
          KL_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[c], 0));
          KL_TRY(nccl_allreduce_sum_on(ctx, out_c, ctx->k * cc, (int)es, ctx->comm_stream));
          if (c == nch - 1) {
            KL_CUDA(cudaEventRecord(ctx->comm_done, ctx->comm_stream));
            ctx->num_reduced = true;
          }
        }
      }
    }
  }
  if (fit && ctx->n == 0 && ctx->num_chunks > 1 && ctx->world > 1) {
    // a shard without rows still takes part in every chunk's all-reduce (the same NCCL call sequence on every rank)
    const int64_t cc = ctx->num_cc;
    for (int c = 0; c < ctx->num_chunks; c++) {
      KL_CUDA(cudaEventRecord(ctx->comm_ev[c], ctx->stream));
      KL_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[c], 0));
      KL_TRY(nccl_allreduce_sum_on(ctx, (char *)ctx->num + (int64_t)c * ctx->k * cc * es, ctx->k * cc, (int)es, ctx->comm_stream));
    }
    KL_CUDA(cudaEventRecord(ctx->comm_done, ctx->comm_stream));
    ctx->num_reduced = true;
  }
  // colsum(W') of this rank joins the all-reduced doubles (slots the sparse objective uses for colsum(W))
  if (fit && centered) KL_TRY(launch_colsum_w(ctx, ctx->W[cur ^ 1], ctx->split ? ctx->Wlo[cur ^ 1] : nullptr, ctx->dred + 2));
  return KLNMF_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Hybrid stacks (learner.py:53-56): a dense modality next to a CSR one.  The reference makes the whole stack
// sparse (array_utils.py:5-9).  Here the CSR blocks form the context's CSR data (ctx->f counts THEIR columns) and go
// through sparse.cu, the dense blocks stay dense (HybridSide) and go through the contraction engine:
//   S_d = W.H_d -> Q_d = (X_d+eps)/(S_d+eps) with Q_d = 0 where X_d = 0 (structural zeros of the reference's stack carry
//   no ratio), objective terms in the dense form -- x log q - x + s is exactly the sparse form's share of the block:
//   the log term vanishes at x = 0 and sum(s) is the block's colsum(W).rowsum(H_d);
//   G_d = Q_d.H_d^T (stored), the rows pass starts its G from it: W' = W (.) (G_d + G_s);
//   N_d = W'^T.Q_d next to the CSR block's numerator, ONE normaliser per component over both (array_utils.py:19-22).
// The dense block multiplies in the mode's one-pass form (TF32 / TF32R: tcgen05 kind::tf32 on plain FP32 operands; FP64:
// DMMA); TF32X3 keeps the all-CSR stack (its FP32-grade promise).  One ratio panel: n x fd must fit the scratch limit.
// ------------------------------------------------------------------------------------------------------------------
struct HybridSide {
  int64_t f_total = 0, fd = 0, ld = 0;       // ld = fd rounded up to 32: pitch of X, Q, H, num
  int64_t panel_rows = 0;                    // rows of the ratio panel Q (the samples are walked in panels of that many)
  void *X = nullptr, *Q = nullptr, *G = nullptr, *num = nullptr;
  void *H[2] = {nullptr, nullptr};
  // FP32 modes: round-to-nearest TF32 copies of what the dense block's contractions multiply -- the dictionary part
  // (k x ld) and the panel's rows of W / W' (panel_rows x ldw); the ratio panel is rounded where it is masked.  The
  // tensor core would TRUNCATE its operands (a bias of 2^-12 per operand that short contractions do not average out:
  // 1.4e-3 on the dictionary of the golden learner after 20 iterations, 3.6e-4 with the rounded copies).
  void *Hr = nullptr, *Wr = nullptr;
  double *total = nullptr;                   // k doubles: the joint normaliser
  struct Range { int dense; int64_t col0, cols, off; };
  std::vector<Range> ranges;                 // the stack's column ranges, in stack order; off = first column inside its part
};

void release_hybrid(klnmf_ctx *ctx) {
  HybridSide *hy = (HybridSide *)ctx->hyb;
  if (!hy) return;
  void *ptrs[] = {hy->X, hy->Q, hy->G, hy->num, hy->H[0], hy->H[1], hy->total, hy->Hr, hy->Wr};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  delete hy;
  ctx->hyb = nullptr;
}

// Rows [r0, r0 + rows) (one panel): ratio + objective of the dense block, its ratio masked at the structural zeros, and
// (unless only the objective is wanted) G_d = Q_d.H_d^T into the rows of hy->G
int hybrid_dense_half(klnmf_ctx *ctx, bool only_error, int64_t r0, int64_t rows) {
  HybridSide *hy = (HybridSide *)ctx->hyb;
  const int cur = ctx->cur, hc = ctx->hcur;
  const int64_t es = ctx->es;
  const int *stop = only_error ? nullptr : ctx->flags + FL_STOP;
  const char *Xp = (const char *)hy->X + r0 * hy->ld * es;
  const void *Wp = (const char *)ctx->W[cur] + r0 * ctx->ldw * es, *Hd = hy->H[hc];
  if (hy->Hr) {   // (hy->Hr was refreshed by the caller: once per pass over the panels)
    KL_TRY(launch_split(ctx, (const float *)Wp, (float *)hy->Wr, nullptr, rows, ctx->ldw, ctx->ldw));
    Wp = hy->Wr; Hd = hy->Hr;
  }
  GemmDesc d{};
  d.M = rows; d.N = hy->fd; d.K = ctx->k;
  d.A = Wp; d.a_sm = ctx->ldw; d.a_sk = 1;
  d.B = Hd; d.b_sk = hy->ld; d.b_sn = 1;
  d.out = hy->Q; d.ldo = hy->ld;
  d.aux = Xp; d.ldaux = hy->ld;
  d.kl = ctx->dred; d.stop = stop; d.only_kl = only_error ? 1 : 0;
  KL_TRY(dense_gemm(ctx, EPI_RATIO, d));
  if (only_error) return KLNMF_OK;
  KL_TRY(launch_mask_ratio(ctx, hy->Q, hy->ld, Xp, hy->ld, rows, hy->fd, stop, hy->Hr ? 1 : 0));
  GemmDesc c{};
  c.M = rows; c.N = ctx->k; c.K = hy->fd;
  c.A = hy->Q; c.a_sm = hy->ld; c.a_sk = 1;
  c.B = Hd; c.b_sk = 1; c.b_sn = hy->ld;
  c.out = (char *)hy->G + r0 * ctx->ldw * es; c.ldo = ctx->ldw; c.stop = stop;
  return dense_gemm(ctx, EPI_STORE, c);
}
// N_d += W'^T.Q_d over the panel's rows, with the updated coefficients (stale ratio, new W: nmf.py:345-349)
int hybrid_dense_numerator(klnmf_ctx *ctx, int64_t r0, int64_t rows) {
  HybridSide *hy = (HybridSide *)ctx->hyb;        // (hy->num was zeroed at the top of the iteration)
  GemmDesc m{};
  m.M = ctx->k; m.N = hy->fd; m.K = rows;
  const void *Wn = (const char *)ctx->W[ctx->cur ^ 1] + r0 * ctx->ldw * (int64_t)ctx->es;
  if (hy->Wr) {
    KL_TRY(launch_split(ctx, (const float *)Wn, (float *)hy->Wr, nullptr, rows, ctx->ldw, ctx->ldw));
    Wn = hy->Wr;
  }
  m.A = Wn; m.a_sm = 1; m.a_sk = ctx->ldw;
  m.B = hy->Q; m.b_sk = hy->ld; m.b_sn = 1;
  m.out = hy->num; m.ldo = hy->ld; m.stop = ctx->flags + FL_STOP;
  return dense_gemm(ctx, EPI_ACC, m);
}
// The coefficient half-step of a hybrid stack (and, fitting, the dense block's numerator), panel by panel: the ratio
// panel of the dense block lives for one panel only, so its numerator is taken before the next panel overwrites it.
int hybrid_round_dictionary(klnmf_ctx *ctx) {
  HybridSide *hy = (HybridSide *)ctx->hyb;
  if (!hy->Hr) return KLNMF_OK;
  return launch_split(ctx, (const float *)hy->H[ctx->hcur], (float *)hy->Hr, nullptr, ctx->k, hy->ld, hy->ld);
}

int hybrid_rows_passes(klnmf_ctx *ctx, int fit, Profiler *prof) {
  HybridSide *hy = (HybridSide *)ctx->hyb;
  KL_TRY(hybrid_round_dictionary(ctx));
  for (int64_t r0 = 0; r0 < ctx->n; r0 += hy->panel_rows) {
    const int64_t rows = ctx->n - r0 < hy->panel_rows ? ctx->n - r0 : hy->panel_rows;
    {
      PhaseTimer t(ctx, prof, PH_RATIO);
      KL_TRY(hybrid_dense_half(ctx, false, r0, rows));
      KL_TRY(sparse_rows(ctx, 0, fit ? 2 : 0, hy->G, r0, rows));
    }
    if (fit) {
      PhaseTimer t(ctx, prof, PH_NUM);
      KL_TRY(hybrid_dense_numerator(ctx, r0, rows));
    }
  }
  return KLNMF_OK;
}

int reset_reduction(klnmf_ctx *ctx) {
  KL_CUDA(cudaMemsetAsync(ctx->dred, 0, ctx->dred_len * 8, ctx->stream));
  KL_CUDA(cudaMemcpyAsync(ctx->dred + 1, ctx->dscal + DS_SUMX, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  return KLNMF_OK;
}

}  // namespace
}  // namespace klnmf

using namespace klnmf;

extern "C" {

int klnmf_abi_version(void) { return KLNMF_ABI_VERSION; }

const char *klnmf_last_error(void) { return g_err; }

int klnmf_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int klnmf_create(klnmf_ctx **out, int device, int64_t n_local, int64_t f, int64_t k, int mode) {
  KL_CHECK(out != nullptr, KLNMF_EINVAL, "klnmf_create: out is NULL");
  *out = nullptr;
  KL_CHECK(n_local >= 0 && f > 0 && k > 0, KLNMF_EINVAL, "klnmf_create: bad shape n=%lld f=%lld k=%lld",
           (long long)n_local, (long long)f, (long long)k);
  KL_CHECK(mode == KLNMF_MODE_TF32 || mode == KLNMF_MODE_TF32X3 || mode == KLNMF_MODE_FP64 || mode == KLNMF_MODE_TF32R, KLNMF_EINVAL,
           "klnmf_create: unknown mode %d", mode);
  const int ndev = klnmf_device_count();
  KL_CHECK(ndev > 0, KLNMF_ENODEVICE, "no CUDA device: libklnmf has no CPU path");
  KL_CHECK(device >= 0 && device < ndev, KLNMF_EINVAL, "device %d out of range (%d devices)", device, ndev);
  KL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KL_CUDA(cudaGetDeviceProperties(&prop, device));
  KL_CHECK(prop.major == 10, KLNMF_ENODEVICE, "device %d is sm_%d%d; libklnmf is built for sm_100a only", device,
           prop.major, prop.minor);
  klnmf_ctx *ctx = new klnmf_ctx();
  ctx->device = device;
  ctx->n = n_local; ctx->f = f; ctx->k = k; ctx->mode = mode;
  ctx->es = mode == KLNMF_MODE_FP64 ? 8 : 4;
  ctx->sm_count = prop.multiProcessorCount;
  const char *dbg = getenv("KLNMF_DEBUG_ENGINE");
  ctx->debug_simt = dbg && strcmp(dbg, "simt") == 0;
  ctx->split = (mode == KLNMF_MODE_TF32X3 || mode == KLNMF_MODE_TF32R) && !ctx->debug_simt;
  ctx->single_pass = mode == KLNMF_MODE_TF32R && !ctx->debug_simt;
  const char *pf = getenv("KLNMF_PROFILE");
  ctx->profile = pf && pf[0] == '1';
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    delete ctx;
    return KLNMF_ECUDA;
  }
  ctx->own_stream = true;
  int r = dmalloc((void **)&ctx->dscal, DS_COUNT * 8);
  if (r == KLNMF_OK) r = dmalloc((void **)&ctx->flags, FL_COUNT * 4);
  if (r != KLNMF_OK || cudaMallocHost((void **)&ctx->pinned, 4096) != cudaSuccess) {
    klnmf_destroy(ctx);
    return r != KLNMF_OK ? r : KLNMF_ECUDA;
  }
  cudaMemsetAsync(ctx->dscal, 0, DS_COUNT * 8, ctx->stream);
  cudaMemsetAsync(ctx->flags, 0, FL_COUNT * 4, ctx->stream);
  *out = ctx;
  return KLNMF_OK;
}

int klnmf_destroy(klnmf_ctx *ctx) {
  if (!ctx) return KLNMF_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  nccl_comm_destroy(ctx);
  tc_release(ctx);
  fused_release(ctx);
  if (ctx->Ht) cudaFree(ctx->Ht);
  release_data(ctx);
  release_hybrid(ctx);
  for (int i = 0; i < 2; i++) {
    if (ctx->W[i]) cudaFree(ctx->W[i]);
    if (ctx->H[i]) cudaFree(ctx->H[i]);
    if (ctx->Wlo[i]) cudaFree(ctx->Wlo[i]);
    if (ctx->Hlo[i]) cudaFree(ctx->Hlo[i]);
  }
  void *ptrs[] = {ctx->num, ctx->rowsumH, ctx->hsum, ctx->colsumW, ctx->rsh32, ctx->dred, ctx->stage, ctx->Q, ctx->Qlo,
                  ctx->dscal, ctx->flags, ctx->errors_dev};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
  for (int c = 0; c < 8; c++)
    if (ctx->comm_ev[c]) cudaEventDestroy(ctx->comm_ev[c]);
  if (ctx->comm_done) cudaEventDestroy(ctx->comm_done);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
  return KLNMF_OK;
}

int klnmf_set_stream(klnmf_ctx *ctx, void *cuda_stream) {
  KL_CHECK(ctx, KLNMF_EINVAL, "ctx is NULL");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (cuda_stream == nullptr) {
    if (!ctx->own_stream) {
      KL_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    return KLNMF_OK;
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return KLNMF_OK;
}

int klnmf_set_scratch_limit(klnmf_ctx *ctx, int64_t bytes) {
  KL_CHECK(ctx && bytes > 0, KLNMF_EINVAL, "bad scratch limit");
  KL_CHECK(!ctx->Q, KLNMF_ESTATE, "scratch limit must be set before the data");
  ctx->scratch_limit = bytes;
  return KLNMF_OK;
}

// ------------------------------------------------------------------------------------------------
int klnmf_set_dense_host(klnmf_ctx *ctx, const void *X, int dtype, int64_t ld) {
  KL_CHECK(ctx && (X || ctx->n == 0), KLNMF_EINVAL, "set_dense_host: NULL argument");
  KL_CHECK(ld >= ctx->f, KLNMF_EINVAL, "set_dense_host: ld %lld < f %lld", (long long)ld, (long long)ctx->f);
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, false));
  release_data(ctx);
  ctx->ldx = round_up(ctx->f, 32);
  const int64_t bytes = (ctx->n > 0 ? ctx->n : 1) * ctx->ldx * ctx->es;
  KL_TRY(dmalloc(&ctx->X, bytes));
  ctx->x_owned = true;
  if (ctx->ldx != ctx->f) KL_CUDA(cudaMemsetAsync(ctx->X, 0, bytes, ctx->stream));
  KL_TRY(upload_matrix(ctx, X, dtype, ld, ctx->X, ctx->ldx, ctx->n, ctx->f, false));
  KL_TRY(ensure_state(ctx));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_set_dense_blocks_host(klnmf_ctx *ctx, int n_blocks, const void *const *X, const int *dtypes,
                                const int64_t *lds, const int64_t *cols, const double *scales, const int *product_f32) {
  KL_CHECK(ctx && n_blocks >= 1 && X && dtypes && lds && cols && scales, KLNMF_EINVAL, "set_dense_blocks_host: NULL argument");
  int64_t total = 0;
  for (int b = 0; b < n_blocks; b++) {
    KL_CHECK(cols[b] >= 0 && lds[b] >= cols[b] && (X[b] || ctx->n == 0 || cols[b] == 0), KLNMF_EINVAL,
             "set_dense_blocks_host: bad block %d (cols %lld, ld %lld)", b, (long long)cols[b], (long long)lds[b]);
    KL_CHECK(dtypes[b] == KLNMF_F32 || dtypes[b] == KLNMF_F64, KLNMF_EINVAL, "set_dense_blocks_host: bad dtype of block %d", b);
    total += cols[b];
  }
  KL_CHECK(total == ctx->f, KLNMF_EINVAL, "set_dense_blocks_host: the blocks have %lld columns, the context %lld",
           (long long)total, (long long)ctx->f);
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, false));
  release_data(ctx);
  ctx->ldx = round_up(ctx->f, 32);
  const int64_t bytes = (ctx->n > 0 ? ctx->n : 1) * ctx->ldx * ctx->es;
  KL_TRY(dmalloc(&ctx->X, bytes));
  ctx->x_owned = true;
  if (ctx->ldx != ctx->f) KL_CUDA(cudaMemsetAsync(ctx->X, 0, bytes, ctx->stream));
  int64_t off = 0;
  for (int b = 0; b < n_blocks; b++) {
    char *dst = (char *)ctx->X + off * ctx->es;
    KL_TRY(upload_matrix(ctx, X[b], dtypes[b], lds[b], dst, ctx->ldx, ctx->n, cols[b], false, scales[b],
                         product_f32 && product_f32[b] && dtypes[b] == KLNMF_F32 ? 1 : 0));
    off += cols[b];
  }
  KL_TRY(ensure_state(ctx));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_set_dense_device(klnmf_ctx *ctx, const void *X_dev, int dtype, int64_t ld) {
  KL_CHECK(ctx && X_dev, KLNMF_EINVAL, "set_dense_device: NULL argument");
  KL_CHECK(ld >= ctx->f, KLNMF_EINVAL, "set_dense_device: ld < f");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, false));
  release_data(ctx);
  const int64_t ses = dtype == KLNMF_F64 ? 8 : 4;
  // The epilogues read whole 32-column chunks of X and rely on columns f..round_up(f, 32) being ZERO (x = 0 against
  // s = 0 gives q = 1 and an objective term of exactly 0).  A caller's buffer is therefore borrowed only when it has
  // no such padding columns (f % 32 == 0); anything else -- a column slice of a wider matrix, an uninitialised
  // pitch -- is copied into an owned, zero-padded buffer.
  const bool borrow = ses == ctx->es && (ld * ses) % 16 == 0 && ((uintptr_t)X_dev % 16) == 0 && ctx->f % 32 == 0;
  if (borrow) {
    ctx->X = const_cast<void *>(X_dev);
    ctx->ldx = ld;
    ctx->x_owned = false;
  } else {
    ctx->ldx = round_up(ctx->f, 32);
    const int64_t bytes = (ctx->n > 0 ? ctx->n : 1) * ctx->ldx * ctx->es;
    KL_TRY(dmalloc(&ctx->X, bytes));
    ctx->x_owned = true;
    if (ctx->ldx != ctx->f) KL_CUDA(cudaMemsetAsync(ctx->X, 0, bytes, ctx->stream));
    KL_TRY(launch_convert(ctx, X_dev, nullptr, dtype, ld, ctx->X, ctx->es, ctx->ldx, ctx->n, ctx->f, false));
  }
  KL_TRY(ensure_state(ctx));
  ctx->have_x = true;
  return KLNMF_OK;
}

static int alloc_csr(klnmf_ctx *ctx, int64_t nnz) {
  KL_TRY(dmalloc((void **)&ctx->indptr, (ctx->n + 1) * 8));
  KL_TRY(dmalloc((void **)&ctx->indices, nnz * 4));
  KL_TRY(dmalloc(&ctx->vals, nnz * ctx->es));
  ctx->csr_owned = true;
  return KLNMF_OK;
}

int klnmf_set_csr_host(klnmf_ctx *ctx, const int64_t *indptr, const int32_t *indices, const void *values, int dtype,
                       int64_t nnz) {
  KL_CHECK(ctx && indptr && nnz >= 0 && (nnz == 0 || (indices && values)), KLNMF_EINVAL, "set_csr_host: bad argument");
  KL_CHECK(ctx->f < ((int64_t)1 << 31), KLNMF_EINVAL, "set_csr_host: int32 column indices need f < 2^31");
  KL_CHECK(indptr[0] == 0 && indptr[ctx->n] == nnz, KLNMF_EINVAL, "set_csr_host: indptr does not span nnz");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, true));
  release_data(ctx);
  ctx->nnz = nnz;
  KL_TRY(alloc_csr(ctx, nnz));
  KL_CUDA(cudaMemcpyAsync(ctx->indptr, indptr, (ctx->n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (nnz > 0) {
    KL_CUDA(cudaMemcpyAsync(ctx->indices, indices, nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    {  // 1-D upload in bounded chunks (the staging buffer converts float64 -> float32 when needed)
      const int64_t ses = dtype == KLNMF_F64 ? 8 : 4, step = (int64_t)8 << 20;
      for (int64_t o = 0; o < nnz; o += step) {
        const int64_t c = nnz - o < step ? nnz - o : step;
        KL_TRY(upload_matrix(ctx, (const char *)values + o * ses, dtype, c, (char *)ctx->vals + o * ctx->es, c, 1, c, false));
      }
    }
  }
  ctx->bytes_h2d += (ctx->n + 1) * 8 + nnz * 4;
  KL_TRY(dmalloc(&ctx->qnz, nnz * ctx->es));
  KL_TRY(ensure_state(ctx));
  KL_TRY(launch_sum_vals(ctx));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_set_hybrid_min_cols(klnmf_ctx *ctx, int64_t cols) {
  KL_CHECK(ctx && cols >= 0, KLNMF_EINVAL, "set_hybrid_min_cols: bad argument");
  ctx->hybrid_min_cols = cols;
  return KLNMF_OK;
}

int klnmf_is_hybrid(klnmf_ctx *ctx) { return ctx && ctx->hyb ? 1 : 0; }

int klnmf_set_stacked_blocks_host(klnmf_ctx *ctx, int n_blocks, const klnmf_block *blocks) {
  KL_CHECK(ctx && blocks, KLNMF_EINVAL, "set_stacked_blocks_host: NULL argument");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, true));
  release_data(ctx);
  // Hybrid form (HybridSide above): the dense blocks stay dense when they are wide enough to pay for their contractions,
  // the mode multiplies in one pass and the context has no state yet.  hybrid_min_cols = 0: never; KLNMF_HYBRID=0: never,
  // =1: whenever possible.  (The shards of a multi-GPU fit must all take the same form: distributed.DeviceGroup decides
  // it once for all of them.)
  int64_t fd = 0, fs = 0, total = 0;
  int n_dense = 0, n_csr = 0;
  for (int b = 0; b < n_blocks; b++) {
    KL_CHECK(blocks[b].cols >= 0 && (blocks[b].kind == 0 || blocks[b].kind == 1), KLNMF_EINVAL, "set_stacked_blocks_host: bad block %d", b);
    if (blocks[b].kind == 0) { fd += blocks[b].cols; n_dense++; } else { fs += blocks[b].cols; n_csr++; }
    total += blocks[b].cols;
  }
  const char *he = getenv("KLNMF_HYBRID");
  const int64_t min_cols = ctx->hybrid_min_cols == 0 ? 0 : (he ? (atoi(he) == 0 ? 0 : 1) : ctx->hybrid_min_cols);   // an explicit 0 wins
  const int64_t ldd = round_up(fd, 32);
  const bool hybrid = min_cols > 0 && fd >= min_cols && n_dense > 0 && n_csr > 0 && fs > 0 && total == ctx->f &&
                      !ctx->W[0] && ctx->mode != KLNMF_MODE_TF32X3 && !ctx->debug_simt;
  if (!hybrid) {
    KL_TRY(stack_blocks_to_csr(ctx, n_blocks, blocks));
    KL_TRY(dmalloc(&ctx->qnz, ctx->nnz * ctx->es));
    KL_TRY(ensure_state(ctx));
    KL_TRY(launch_sum_vals(ctx));
    KL_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->have_x = true;
    return KLNMF_OK;
  }
  HybridSide *hy = new HybridSide();
  ctx->hyb = hy;
  hy->f_total = total; hy->fd = fd; hy->ld = ldd;
  const int64_t es = ctx->es;
  std::vector<klnmf_block> csr;
  int64_t col0 = 0, offd = 0, offs = 0;
  int rc = KLNMF_OK;
  auto fail = [&](int code) { ctx->f = total; release_data(ctx); release_hybrid(ctx); return code; };
  if ((rc = dmalloc(&hy->X, ctx->n * ldd * es)) != KLNMF_OK) return fail(rc);
  if (cudaMemsetAsync(hy->X, 0, ctx->n * ldd * es, ctx->stream) != cudaSuccess) return fail(KLNMF_ECUDA);
  for (int b = 0; b < n_blocks; b++) {
    const klnmf_block &s = blocks[b];
    if (s.kind == 0) {
      if (!(s.ld >= s.cols && (s.dense || s.cols == 0)) || !(s.dtype == KLNMF_F32 || s.dtype == KLNMF_F64)) {
        set_error("set_stacked_blocks_host: bad dense block %d", b);
        return fail(KLNMF_EINVAL);
      }
      hy->ranges.push_back({1, col0, s.cols, offd});
      rc = upload_matrix(ctx, s.dense, s.dtype, s.ld, (char *)hy->X + offd * es, ldd, ctx->n, s.cols, false, s.scale,
                         s.product_f32 && s.dtype == KLNMF_F32 ? 1 : 0);
      if (rc != KLNMF_OK) return fail(rc);
      offd += s.cols;
    } else {
      hy->ranges.push_back({0, col0, s.cols, offs});
      csr.push_back(s);
      offs += s.cols;
    }
    col0 += s.cols;
  }
  ctx->f = fs;                                   // from here on the context's CSR machinery sees the CSR columns only
  if ((rc = stack_blocks_to_csr(ctx, (int)csr.size(), csr.data())) != KLNMF_OK) return fail(rc);
  if ((rc = dmalloc(&ctx->qnz, ctx->nnz * es)) != KLNMF_OK || (rc = ensure_state(ctx)) != KLNMF_OK ||
      (rc = launch_sum_vals(ctx)) != KLNMF_OK)
    return fail(rc);
  const int64_t hb = ctx->k * ldd * es;
  {   // the ratio panel: as many rows as the scratch limit allows (like the dense path's panel, ensure_state)
    int64_t pr = ctx->scratch_limit / (ldd * es) / 128 * 128;
    if (pr < 128) pr = 128;
    if (pr > round_up(ctx->n > 0 ? ctx->n : 1, 128)) pr = round_up(ctx->n > 0 ? ctx->n : 1, 128);
    hy->panel_rows = pr;
  }
  if ((rc = dmalloc(&hy->Q, hy->panel_rows * ldd * es)) != KLNMF_OK || (rc = dmalloc(&hy->G, ctx->n * ctx->ldw * es)) != KLNMF_OK ||
      (rc = dmalloc(&hy->num, hb)) != KLNMF_OK || (rc = dmalloc(&hy->H[0], hb)) != KLNMF_OK ||
      (rc = dmalloc(&hy->H[1], hb)) != KLNMF_OK || (rc = dmalloc((void **)&hy->total, (ctx->k + 1) * 8)) != KLNMF_OK)
    return fail(rc);
  if (es == 4 && ((rc = dmalloc(&hy->Hr, hb)) != KLNMF_OK || (rc = dmalloc(&hy->Wr, hy->panel_rows * ctx->ldw * es)) != KLNMF_OK))
    return fail(rc);
  cudaMemsetAsync(hy->H[0], 0, hb, ctx->stream);
  cudaMemsetAsync(hy->H[1], 0, hb, ctx->stream);
  cudaMemsetAsync(hy->num, 0, hb, ctx->stream);
  cudaMemsetAsync(hy->G, 0, ctx->n * ctx->ldw * es, ctx->stream);
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_set_csr_device(klnmf_ctx *ctx, const int64_t *indptr_dev, const int32_t *indices_dev, const void *values_dev,
                         int dtype, int64_t nnz) {
  KL_CHECK(ctx && indptr_dev && nnz >= 0, KLNMF_EINVAL, "set_csr_device: bad argument");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, true));
  release_data(ctx);
  ctx->nnz = nnz;
  const int64_t ses = dtype == KLNMF_F64 ? 8 : 4;
  ctx->indptr = const_cast<int64_t *>(indptr_dev);
  ctx->indices = const_cast<int32_t *>(indices_dev);
  ctx->csr_owned = false;
  if (ses == ctx->es) {
    ctx->vals = const_cast<void *>(values_dev);
  } else {
    // value type differs from the arithmetic mode: keep a converted private copy of everything
    KL_TRY(alloc_csr(ctx, nnz));
    KL_CUDA(cudaMemcpyAsync(ctx->indptr, indptr_dev, (ctx->n + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    KL_CUDA(cudaMemcpyAsync(ctx->indices, indices_dev, nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    KL_TRY(launch_convert(ctx, values_dev, nullptr, dtype, nnz, ctx->vals, ctx->es, nnz, 1, nnz, false));
  }
  KL_TRY(dmalloc(&ctx->qnz, nnz * ctx->es));
  KL_TRY(ensure_state(ctx));
  KL_TRY(launch_sum_vals(ctx));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_create_column_view(klnmf_ctx *parent, int n_ranges, const int64_t *starts, const int64_t *widths, int64_t k,
                             klnmf_ctx **out) {
  KL_CHECK(parent && out && n_ranges >= 1 && starts && widths, KLNMF_EINVAL, "create_column_view: bad argument");
  KL_CHECK(parent->have_x && !parent->sparse, KLNMF_ESTATE, "create_column_view: the parent holds no dense data");
  *out = nullptr;
  int64_t f_sub = 0;
  for (int r = 0; r < n_ranges; r++) {
    KL_CHECK(starts[r] >= 0 && widths[r] > 0 && starts[r] + widths[r] <= parent->f, KLNMF_EINVAL,
             "create_column_view: range %d = [%lld, +%lld) leaves the parent's %lld columns", r, (long long)starts[r],
             (long long)widths[r], (long long)parent->f);
    f_sub += widths[r];
  }
  klnmf_ctx *ctx = nullptr;
  KL_TRY(klnmf_create(&ctx, parent->device, parent->n, f_sub, k, parent->mode));
  auto fail = [&](int rc) { klnmf_destroy(ctx); return rc; };
  ctx->sparse = false;
  ctx->ldx = round_up(f_sub, 32);
  const int64_t es = ctx->es;
  const int64_t bytes = (ctx->n > 0 ? ctx->n : 1) * ctx->ldx * es;
  int rc = dmalloc(&ctx->X, bytes);
  if (rc != KLNMF_OK) return fail(rc);
  ctx->x_owned = true;
  // the parent's pending work (its upload, the per-modality scaling) comes first; then the gather runs on the view's stream
  if (cudaStreamSynchronize(parent->stream) != cudaSuccess) { set_error("create_column_view: parent stream error"); return fail(KLNMF_ECUDA); }
  if (ctx->ldx != f_sub && cudaMemsetAsync(ctx->X, 0, bytes, ctx->stream) != cudaSuccess) return fail(KLNMF_ECUDA);
  int64_t off = 0;
  for (int r = 0; r < n_ranges && ctx->n > 0; r++) {
    if (cudaMemcpy2DAsync((char *)ctx->X + off * es, ctx->ldx * es, (const char *)parent->X + starts[r] * es, parent->ldx * es,
                          widths[r] * es, ctx->n, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) {
      set_error("create_column_view: device copy of range %d failed", r);
      return fail(KLNMF_ECUDA);
    }
    off += widths[r];
  }
  rc = ensure_state(ctx);
  if (rc != KLNMF_OK) return fail(rc);
  ctx->have_x = true;
  *out = ctx;
  return KLNMF_OK;
}

int klnmf_check_input(klnmf_ctx *ctx, int32_t out[2]) {
  KL_CHECK(ctx && out && ctx->have_x, KLNMF_ESTATE, "check_input: no data set");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_CUDA(cudaMemsetAsync(ctx->flags + FL_NEG, 0, 8, ctx->stream));
  if (ctx->sparse) KL_TRY(launch_check(ctx, ctx->vals, ctx->es, 1, ctx->nnz, ctx->nnz));
  else KL_TRY(launch_check(ctx, ctx->X, ctx->es, ctx->n, ctx->f, ctx->ldx));
  if (ctx->hyb) {
    const HybridSide *hy = (const HybridSide *)ctx->hyb;
    KL_TRY(launch_check(ctx, hy->X, ctx->es, ctx->n, hy->fd, hy->ld));
  }
  int *h = (int *)ctx->pinned;
  KL_CUDA(cudaMemcpyAsync(h, ctx->flags + FL_NEG, 8, cudaMemcpyDeviceToHost, ctx->stream));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  out[0] = h[0];
  out[1] = h[1];
  return KLNMF_OK;
}

// ------------------------------------------------------------------------------------------------
int klnmf_set_dictionary_host(klnmf_ctx *ctx, const double *H, int64_t ld) {
  KL_CHECK(ctx && H, KLNMF_EINVAL, "set_dictionary_host: NULL argument");
  KL_CHECK(ctx->have_x, KLNMF_ESTATE, "set the data before the dictionary");
  KL_CHECK(ld >= (ctx->hyb ? ((HybridSide *)ctx->hyb)->f_total : ctx->f), KLNMF_EINVAL, "set_dictionary_host: ld < f");
  KL_CUDA(cudaSetDevice(ctx->device));
  const int hc = ctx->hcur;
  if (ctx->hyb) {
    // the stack's columns go to their part: dense ranges k x fd row-major, CSR ranges transposed f_s x k
    HybridSide *hy = (HybridSide *)ctx->hyb;
    for (const auto &r : hy->ranges) {
      if (r.dense) KL_TRY(upload_matrix(ctx, H + r.col0, KLNMF_F64, ld, (char *)hy->H[hc] + r.off * ctx->es, hy->ld, ctx->k, r.cols, false));
      else KL_TRY(upload_matrix(ctx, H + r.col0, KLNMF_F64, ld, (char *)ctx->H[hc] + r.off * ctx->ldh * ctx->es, ctx->ldh, ctx->k, r.cols, true));
    }
    KL_TRY(launch_rowsum_h(ctx));              // row sums of the CSR part: what the sparse objective multiplies colsum(W) with
    ctx->have_h = true;
    ctx->ht_stale = true;
    return KLNMF_OK;
  }
  // dense layout k x f; sparse layout transposed f x k
  KL_TRY(upload_matrix(ctx, H, KLNMF_F64, ld, ctx->H[hc], ctx->ldh, ctx->k, ctx->f, ctx->sparse));
  if (ctx->split)
    KL_TRY(launch_split(ctx, (const float *)ctx->H[hc], (float *)ctx->H[hc], (float *)ctx->Hlo[hc], ctx->k, ctx->f,
                        ctx->ldh));
  KL_TRY(launch_rowsum_h(ctx));
  ctx->have_h = true;
  ctx->ht_stale = true;
  return KLNMF_OK;
}

int klnmf_get_dictionary_host(klnmf_ctx *ctx, double *H, int64_t ld) {
  KL_CHECK(ctx && H && ctx->have_h, KLNMF_ESTATE, "get_dictionary_host: no dictionary");
  KL_CUDA(cudaSetDevice(ctx->device));
  const int hc = ctx->hcur;
  if (ctx->hyb) {
    HybridSide *hy = (HybridSide *)ctx->hyb;
    KL_CHECK(ld >= hy->f_total, KLNMF_EINVAL, "get_dictionary_host: ld < f");
    for (const auto &r : hy->ranges) {
      if (r.dense) KL_TRY(download_matrix(ctx, (char *)hy->H[hc] + r.off * ctx->es, nullptr, hy->ld, H + r.col0, KLNMF_F64, ld, ctx->k, r.cols, false));
      else KL_TRY(download_matrix(ctx, (char *)ctx->H[hc] + r.off * ctx->ldh * ctx->es, nullptr, ctx->ldh, H + r.col0, KLNMF_F64, ld, ctx->k, r.cols, true));
    }
    return KLNMF_OK;
  }
  return download_matrix(ctx, ctx->H[hc], ctx->split ? ctx->Hlo[hc] : nullptr, ctx->ldh, H, KLNMF_F64, ld, ctx->k,
                         ctx->f, ctx->sparse);
}

int klnmf_init_coefficients(klnmf_ctx *ctx) {
  KL_CHECK(ctx && ctx->have_x && ctx->have_h, KLNMF_ESTATE, "init_coefficients needs data and dictionary");
  KL_CUDA(cudaSetDevice(ctx->device));
  if (ctx->n == 0) { ctx->have_w = true; return KLNMF_OK; }
  if (ctx->sparse) {
    if (ctx->hyb) {   // W0 = X_d.H_d^T + X_s.H_s^T: the dense block's product first, the CSR pass starts from it
      HybridSide *hy = (HybridSide *)ctx->hyb;
      KL_TRY(hybrid_round_dictionary(ctx));
      for (int64_t r0 = 0; r0 < ctx->n; r0 += hy->panel_rows) {
        GemmDesc c{};
        c.M = ctx->n - r0 < hy->panel_rows ? ctx->n - r0 : hy->panel_rows; c.N = ctx->k; c.K = hy->fd;
        c.A = (const char *)hy->X + r0 * hy->ld * ctx->es; c.a_sm = hy->ld; c.a_sk = 1;
        c.B = hy->H[ctx->hcur]; c.b_sk = 1; c.b_sn = hy->ld;
        if (hy->Hr) {   // rounded copies of the data panel (through the ratio panel's buffer) and of the dictionary
          KL_TRY(launch_split(ctx, (const float *)c.A, (float *)hy->Q, nullptr, c.M, hy->ld, hy->ld));
          c.A = hy->Q; c.B = hy->Hr;
        }
        c.out = (char *)hy->G + r0 * ctx->ldw * ctx->es; c.ldo = ctx->ldw;
        KL_TRY(dense_gemm(ctx, EPI_STORE, c));
      }
      KL_TRY(sparse_init_w(ctx, hy->G));
    } else {
      KL_TRY(sparse_init_w(ctx));
    }
    ctx->have_w = true;
    return KLNMF_OK;
  }
  const int64_t es = ctx->es;
  const int cur = ctx->cur, hc = ctx->hcur;
  for (int64_t r0 = 0; r0 < ctx->n; r0 += ctx->panel_rows) {
    const int64_t rows = ctx->n - r0 < ctx->panel_rows ? ctx->n - r0 : ctx->panel_rows;
    GemmDesc d{};
    d.M = rows; d.N = ctx->k; d.K = ctx->f;
    const char *Xp = (const char *)ctx->X + r0 * ctx->ldx * es;
    if (ctx->split) {
      // X must enter the split-TF32 contraction as (hi, lo) too: stage the panel through the Q scratch
      for (int64_t rr = 0; rr < rows; rr += 1 << 20) {
        const int64_t r = rows - rr < (1 << 20) ? rows - rr : (1 << 20);
        KL_CUDA(cudaMemcpy2DAsync((char *)ctx->Q + rr * ctx->ldq * es, ctx->ldq * es, Xp + rr * ctx->ldx * es,
                                  ctx->ldx * es, ctx->f * es, r, cudaMemcpyDeviceToDevice, ctx->stream));
      }
      // (TF32R multiplies the rounded high parts only: no low-part panel)
      KL_TRY(launch_split(ctx, (const float *)ctx->Q, (float *)ctx->Q, ctx->single_pass ? nullptr : (float *)ctx->Qlo, rows,
                          ctx->f, ctx->ldq));
      d.A = ctx->Q; d.a_sm = ctx->ldq; d.A_lo = ctx->single_pass ? nullptr : ctx->Qlo;
    } else {
      d.A = Xp; d.a_sm = ctx->ldx;
    }
    d.a_sk = 1;
    d.B = ctx->H[hc]; d.b_sk = 1; d.b_sn = ctx->ldh; d.B_lo = ctx->split ? ctx->Hlo[hc] : nullptr;
    d.out = (char *)ctx->W[cur] + r0 * ctx->ldw * es; d.ldo = ctx->ldw;
    d.out_lo = ctx->split ? (char *)ctx->Wlo[cur] + r0 * ctx->ldw * es : nullptr;
    d.single_pass = ctx->single_pass ? 1 : 0;
    KL_TRY(dense_gemm(ctx, EPI_STORE, d));
  }
  ctx->have_w = true;
  return KLNMF_OK;
}

int klnmf_set_coefficients_host(klnmf_ctx *ctx, const double *W, int64_t ld) {
  KL_CHECK(ctx && (W || ctx->n == 0), KLNMF_EINVAL, "set_coefficients_host: NULL argument");
  KL_CHECK(ctx->have_x, KLNMF_ESTATE, "set the data before the coefficients");
  KL_CHECK(ld >= ctx->k, KLNMF_EINVAL, "set_coefficients_host: ld < k");
  KL_CUDA(cudaSetDevice(ctx->device));
  const int cur = ctx->cur;
  KL_TRY(upload_matrix(ctx, W, KLNMF_F64, ld, ctx->W[cur], ctx->ldw, ctx->n, ctx->k, false));
  if (ctx->split)
    KL_TRY(launch_split(ctx, (const float *)ctx->W[cur], (float *)ctx->W[cur], (float *)ctx->Wlo[cur], ctx->n, ctx->k,
                        ctx->ldw));
  ctx->have_w = true;
  return KLNMF_OK;
}

int klnmf_get_coefficients_host(klnmf_ctx *ctx, double *W, int64_t ld) {
  KL_CHECK(ctx && (W || ctx->n == 0) && ctx->have_w, KLNMF_ESTATE, "get_coefficients_host: no coefficients");
  KL_CUDA(cudaSetDevice(ctx->device));
  const int cur = ctx->cur;
  return download_matrix(ctx, ctx->W[cur], ctx->split ? ctx->Wlo[cur] : nullptr, ctx->ldw, W, KLNMF_F64, ld, ctx->n,
                         ctx->k, false);
}

int klnmf_coefficients_device(klnmf_ctx *ctx, void **ptr, int64_t *ld, int *dtype) {
  KL_CHECK(ctx && ptr && ld && dtype && ctx->have_w, KLNMF_ESTATE, "coefficients_device: no coefficients");
  KL_CHECK(!ctx->split, KLNMF_ESTATE, "split-TF32 state is a (hi, lo) pair; use get_coefficients_host");
  *ptr = ctx->W[ctx->cur]; *ld = ctx->ldw; *dtype = ctx->es == 8 ? KLNMF_F64 : KLNMF_F32;
  return KLNMF_OK;
}

int klnmf_dictionary_device(klnmf_ctx *ctx, void **ptr, int64_t *ld, int *dtype) {
  KL_CHECK(ctx && ptr && ld && dtype && ctx->have_h, KLNMF_ESTATE, "dictionary_device: no dictionary");
  KL_CHECK(!ctx->hyb, KLNMF_ESTATE, "dictionary_device: a hybrid stack keeps its dictionary in two parts; use get_dictionary_host");
  KL_CHECK(!ctx->split, KLNMF_ESTATE, "split-TF32 state is a (hi, lo) pair; use get_dictionary_host");
  *ptr = ctx->H[ctx->hcur]; *ld = ctx->ldh; *dtype = ctx->es == 8 ? KLNMF_F64 : KLNMF_F32;
  return KLNMF_OK;
}

// ------------------------------------------------------------------------------------------------
int klnmf_run(klnmf_ctx *ctx, int max_iter, double tol_abs, int fit, double *errors_out, int *n_errors, int *n_iter) {
  return klnmf_run_resume(ctx, max_iter, tol_abs, fit, INFINITY, errors_out, n_errors, n_iter);
}

int klnmf_run_resume(klnmf_ctx *ctx, int max_iter, double tol_abs, int fit, double prev_objective, double *errors_out,
                     int *n_errors, int *n_iter) {
  KL_CHECK(ctx && max_iter >= 1, KLNMF_EINVAL, "klnmf_run: max_iter must be >= 1 (reference: range(1, max_iter+1))");
  KL_CHECK(!(prev_objective != prev_objective), KLNMF_EINVAL, "klnmf_run_resume: previous objective is NaN");
  KL_CHECK(ctx->have_x && ctx->have_h && ctx->have_w, KLNMF_ESTATE, "klnmf_run needs data, dictionary and coefficients");
  KL_CUDA(cudaSetDevice(ctx->device));
  if (ctx->errors_cap < max_iter) {
    if (ctx->errors_dev) { KL_CUDA(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->errors_dev); }
    KL_TRY(dmalloc((void **)&ctx->errors_dev, (int64_t)max_iter * 8));
    ctx->errors_cap = max_iter;
  }
  // tol == 0 is the learner's "run exactly `iterations` updates" idiom (learner.py:12,39): the float64
  // reference then breaks only if the objective RISES.  Objective noise of the reduced-precision
  // contractions must not trigger that break, so the rise has to exceed the mode's noise floor.
  // The same holds for any tol below that floor (a tiny positive tol near convergence): decide_kernel applies the
  // slack whenever tol_abs <= slack * |previous objective|, and the plain reference test above it.
  const double slack = ctx->mode == KLNMF_MODE_FP64 ? 0.0 : (ctx->mode == KLNMF_MODE_TF32X3 ? 1e-6 : (ctx->mode == KLNMF_MODE_TF32R ? 1e-5 : 1e-4));
  double *hp = ctx->pinned;
  hp[DS_KL] = 0.0; hp[DS_PREV] = prev_objective; hp[DS_WHSUM] = slack; hp[DS_TOL] = tol_abs;
  KL_CUDA(cudaMemcpyAsync(ctx->dscal + DS_KL, hp + DS_KL, 8, cudaMemcpyHostToDevice, ctx->stream));
  KL_CUDA(cudaMemcpyAsync(ctx->dscal + DS_PREV, hp + DS_PREV, 8, cudaMemcpyHostToDevice, ctx->stream));
  KL_CUDA(cudaMemcpyAsync(ctx->dscal + DS_WHSUM, hp + DS_WHSUM, 8, cudaMemcpyHostToDevice, ctx->stream));
  KL_CUDA(cudaMemcpyAsync(ctx->dscal + DS_TOL, hp + DS_TOL, 8, cudaMemcpyHostToDevice, ctx->stream));
  KL_CUDA(cudaMemsetAsync(ctx->flags, 0, 2 * 4, ctx->stream));
  KL_TRY(reset_reduction(ctx));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));   // pinned scratch is reused below

  KL_TRY(setup_num_chunks(ctx, fit));
  Profiler prof_store;
  Profiler *prof = ctx->profile ? &prof_store : nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEventCreate(&ev0);
  cudaEventCreate(&ev1);
  cudaEventRecord(ev0, ctx->stream);

  const int cur0 = ctx->cur, hcur0 = ctx->hcur;
  const int64_t hbytes = (ctx->sparse ? ctx->f : ctx->k) * ctx->ldh * (int64_t)ctx->es;
  int *hflags = (int *)(ctx->pinned + 16);
  int rc = KLNMF_OK;
  hflags[FL_STOP] = 0; hflags[FL_NERR] = 0;
  // one reference iteration (nmf.py:212-222), enqueued on the context's stream; nothing in it synchronises
  auto enqueue_iteration = [&](int it) -> int {
    int r = KLNMF_OK;
    ctx->num_reduced = false;
    if (fit) r = launch_zero(ctx, ctx->num, ctx->num_bytes);
    if (fit && r == KLNMF_OK && ctx->hyb) {
      HybridSide *hy = (HybridSide *)ctx->hyb;
      r = launch_zero(ctx, hy->num, ctx->k * hy->ld * (int64_t)ctx->es);
    }
    if (r == KLNMF_OK && ctx->sparse && ctx->n > 0) {
      if (ctx->hyb) {
        r = hybrid_rows_passes(ctx, fit, prof);
      } else {
        PhaseTimer t(ctx, prof, PH_RATIO);
        r = sparse_rows(ctx, 0, fit ? 2 : 0);
      }
      if (r == KLNMF_OK && fit) {
        PhaseTimer t(ctx, prof, PH_NUM);
        r = sparse_scatter(ctx, false);
      }
    } else if (r == KLNMF_OK) {
      r = dense_iteration(ctx, fit, false, prof);
    }
    if (r == KLNMF_OK && ctx->world > 1) {
      PhaseTimer t(ctx, prof, PH_COMM);
      // one NCCL group: the k x f numerator and the handful of FP64 partials travel in one launch
      r = nccl_group_start();
      if (r == KLNMF_OK) r = nccl_allreduce_sum_f64(ctx, ctx->dred, 2 + ctx->k);
      if (r == KLNMF_OK && fit && !ctx->num_reduced)     // (a shard without rows, the sparse path: the whole buffer at once)
        r = nccl_allreduce_sum(ctx, ctx->num, num_elems(ctx), ctx->es);
      if (r == KLNMF_OK && fit && ctx->hyb) {           // hybrid stack: the dense block's numerator travels in the same group
        HybridSide *hy = (HybridSide *)ctx->hyb;
        r = nccl_allreduce_sum(ctx, hy->num, ctx->k * hy->ld, ctx->es);
      }
      const int r_end = nccl_group_end();
      if (r == KLNMF_OK) r = r_end;
      if (r == KLNMF_OK && ctx->num_reduced && cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0) != cudaSuccess) {
        set_error("klnmf_run: cudaStreamWaitEvent failed");
        r = KLNMF_ECUDA;
      }
    }
    if (r == KLNMF_OK) r = launch_decide(ctx, it);
    if (r == KLNMF_OK && fit) {
      PhaseTimer t(ctx, prof, PH_DICT);
      const int hc = ctx->hcur;
      if (ctx->hyb) {
        HybridSide *hy = (HybridSide *)ctx->hyb;
        r = launch_dict_update_hybrid(ctx, ctx->H[hc], ctx->H[hc ^ 1], hy->H[hc], hy->H[hc ^ 1], hy->num, hy->fd, hy->ld, hy->total);
      } else
      r = ctx->sparse ? launch_dict_update_t(ctx, ctx->H[hc], ctx->H[hc ^ 1])
                      : launch_dict_update(ctx, ctx->H[hc], ctx->H[hc ^ 1], ctx->split ? ctx->Hlo[hc ^ 1] : nullptr,
                                           ctx->centered ? ctx->colsumW : nullptr);
      ctx->hcur ^= 1;
      ctx->ht_stale = true;
    }
    ctx->cur ^= 1;
    return r;
  };
  // with tol > 0 the host looks at the stop flag every 32 iterations (the kernels gate themselves on it anyway)
  auto stopped_early = [&](int it_done) -> bool {
    if (!(tol_abs > 0.0) || it_done >= max_iter) return false;
    cudaMemcpyAsync(hflags, ctx->flags, 8, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return hflags[FL_STOP] != 0;
  };
  // (Replaying two captured iterations as a CUDA graph was measured and dropped: the launches above are already
  // asynchronous with no host synchronisation in the loop, and instantiating the graph costs more per call than it
  // saves -- cfg1, 100 iterations: 10 600 it/s with plain launches, 8400 it/s with the graph; profiles/r1_s4_run51_*.log.)
  int it = 0;
  for (; it < max_iter && rc == KLNMF_OK; it++) {
    rc = enqueue_iteration(it);
    if (rc == KLNMF_OK && (it % 32) == 31 && stopped_early(it + 1)) break;
  }
  cudaEventRecord(ev1, ctx->stream);
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  if (rc == KLNMF_OK && se != cudaSuccess) {
    set_error("klnmf_run: device error: %s", cudaGetErrorString(se));
    rc = KLNMF_ECUDA;
  }
  if (rc == KLNMF_OK) {
    cudaMemcpy(hflags, ctx->flags, 8, cudaMemcpyDeviceToHost);
    const int ne = hflags[FL_NERR], stopped = hflags[FL_STOP];
    // updates applied == errors recorded; later (gated) iterations wrote nothing
    ctx->cur = (cur0 + ne) & 1;
    ctx->hcur = fit ? (hcur0 + ne) & 1 : hcur0;
    if (n_errors) *n_errors = ne;
    if (n_iter) *n_iter = stopped ? ne + 1 : max_iter;
    if (errors_out && ne > 0) {
      cudaMemcpy(errors_out, ctx->errors_dev, (size_t)ne * 8, cudaMemcpyDeviceToHost);
      ctx->bytes_d2h += (int64_t)ne * 8;
    }
    // a stop leaves FL_STOP set; clear it so that later calls (error, reconstruct) are not gated
    cudaMemsetAsync(ctx->flags, 0, 4, ctx->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    for (int i = 0; i < 6; i++) ctx->prof_ms[i] = 0.0;
    for (int i = 0; i < 5; i++) ctx->prof_cnt[i] = 0;
    ctx->prof_ms[PH_TOTAL] = ms;
    if (prof)
      for (auto &e : prof->ev) {
        float t = 0.f;
        cudaEventElapsedTime(&t, e.a, e.b);
        ctx->prof_ms[e.phase] += t;
        ctx->prof_cnt[e.phase] += 1;
      }
  } else {
    ctx->cur = cur0;
    ctx->hcur = hcur0;
  }
  if (prof)
    for (auto &e : prof->ev) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
  return rc;
}

int klnmf_error(klnmf_ctx *ctx, double *out) {
  KL_CHECK(ctx && out, KLNMF_EINVAL, "klnmf_error: NULL argument");
  KL_CHECK(ctx->have_x && ctx->have_h && ctx->have_w, KLNMF_ESTATE, "klnmf_error needs data, dictionary and coefficients");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(reset_reduction(ctx));
  if (ctx->sparse) {
    if (ctx->n > 0) {
      if (ctx->hyb) {
        const HybridSide *hy = (const HybridSide *)ctx->hyb;
        KL_TRY(hybrid_round_dictionary(ctx));
        for (int64_t r0 = 0; r0 < ctx->n; r0 += hy->panel_rows)
          KL_TRY(hybrid_dense_half(ctx, true, r0, ctx->n - r0 < hy->panel_rows ? ctx->n - r0 : hy->panel_rows));
      }
      KL_TRY(sparse_rows(ctx, 1));
    }
  }
  else KL_TRY(dense_iteration(ctx, 0, true, nullptr));
  if (ctx->world > 1) KL_TRY(nccl_allreduce_sum_f64(ctx, ctx->dred, 2 + ctx->k));
  std::vector<double> h(2 + ctx->k), rs(ctx->k);
  KL_CUDA(cudaMemcpyAsync(h.data(), ctx->dred, (2 + ctx->k) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  KL_CUDA(cudaMemcpyAsync(rs.data(), ctx->rowsumH, ctx->k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  double e = h[0];
  if (ctx->sparse) {
    double wh = 0.0;
    for (int64_t a = 0; a < ctx->k; a++) wh += h[2 + a] * rs[a];
    e = e - h[1] + wh;
  }
  *out = e;
  KL_TRY(reset_reduction(ctx));
  return KLNMF_OK;
}

// _updated_H with Q=None (nmf.py:345-351): H <- rownorm(H (.) W^T Q(W,H)) with the CURRENT W.
int klnmf_dictionary_step(klnmf_ctx *ctx) {
  KL_CHECK(ctx && ctx->have_x && ctx->have_h && ctx->have_w, KLNMF_ESTATE, "dictionary_step needs data, dictionary, coefficients");
  KL_CHECK(!ctx->hyb, KLNMF_ESTATE, "dictionary_step: not available on a hybrid (dense + CSR) stack");
  KL_CUDA(cudaSetDevice(ctx->device));
  const int64_t hbytes = (ctx->sparse ? ctx->f : ctx->k) * ctx->ldh * (int64_t)ctx->es;
  KL_TRY(setup_num_chunks(ctx, 1));
  KL_TRY(reset_reduction(ctx));
  ctx->num_reduced = false;
  KL_TRY(launch_zero(ctx, ctx->num, ctx->num_bytes));
  if (ctx->sparse) {
    if (ctx->n > 0) {
      KL_TRY(sparse_rows(ctx, 0, 2));
      KL_TRY(sparse_scatter(ctx, true));
    }
  } else {
    KL_TRY(dense_iteration(ctx, 1, false, nullptr, true));
  }
  if (ctx->world > 1) {
    if (!ctx->num_reduced) KL_TRY(nccl_allreduce_sum(ctx, ctx->num, num_elems(ctx), ctx->es));
    else KL_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->comm_done, 0));
  }
  const int hc = ctx->hcur;
  KL_TRY(ctx->sparse ? launch_dict_update_t(ctx, ctx->H[hc], ctx->H[hc ^ 1])
                     : launch_dict_update(ctx, ctx->H[hc], ctx->H[hc ^ 1], ctx->split ? ctx->Hlo[hc ^ 1] : nullptr));
  ctx->hcur ^= 1;
  ctx->ht_stale = true;
  KL_TRY(reset_reduction(ctx));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  return KLNMF_OK;
}

// _Q (nmf.py:325-336): dense -> n x f matrix; CSR -> nnz values in the order of the stored entries.
int klnmf_ratio_host(klnmf_ctx *ctx, void *out, int dtype, int64_t ld) {
  KL_CHECK(ctx && out && ctx->have_x && ctx->have_h && ctx->have_w, KLNMF_ESTATE, "ratio_host needs data, dictionary, coefficients");
  KL_CHECK(!ctx->hyb, KLNMF_ESTATE, "ratio_host: not available on a hybrid (dense + CSR) stack");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(reset_reduction(ctx));
  if (ctx->sparse) {
    if (ctx->nnz > 0) {
      KL_TRY(sparse_rows(ctx, 0, 1));
      KL_TRY(download_matrix(ctx, ctx->qnz, nullptr, ctx->nnz, out, dtype, ctx->nnz, 1, ctx->nnz, false));
    }
  } else {
    KL_CHECK(ld >= ctx->f, KLNMF_EINVAL, "ratio_host: ld < f");
    KL_TRY(dense_iteration(ctx, 0, false, nullptr, false, out, dtype, ld));
  }
  KL_TRY(reset_reduction(ctx));
  KL_CUDA(cudaStreamSynchronize(ctx->stream));
  return KLNMF_OK;
}

// _special_sparse_dot (nmf.py:52-70): (W.H) sampled at the stored entries of the CSR data.
int klnmf_sddmm_host(klnmf_ctx *ctx, double *out_vals) {
  KL_CHECK(ctx && ctx->sparse && ctx->have_x && ctx->have_h && ctx->have_w, KLNMF_ESTATE, "sddmm_host needs CSR data, dictionary, coefficients");
  KL_CHECK(!ctx->hyb, KLNMF_ESTATE, "sddmm_host: not available on a hybrid (dense + CSR) stack");
  KL_CUDA(cudaSetDevice(ctx->device));
  if (ctx->nnz == 0) return KLNMF_OK;
  KL_CHECK(out_vals, KLNMF_EINVAL, "sddmm_host: NULL output");
  KL_TRY(reset_reduction(ctx));
  KL_TRY(sparse_rows(ctx, 3));
  KL_TRY(download_matrix(ctx, ctx->qnz, nullptr, ctx->nnz, out_vals, KLNMF_F64, ctx->nnz, 1, ctx->nnz, false));
  KL_TRY(reset_reduction(ctx));
  return KLNMF_OK;
}

int klnmf_reconstruct_host(klnmf_ctx *ctx, const double *H_dest, int64_t f_dest, int64_t ld_h, double *out,
                           int64_t ld_out) {
  KL_CHECK(ctx && H_dest && (out || ctx->n == 0) && f_dest > 0, KLNMF_EINVAL, "reconstruct_host: bad argument");
  KL_CHECK(ctx->have_w, KLNMF_ESTATE, "reconstruct_host: no coefficients");
  KL_CHECK(ld_h >= f_dest && ld_out >= f_dest, KLNMF_EINVAL, "reconstruct_host: leading dimension too small");
  KL_CUDA(cudaSetDevice(ctx->device));
  if (ctx->n == 0) return KLNMF_OK;
  const int64_t es = ctx->es;
  const int64_t ldd = round_up(f_dest, 32);
  void *Hd = nullptr, *Hdlo = nullptr, *panel = nullptr;
  KL_TRY(dmalloc(&Hd, ctx->k * ldd * es));
  cudaMemsetAsync(Hd, 0, ctx->k * ldd * es, ctx->stream);
  int rc = upload_matrix(ctx, H_dest, KLNMF_F64, ld_h, Hd, ldd, ctx->k, f_dest, false);
  if (rc == KLNMF_OK && ctx->split) {
    rc = dmalloc(&Hdlo, ctx->k * ldd * es);
    if (rc == KLNMF_OK) {
      cudaMemsetAsync(Hdlo, 0, ctx->k * ldd * es, ctx->stream);
      rc = launch_split(ctx, (const float *)Hd, (float *)Hd, (float *)Hdlo, ctx->k, f_dest, ldd);
    }
  }
  int64_t prow = ((int64_t)1 << 30) / (ldd * es);
  prow = prow / 128 * 128;
  if (prow < 128) prow = 128;
  if (prow > ctx->n) prow = ctx->n;
  if (rc == KLNMF_OK) rc = dmalloc(&panel, prow * ldd * es);
  const int cur = ctx->cur;
  for (int64_t r0 = 0; r0 < ctx->n && rc == KLNMF_OK; r0 += prow) {
    const int64_t rows = ctx->n - r0 < prow ? ctx->n - r0 : prow;
    GemmDesc d{};
    d.M = rows; d.N = f_dest; d.K = ctx->k;
    d.A = (const char *)ctx->W[cur] + r0 * ctx->ldw * es; d.a_sm = ctx->ldw; d.a_sk = 1;
    d.A_lo = ctx->split ? (const char *)ctx->Wlo[cur] + r0 * ctx->ldw * es : nullptr;
    d.B = Hd; d.b_sk = ldd; d.b_sn = 1; d.B_lo = Hdlo;
    d.out = panel; d.ldo = ldd;
    rc = dense_gemm(ctx, EPI_STORE, d);
    if (rc == KLNMF_OK)
      rc = download_matrix(ctx, panel, nullptr, ldd, out + r0 * ld_out, KLNMF_F64, ld_out, rows, f_dest, false);
  }
  cudaStreamSynchronize(ctx->stream);
  if (Hd) cudaFree(Hd);
  if (Hdlo) cudaFree(Hdlo);
  if (panel) cudaFree(panel);
  return rc;
}

// ------------------------------------------------------------------------------------------------
int klnmf_nccl_load(const char *libnccl_path) { return nccl_load(libnccl_path); }
int klnmf_nccl_unique_id(void *id128) {
  KL_CHECK(id128, KLNMF_EINVAL, "nccl_unique_id: NULL");
  return nccl_unique_id(id128);
}
int klnmf_comm_init(klnmf_ctx *ctx, const void *id128, int rank, int world) {
  KL_CHECK(ctx && id128, KLNMF_EINVAL, "comm_init: NULL argument");
  return nccl_comm_init(ctx, id128, rank, world);
}
int klnmf_comm_create(void **comm, int device, const void *id128, int rank, int world) {
  KL_CHECK(comm && id128, KLNMF_EINVAL, "comm_create: NULL argument");
  const int ndev = klnmf_device_count();
  KL_CHECK(ndev > 0, KLNMF_ENODEVICE, "no CUDA device: libklnmf has no CPU path");
  KL_CHECK(device >= 0 && device < ndev, KLNMF_EINVAL, "device %d out of range (%d devices)", device, ndev);
  return nccl_comm_create(comm, device, id128, rank, world);
}
int klnmf_comm_attach(klnmf_ctx *ctx, void *comm, int rank, int world) {
  KL_CHECK(ctx && comm && world >= 1 && rank >= 0 && rank < world, KLNMF_EINVAL, "comm_attach: bad argument");
  nccl_comm_destroy(ctx);          // an owned communicator of an earlier klnmf_comm_init
  ctx->comm = comm;
  ctx->comm_owned = false;
  ctx->rank = rank;
  ctx->world = world;
  return KLNMF_OK;
}
int klnmf_comm_destroy(void *comm) { return nccl_comm_free(comm); }

// ------------------------------------------------------------------------------------------------
int klnmf_fill_dense_synthetic(klnmf_ctx *ctx, uint64_t seed) {
  KL_CHECK(ctx, KLNMF_EINVAL, "ctx is NULL");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, false));
  release_data(ctx);
  ctx->ldx = round_up(ctx->f, 32);
  const int64_t bytes = (ctx->n > 0 ? ctx->n : 1) * ctx->ldx * ctx->es;
  KL_TRY(dmalloc(&ctx->X, bytes));
  ctx->x_owned = true;
  if (ctx->ldx != ctx->f) KL_CUDA(cudaMemsetAsync(ctx->X, 0, bytes, ctx->stream));
  KL_TRY(launch_fill_uniform(ctx, ctx->X, ctx->es, ctx->n, ctx->f, ctx->ldx, seed));
  KL_TRY(ensure_state(ctx));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_fill_csr_synthetic(klnmf_ctx *ctx, int64_t nnz_per_row, uint64_t seed) {
  KL_CHECK(ctx && nnz_per_row > 0 && nnz_per_row <= ctx->f, KLNMF_EINVAL, "fill_csr_synthetic: bad nnz_per_row");
  KL_CUDA(cudaSetDevice(ctx->device));
  KL_TRY(kind_guard(ctx, true));
  release_data(ctx);
  ctx->nnz = ctx->n * nnz_per_row;
  KL_TRY(alloc_csr(ctx, ctx->nnz));
  KL_TRY(dmalloc(&ctx->qnz, ctx->nnz * ctx->es));
  KL_TRY(sparse_fill_synthetic(ctx, nnz_per_row, seed));
  KL_TRY(ensure_state(ctx));
  KL_TRY(launch_sum_vals(ctx));
  ctx->have_x = true;
  return KLNMF_OK;
}

int klnmf_get_dense_host(klnmf_ctx *ctx, void *X, int dtype, int64_t ld) {
  KL_CHECK(ctx && X && ctx->have_x && !ctx->sparse, KLNMF_ESTATE, "get_dense_host: no dense data");
  KL_CUDA(cudaSetDevice(ctx->device));
  return download_matrix(ctx, ctx->X, nullptr, ctx->ldx, X, dtype, ld, ctx->n, ctx->f, false);
}

int klnmf_counters(klnmf_ctx *ctx, int64_t out[4]) {
  KL_CHECK(ctx && out, KLNMF_EINVAL, "counters: NULL argument");
  out[0] = ctx->n_launch; out[1] = ctx->n_nccl; out[2] = ctx->bytes_h2d; out[3] = ctx->bytes_d2h;
  return KLNMF_OK;
}

int klnmf_last_run_profile(klnmf_ctx *ctx, double ms[6], int64_t counts[5]) {
  KL_CHECK(ctx && ms && counts, KLNMF_EINVAL, "last_run_profile: NULL argument");
  for (int i = 0; i < 6; i++) ms[i] = ctx->prof_ms[i];
  for (int i = 0; i < 5; i++) counts[i] = ctx->prof_cnt[i];
  return KLNMF_OK;
}

// Diagnostic / test hook: one contraction out = op(A).op(B) through the context-free engine of
// `mode` (the same kernels klnmf_run uses), host float64 in and out.
int klnmf_contract_host(int device, int mode, int64_t M, int64_t N, int64_t K, const double *A, int a_trans,
                        const double *B, int b_trans, double *out) {
  KL_CHECK(A && B && out && M > 0 && N > 0 && K > 0, KLNMF_EINVAL, "contract_host: bad argument");
  klnmf_ctx *ctx = nullptr;
  KL_TRY(klnmf_create(&ctx, device, M, N, K, mode));
  const int64_t es = ctx->es;
  // A: a_trans=0 -> M x K row-major (K contiguous); a_trans=1 -> K x M row-major (M contiguous)
  const int64_t a_rows = a_trans ? K : M, a_cols = a_trans ? M : K;
  const int64_t b_rows = b_trans ? N : K, b_cols = b_trans ? K : N;
  const int64_t lda = round_up(a_cols, 32), ldb = round_up(b_cols, 32), ldo = round_up(N, 32);
  void *dA = nullptr, *dAlo = nullptr, *dB = nullptr, *dBlo = nullptr, *dO = nullptr;
  int rc = dmalloc(&dA, a_rows * lda * es);
  if (rc == KLNMF_OK) rc = dmalloc(&dB, b_rows * ldb * es);
  if (rc == KLNMF_OK) rc = dmalloc(&dO, M * ldo * es);
  if (rc == KLNMF_OK) {
    cudaMemsetAsync(dA, 0, a_rows * lda * es, ctx->stream);
    cudaMemsetAsync(dB, 0, b_rows * ldb * es, ctx->stream);
    cudaMemsetAsync(dO, 0, M * ldo * es, ctx->stream);
    rc = upload_matrix(ctx, A, KLNMF_F64, a_cols, dA, lda, a_rows, a_cols, false);
  }
  if (rc == KLNMF_OK) rc = upload_matrix(ctx, B, KLNMF_F64, b_cols, dB, ldb, b_rows, b_cols, false);
  if (rc == KLNMF_OK && ctx->split) {
    rc = dmalloc(&dAlo, a_rows * lda * es);
    if (rc == KLNMF_OK) rc = dmalloc(&dBlo, b_rows * ldb * es);
    if (rc == KLNMF_OK) {
      cudaMemsetAsync(dAlo, 0, a_rows * lda * es, ctx->stream);
      cudaMemsetAsync(dBlo, 0, b_rows * ldb * es, ctx->stream);
      rc = launch_split(ctx, (const float *)dA, (float *)dA, (float *)dAlo, a_rows, a_cols, lda);
    }
    if (rc == KLNMF_OK) rc = launch_split(ctx, (const float *)dB, (float *)dB, (float *)dBlo, b_rows, b_cols, ldb);
  }
  if (rc == KLNMF_OK) {
    GemmDesc d{};
    d.M = M; d.N = N; d.K = K;
    d.A = dA; d.A_lo = dAlo; d.a_sm = a_trans ? 1 : lda; d.a_sk = a_trans ? lda : 1;
    d.B = dB; d.B_lo = dBlo; d.b_sk = b_trans ? 1 : ldb; d.b_sn = b_trans ? ldb : 1;
    d.out = dO; d.ldo = ldo;
    rc = dense_gemm(ctx, EPI_STORE, d);
  }
  if (rc == KLNMF_OK) rc = download_matrix(ctx, dO, nullptr, ldo, out, KLNMF_F64, N, M, N, false);
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  if (rc == KLNMF_OK && se != cudaSuccess) {
    set_error("contract_host: device error: %s", cudaGetErrorString(se));
    rc = KLNMF_ECUDA;
  }
  void *ptrs[] = {dA, dAlo, dB, dBlo, dO};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  klnmf_destroy(ctx);
  return rc;
}

// Diagnostic: device time of `iters` back-to-back contractions of the given shape and operand majors
// on synthetic device data (no host copies); ms_out = average milliseconds per contraction.
int klnmf_contract_bench(int device, int mode, int64_t M, int64_t N, int64_t K, int a_trans, int b_trans, int iters,
                         double *ms_out) {
  KL_CHECK(ms_out && M > 0 && N > 0 && K > 0 && iters > 0, KLNMF_EINVAL, "contract_bench: bad argument");
  klnmf_ctx *ctx = nullptr;
  KL_TRY(klnmf_create(&ctx, device, M, N, K, mode));
  const int64_t es = ctx->es;
  const int64_t a_rows = a_trans ? K : M, a_cols = a_trans ? M : K;
  const int64_t b_rows = b_trans ? N : K, b_cols = b_trans ? K : N;
  const int64_t lda = round_up(a_cols, 32), ldb = round_up(b_cols, 32), ldo = round_up(N, 32);
  void *dA = nullptr, *dAlo = nullptr, *dB = nullptr, *dBlo = nullptr, *dO = nullptr;
  int rc = dmalloc(&dA, a_rows * lda * es);
  if (rc == KLNMF_OK) rc = dmalloc(&dB, b_rows * ldb * es);
  if (rc == KLNMF_OK) rc = dmalloc(&dO, M * ldo * es);
  if (rc == KLNMF_OK) rc = launch_fill_uniform(ctx, dA, (int)es, a_rows, a_cols, lda, 1);
  if (rc == KLNMF_OK) rc = launch_fill_uniform(ctx, dB, (int)es, b_rows, b_cols, ldb, 2);
  if (rc == KLNMF_OK && ctx->split) {
    rc = dmalloc(&dAlo, a_rows * lda * es);
    if (rc == KLNMF_OK) rc = dmalloc(&dBlo, b_rows * ldb * es);
    if (rc == KLNMF_OK) rc = launch_split(ctx, (const float *)dA, (float *)dA, (float *)dAlo, a_rows, a_cols, lda);
    if (rc == KLNMF_OK) rc = launch_split(ctx, (const float *)dB, (float *)dB, (float *)dBlo, b_rows, b_cols, ldb);
  }
  GemmDesc d{};
  d.M = M; d.N = N; d.K = K;
  d.A = dA; d.A_lo = dAlo; d.a_sm = a_trans ? 1 : lda; d.a_sk = a_trans ? lda : 1;
  d.B = dB; d.B_lo = dBlo; d.b_sk = b_trans ? 1 : ldb; d.b_sn = b_trans ? ldb : 1;
  d.out = dO; d.ldo = ldo;
  // KLNMF_BENCH_EPI=ratio|ratio_kl: time the ratio epilogue (with / without the Q store) instead of a plain store
  int epi = EPI_STORE;
  void *dX = nullptr;
  double *dkl = nullptr;
  const char *be = getenv("KLNMF_BENCH_EPI");
  if (rc == KLNMF_OK && be && strncmp(be, "ratio", 5) == 0) {
    epi = EPI_RATIO;
    rc = dmalloc(&dX, M * ldo * es);
    if (rc == KLNMF_OK) rc = dmalloc((void **)&dkl, 8);
    if (rc == KLNMF_OK) rc = launch_fill_uniform(ctx, dX, (int)es, M, N, ldo, 3);
    if (rc == KLNMF_OK) cudaMemsetAsync(dkl, 0, 8, ctx->stream);
    d.aux = dX; d.ldaux = ldo; d.kl = dkl; d.only_kl = strcmp(be, "ratio_kl") == 0;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 2 && rc == KLNMF_OK; i++) rc = dense_gemm(ctx, epi, d);
  cudaEventRecord(e0, ctx->stream);
  for (int i = 0; i < iters && rc == KLNMF_OK; i++) rc = dense_gemm(ctx, epi, d);
  cudaEventRecord(e1, ctx->stream);
  cudaError_t se = cudaStreamSynchronize(ctx->stream);
  if (rc == KLNMF_OK && se != cudaSuccess) {
    set_error("contract_bench: device error: %s", cudaGetErrorString(se));
    rc = KLNMF_ECUDA;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = (double)ms / iters;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  void *ptrs[] = {dA, dAlo, dB, dBlo, dO, dX, dkl};
  for (void *q : ptrs)
    if (q) cudaFree(q);
  klnmf_destroy(ctx);
  return rc;
}

int klnmf_l2_read_bench(int device, int64_t bytes, int iters, double *gbps_out) {
  KL_CHECK(gbps_out && bytes >= (1 << 20) && iters >= 1, KLNMF_EINVAL, "l2_read_bench: bad argument");
  KL_CHECK(klnmf_device_count() > 0, KLNMF_ENODEVICE, "no CUDA device: libklnmf has no CPU path");
  return l2_read_bench(device, bytes, iters, gbps_out);
}

const char *klnmf_engine_name(klnmf_ctx *ctx) {
  if (!ctx) return "none";
  if (ctx->es == 8) return "dmma_f64";
  if (ctx->debug_simt) return "simt_f32_debug";
  return ctx->mode == KLNMF_MODE_TF32X3 ? "tcgen05_tf32x3" : (ctx->mode == KLNMF_MODE_TF32R ? "tcgen05_tf32r" : "tcgen05_tf32");
}

}  // extern "C"
