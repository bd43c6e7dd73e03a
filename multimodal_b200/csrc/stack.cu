// The learner's stack of modalities when at least one of them is sparse (learner.py:53-56, array_utils.py:5-9):
//     safe_hstack([coef_m * X_m])  ->  scipy.sparse.hstack(...)  ->  CSR,
// i.e. the reference sparsifies the dense modalities, scales every block and concatenates them on the host.  Here the
// blocks are uploaded as they are (dense blocks dense, CSR blocks CSR) and the scaled, stacked CSR matrix is built on
// the device: count the non-zeros of every row over all blocks -> exclusive scan -> fill.  Explicit zeros (and dense
// zeros) are dropped, which is what the reference's eliminate_zeros() does to the stack before its first use
// (nmf.py:66); column indices inside a row come out sorted when the CSR blocks' rows are.
#include <cub/device/device_scan.cuh>

#include <vector>

#include "common.cuh"

namespace klnmf {

namespace {

constexpr int MAX_BLOCKS = 16;

struct DevBlock {
  int kind;                 // 0 dense, 1 CSR
  int es;                   // element size of the block's values (4 / 8)
  int f32prod;              // scale * x formed in float
  int64_t cols, col0, ld;
  double scale;
  const void *dense;        // n x ld
  const int64_t *indptr;
  const int32_t *indices;
  const void *vals;
};
struct DevBlocks {
  int n;
  DevBlock b[MAX_BLOCKS];
};

__device__ __forceinline__ double scaled(const DevBlock &b, const void *base, int64_t idx) {
  if (b.es == 4) {
    const float x = ((const float *)base)[idx];
    return b.f32prod ? (double)(x * (float)b.scale) : (double)x * b.scale;
  }
  return ((const double *)base)[idx] * b.scale;
}

// one warp per row; FILL = false: cnt[i] = stored entries of the stacked row; FILL = true: write them at indptr[i]
template <typename T, bool FILL>
__global__ void __launch_bounds__(256) stack_rows_kernel(const DevBlocks blocks, int64_t n, int64_t *__restrict__ cnt,
                                                          const int64_t *__restrict__ indptr, int32_t *__restrict__ indices,
                                                          T *__restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp_global; i < n; i += n_warps) {
    int64_t pos = FILL ? indptr[i] : 0;
    for (int bi = 0; bi < blocks.n; bi++) {
      const DevBlock &b = blocks.b[bi];
      int64_t lo, hi;
      if (b.kind == 0) { lo = 0; hi = b.cols; } else { lo = b.indptr[i]; hi = b.indptr[i + 1]; }
      for (int64_t t0 = lo; t0 < hi; t0 += 32) {
        const int64_t t = t0 + lane;
        double v = 0.0;
        int32_t col = 0;
        if (t < hi) {
          if (b.kind == 0) { v = scaled(b, b.dense, i * b.ld + t); col = (int32_t)(b.col0 + t); }
          else { v = scaled(b, b.vals, t); col = (int32_t)(b.col0 + b.indices[t]); }
        }
        const bool keep = t < hi && v != 0.0;          // NaN compares unequal to zero and is kept: the input check sees it
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
          const int64_t o = pos + __popc(m & ((1u << lane) - 1u));
          indices[o] = col;
          vals[o] = (T)v;
        }
        pos += __popc(m);
      }
    }
    if (!FILL && lane == 0) cnt[i] = pos;
  }
}

}  // namespace

int stack_blocks_to_csr(klnmf_ctx *ctx, int n_blocks, const klnmf_block *blocks) {
  KL_CHECK(n_blocks >= 1 && n_blocks <= MAX_BLOCKS, KLNMF_EINVAL, "set_stacked_blocks_host: 1..%d blocks", MAX_BLOCKS);
  const int64_t n = ctx->n;
  DevBlocks db{};
  db.n = n_blocks;
  std::vector<void *> temps;
  auto cleanup = [&]() { for (void *p : temps) cudaFree(p); };
  auto fail = [&](int rc) { cudaStreamSynchronize(ctx->stream); cleanup(); return rc; };
  auto up = [&](const void *src, int64_t bytes, void **dst) -> int {
    *dst = nullptr;
    if (cudaMalloc(dst, (size_t)(bytes > 0 ? bytes : 16)) != cudaSuccess) { set_error("set_stacked_blocks_host: cudaMalloc(%lld) failed", (long long)bytes); cudaGetLastError(); return KLNMF_ENOMEM; }
    temps.push_back(*dst);
    if (bytes > 0 && cudaMemcpyAsync(*dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { set_error("set_stacked_blocks_host: upload failed"); return KLNMF_ECUDA; }
    ctx->bytes_h2d += bytes;
    return KLNMF_OK;
  };
  int64_t col0 = 0;
  for (int bi = 0; bi < n_blocks; bi++) {
    const klnmf_block &s = blocks[bi];
    DevBlock &d = db.b[bi];
    KL_CHECK(s.dtype == KLNMF_F32 || s.dtype == KLNMF_F64, KLNMF_EINVAL, "set_stacked_blocks_host: bad dtype of block %d", bi);
    KL_CHECK(s.cols >= 0, KLNMF_EINVAL, "set_stacked_blocks_host: bad width of block %d", bi);
    d.kind = s.kind; d.es = s.dtype == KLNMF_F64 ? 8 : 4; d.f32prod = (s.product_f32 && s.dtype == KLNMF_F32) ? 1 : 0;
    d.cols = s.cols; d.col0 = col0; d.scale = s.scale;
    int rc = KLNMF_OK;
    if (s.kind == 0) {
      if (!(s.ld >= s.cols && (s.dense || n == 0 || s.cols == 0))) { set_error("set_stacked_blocks_host: bad dense block %d", bi); return fail(KLNMF_EINVAL); }
      d.ld = s.cols;
      void *p = nullptr;
      if (cudaMalloc(&p, (size_t)(n * s.cols * d.es > 0 ? n * s.cols * d.es : 16)) != cudaSuccess) { set_error("set_stacked_blocks_host: cudaMalloc failed"); cudaGetLastError(); return fail(KLNMF_ENOMEM); }
      temps.push_back(p);
      if (n * s.cols > 0 && cudaMemcpy2DAsync(p, (size_t)(s.cols * d.es), s.dense, (size_t)(s.ld * d.es), (size_t)(s.cols * d.es), (size_t)n,
                                              cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { set_error("set_stacked_blocks_host: upload failed"); return fail(KLNMF_ECUDA); }
      ctx->bytes_h2d += n * s.cols * d.es;
      d.dense = p;
    } else if (s.kind == 1) {
      if (!(s.indptr && s.nnz >= 0 && (s.nnz == 0 || (s.indices && s.values)) && s.indptr[0] == 0 && s.indptr[n] == s.nnz)) {
        set_error("set_stacked_blocks_host: bad CSR block %d", bi);
        return fail(KLNMF_EINVAL);
      }
      void *p = nullptr;
      if ((rc = up(s.indptr, (n + 1) * 8, &p)) != KLNMF_OK) return fail(rc);
      d.indptr = (const int64_t *)p;
      if ((rc = up(s.indices, s.nnz * 4, &p)) != KLNMF_OK) return fail(rc);
      d.indices = (const int32_t *)p;
      if ((rc = up(s.values, s.nnz * d.es, &p)) != KLNMF_OK) return fail(rc);
      d.vals = p;
    } else {
      set_error("set_stacked_blocks_host: block %d is neither dense (0) nor CSR (1)", bi);
      return fail(KLNMF_EINVAL);
    }
    col0 += s.cols;
  }
  if (col0 != ctx->f) { set_error("set_stacked_blocks_host: the blocks have %lld columns, the context %lld", (long long)col0, (long long)ctx->f); return fail(KLNMF_EINVAL); }
  if (ctx->f >= ((int64_t)1 << 31)) { set_error("set_stacked_blocks_host: int32 column indices need f < 2^31"); return fail(KLNMF_EINVAL); }

  // count -> scan -> fill
  int64_t *cnt = nullptr, *indptr = nullptr;
  void *scan_tmp = nullptr;
  if (cudaMalloc((void **)&cnt, (size_t)(n + 1) * 8) != cudaSuccess) { set_error("set_stacked_blocks_host: cudaMalloc failed"); return fail(KLNMF_ENOMEM); }
  temps.push_back(cnt);
  if (cudaMalloc((void **)&indptr, (size_t)(n + 1) * 8) != cudaSuccess) { set_error("set_stacked_blocks_host: cudaMalloc failed"); return fail(KLNMF_ENOMEM); }
  cudaMemsetAsync(cnt, 0, (size_t)(n + 1) * 8, ctx->stream);
  const int grid = (int)(ceil_div(n > 0 ? n : 1, 8) < (int64_t)ctx->sm_count * 16 ? ceil_div(n > 0 ? n : 1, 8) : (int64_t)ctx->sm_count * 16);
  if (n > 0) {
    if (ctx->es == 8) stack_rows_kernel<double, false><<<grid, 256, 0, ctx->stream>>>(db, n, cnt, nullptr, nullptr, nullptr);
    else stack_rows_kernel<float, false><<<grid, 256, 0, ctx->stream>>>(db, n, cnt, nullptr, nullptr, nullptr);
    ctx->n_launch++;
  }
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, indptr, (int64_t)(n + 1), ctx->stream);
  if (cudaMalloc(&scan_tmp, tmp_bytes > 0 ? tmp_bytes : 16) != cudaSuccess) { cudaFree(indptr); set_error("set_stacked_blocks_host: cudaMalloc failed"); return fail(KLNMF_ENOMEM); }
  temps.push_back(scan_tmp);
  cub::DeviceScan::ExclusiveSum(scan_tmp, tmp_bytes, cnt, indptr, (int64_t)(n + 1), ctx->stream);
  ctx->n_launch++;
  int64_t nnz = 0;
  if (cudaMemcpyAsync(&nnz, indptr + n, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) { cudaFree(indptr); set_error("set_stacked_blocks_host: count failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(KLNMF_ECUDA); }
  ctx->nnz = nnz;
  ctx->indptr = indptr;
  ctx->csr_owned = true;
  ctx->indices = nullptr; ctx->vals = nullptr;
  if (cudaMalloc((void **)&ctx->indices, (size_t)(nnz > 0 ? nnz : 1) * 4) != cudaSuccess ||
      cudaMalloc(&ctx->vals, (size_t)(nnz > 0 ? nnz : 1) * ctx->es) != cudaSuccess) { set_error("set_stacked_blocks_host: cudaMalloc failed"); cudaGetLastError(); return fail(KLNMF_ENOMEM); }
  if (n > 0 && nnz > 0) {
    if (ctx->es == 8) stack_rows_kernel<double, true><<<grid, 256, 0, ctx->stream>>>(db, n, nullptr, indptr, ctx->indices, (double *)ctx->vals);
    else stack_rows_kernel<float, true><<<grid, 256, 0, ctx->stream>>>(db, n, nullptr, indptr, ctx->indices, (float *)ctx->vals);
    ctx->n_launch++;
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("set_stacked_blocks_host: device error: %s", cudaGetErrorString(e)); return KLNMF_ECUDA; }
  return KLNMF_OK;
}

}  // namespace klnmf
