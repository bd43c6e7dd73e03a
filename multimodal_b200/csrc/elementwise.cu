// Vectorised / warp-reduction kernels around the contractions:
// dictionary update + row normalisation (nmf.py:345-351, array_utils.py:19-22), the
// device-side stopping test (nmf.py:214-220), split-TF32 operand preparation, layout
// conversion, input checks, synthetic data.
#include "common.cuh"

namespace klnmf {

namespace {

__device__ __forceinline__ float tf32_hi(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

__device__ __forceinline__ double block_sum(double v, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (warp == 0) {
    s = lane < nw ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[0] = s;
  }
  __syncthreads();
  s = red[0];
  __syncthreads();
  return s;
}

// H_new[a,:] = H[a,:]*N[a,:] / (1e-16 + sum_j H[a,j]*N[a,j]); one block per dictionary row.
template <typename T>
__global__ void __launch_bounds__(512) dict_update_kernel(const T *__restrict__ H, const T *__restrict__ Hlo,
                                                          const T *__restrict__ num, T *__restrict__ Hn,
                                                          T *__restrict__ Hnlo, int64_t f, int64_t ld,
                                                          double *__restrict__ rowsum, const double *__restrict__ rowadd,
                                                          const int *stop, int64_t cc, int64_t krows) {
  if (*stop != 0) return;
  __shared__ double red[32];
  const int64_t a = blockIdx.x;
  const T *h = H + a * ld;
  // numerator layout: plain k x ld (cc == 0) or f-chunks [chunk][krows][cc] (klnmf_ctx::num_cc)
  auto nmv = [&](int64_t j) -> T { return cc ? num[(j / cc) * (krows * cc) + a * cc + (j % cc)] : num[a * ld + j]; };
  const T *hl = Hlo ? Hlo + a * ld : nullptr;
  // centered ratio (api.cu, dense_iteration): num holds W'^T.(Q - 1); the missing W'^T.1 = colsum(W') is the same
  // for every feature of a dictionary row
  const T add = rowadd ? (T)rowadd[a] : (T)0;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < f; j += blockDim.x) {
    T v = h[j];
    if (hl) v += hl[j];
    const T nj = nmv(j);
    s += (double)(v * (rowadd ? fmax(nj + add, (T)0) : nj));
  }
  s = block_sum(s, red);
  const double inv = 1.0 / (KL_NORM_EPS + s);
  for (int64_t j = threadIdx.x; j < f; j += blockDim.x) {
    T v = h[j];
    if (hl) v += hl[j];
    const T nj = nmv(j);
    T r = (T)((double)(v * (rowadd ? fmax(nj + add, (T)0) : nj)) * inv);   // the centered numerator, clamped like G
    if (Hnlo) {
      float hi = tf32_hi((float)r);
      Hn[a * ld + j] = (T)hi;
      Hnlo[a * ld + j] = (T)((float)r - hi);
    } else {
      Hn[a * ld + j] = r;
    }
  }
  if (threadIdx.x == 0 && rowsum) rowsum[a] = s * inv;
}

// transposed (sparse-path) layout: Ht is f x k.  Pass 1: column sums of Ht*Nt.
template <typename T>
__global__ void __launch_bounds__(256) dict_colsum_t_kernel(const T *__restrict__ Ht, const T *__restrict__ Nt,
                                                            int64_t f, int64_t k, int64_t ld,
                                                            double *__restrict__ hsum, int rows_per_block,
                                                            const int *stop) {
  if (*stop != 0) return;
  const int64_t j0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t j1 = j0 + rows_per_block < f ? j0 + rows_per_block : f;
  for (int64_t a = threadIdx.x; a < k; a += blockDim.x) {
    double s = 0.0;
    for (int64_t j = j0; j < j1; j++) s += (double)(Ht[j * ld + a] * Nt[j * ld + a]);
    atomicAdd(&hsum[a], s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) dict_scale_t_kernel(const T *__restrict__ Ht, const T *__restrict__ Nt,
                                                           T *__restrict__ Htn, int64_t f, int64_t k, int64_t ld,
                                                           const double *__restrict__ hsum,
                                                           double *__restrict__ rowsum, const int *stop,
                                                           const double *__restrict__ rs_part) {
  // rs_part (hybrid stacks): the share of hsum that belongs to THESE columns -- rowsum then is the row sum of this
  // block of the dictionary only, which is what the sparse objective multiplies colsum(W) with
  if (*stop != 0) return;
  const int64_t total = f * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t j = i / k, a = i - j * k;
    double inv = 1.0 / (KL_NORM_EPS + hsum[a]);
    Htn[j * ld + a] = (T)((double)(Ht[j * ld + a] * Nt[j * ld + a]) * inv);
    if (j == 0) rowsum[a] = (rs_part ? rs_part[a] : hsum[a]) * inv;
  }
}

// ---- hybrid stacks (a dense block next to the CSR block, api.cu: HybridSide) -------------------------------------
// structural zeros of the stack carry no ratio: the reference sparsifies the whole stack (array_utils.py:5-9) and
// forms the ratio at its stored entries only (nmf.py:52-70, 332-336)
// (round_q, FP32 only: the ratio is also rounded to nearest TF32, so that the tensor core, which would truncate it,
// multiplies exactly what is stored -- the one-pass modes' treatment of every operand)
__device__ __forceinline__ float round_operand(float v) { return tf32_hi(v); }
__device__ __forceinline__ double round_operand(double v) { return v; }
template <typename T>
__global__ void __launch_bounds__(256) mask_ratio_kernel(T *__restrict__ Q, int64_t ldq, const T *__restrict__ X, int64_t ldx,
                                                         int64_t rows, int64_t cols, const int *stop, int round_q) {
  if (stop != nullptr && *stop != 0) return;
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    if (X[r * ldx + c] == (T)0) Q[r * ldq + c] = (T)0;
    else if (round_q) Q[r * ldq + c] = round_operand(Q[r * ldq + c]);
  }
}
// hsum_d[a] = sum_j H[a,j] N[a,j] over the dense block's columns (k x f layout), one CTA per component
template <typename T>
__global__ void __launch_bounds__(256) hyb_rowsum_kernel(const T *__restrict__ H, const T *__restrict__ N, int64_t f, int64_t ld,
                                                         double *__restrict__ hsum_d, const int *stop) {
  if (*stop != 0) return;
  __shared__ double red[32];
  const int64_t a = blockIdx.x;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < f; j += blockDim.x) s += (double)(H[a * ld + j] * N[a * ld + j]);
  s = block_sum(s, red);
  if (threadIdx.x == 0) hsum_d[a] = s;
}
__global__ void hyb_total_kernel(double *__restrict__ hsum_d, const double *__restrict__ hsum_s, int64_t k, const int *stop) {
  if (*stop != 0) return;
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < k; a += (int64_t)gridDim.x * blockDim.x)
    hsum_d[a] += hsum_s[a];
}
// H'[a,j] = H[a,j] N[a,j] / (1e-16 + total[a])  (array_utils.py:19-22 over the WHOLE row of the stacked dictionary)
template <typename T>
__global__ void __launch_bounds__(256) hyb_scale_kernel(const T *__restrict__ H, const T *__restrict__ N, T *__restrict__ Hn,
                                                        int64_t k, int64_t f, int64_t ld, const double *__restrict__ total,
                                                        const int *stop) {
  if (*stop != 0) return;
  const int64_t cells = k * f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = i / f, j = i - a * f;
    const double inv = 1.0 / (KL_NORM_EPS + total[a]);
    Hn[a * ld + j] = (T)((double)(H[a * ld + j] * N[a * ld + j]) * inv);
  }
}

// The reference's stopping test on the device (nmf.py:214-220).  dred = [kl, sum(X.data), colsum(W)...]
__global__ void decide_kernel(double *dred, double *dscal, int *flags, double *errors, int errors_cap,
                              const double *rowsumH, int64_t k, int sparse, double *colsum_out) {
  // one warp: the k-long loops are strided over its lanes (one thread walking k = 512 global values took 90 us)
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  double wh = 0.0;
  if (sparse)
    for (int64_t a = lane; a < k; a += 32) wh += dred[2 + a] * rowsumH[a];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wh += __shfl_xor_sync(0xffffffffu, wh, o);
  if (lane == 0) {
    double e = dred[0];
    if (sparse) e = e - dred[1] + wh;             // nmf.py:301-308
    if (flags[FL_STOP] == 0) {
      const double prev = dscal[DS_PREV];
      // DS_WHSUM holds the relative objective noise of the arithmetic mode (see klnmf_run): a tolerance at or below that
      // noise (tol == 0 above all) asks "did the objective rise", and a rise has to exceed the noise to count
      const double noise = dscal[DS_WHSUM] * (isfinite(prev) ? fabs(prev) : 0.0);
      double tol = dscal[DS_TOL];
      if (tol <= noise) tol -= noise;
      if (prev - e < tol) {
        flags[FL_STOP] = 1;
      } else {
        dscal[DS_PREV] = e;
        int ne = flags[FL_NERR];
        if (ne < errors_cap) errors[ne] = e;
        flags[FL_NERR] = ne + 1;
      }
    }
    dscal[DS_KL] = e;
    dred[0] = 0.0;
    dred[1] = dscal[DS_SUMX];
  }
  __syncwarp();
  for (int64_t a = lane; a < k; a += 32) {
    if (colsum_out) colsum_out[a] = dred[2 + a];     // colsum(W') of all ranks, for the dictionary update
    dred[2 + a] = 0.0;
  }
}

template <typename T>
__global__ void scale_block_kernel(T *__restrict__ p, int64_t rows, int64_t cols, int64_t ld, double scale, int f32) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    p[r * ld + c] = f32 ? (T)((float)p[r * ld + c] * (float)scale) : (T)((double)p[r * ld + c] * scale);
  }
}

__global__ void rsh32_kernel(const double *__restrict__ rowsumH, float *__restrict__ out, int64_t k, int64_t len) {
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < len; a += (int64_t)gridDim.x * blockDim.x)
    out[a] = a < k ? (float)rowsumH[a] : 0.f;
}

// out[a] += sum_i W[i,a] (+ Wlo): a warp reads whole rows (up to 512 columns per sweep), FP64 accumulation
template <typename T>
__global__ void __launch_bounds__(256) colsum_w_kernel(const T *__restrict__ W, const T *__restrict__ Wlo, int64_t n,
                                                       int64_t k, int64_t ld, double *__restrict__ out, const int *stop) {
  if (*stop != 0) return;
  __shared__ double red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int64_t cb = 0; cb < k; cb += 512) {
    double acc[16];
#pragma unroll
    for (int m = 0; m < 16; m++) acc[m] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * 8 + ty; i < n; i += (int64_t)gridDim.x * 8) {
      const T *w = W + i * ld + cb;
      const T *wl = Wlo ? Wlo + i * ld + cb : nullptr;
#pragma unroll
      for (int m = 0; m < 16; m++) {
        const int64_t c = cb + 32 * m + tx;
        if (c < k) {
          acc[m] += (double)w[32 * m + tx];
          if (wl) acc[m] += (double)wl[32 * m + tx];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 16; m++) {
      const int64_t c = cb + 32 * m + tx;
      red[ty][tx] = acc[m];
      __syncthreads();
      if (ty == 0 && c < k) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < 8; r++) t += red[r][tx];
        atomicAdd(&out[c], t);
      }
      __syncthreads();
    }
  }
}

// The same sums for FP32 state (the one-pass and split modes run it once per fit iteration over all of W': 16 GB at cfg5).
// The generic kernel above converts and adds every element in FP64 (4-byte loads, F2F + DADD per element: 2.4 TB/s,
// 6.8 ms per iteration at cfg5 under the power cap, outside every timed phase).  Here a thread owns one float4 column
// group, adds hi + lo of 32 rows in FP32 (non-negative terms: <= 32 * 2^-24 relative) and only then in FP64.
__global__ void __launch_bounds__(256) colsum_w_f32_kernel(const float *__restrict__ W, const float *__restrict__ Wlo, int64_t n,
                                                           int64_t ld, double *__restrict__ out, const int *stop) {
  if (*stop != 0) return;
  const int G = (int)(ld >> 2);                       // float4 groups per row (ld is a multiple of 32)
  const int per = 256 / G > 0 ? 256 / G : 1;          // rows per sweep of the CTA (G <= 256)
  const int g = threadIdx.x % G, r0 = threadIdx.x / G;
  extern __shared__ double4 part[];                   // per x G partial sums
  double4 acc = make_double4(0.0, 0.0, 0.0, 0.0);
  if (r0 < per) {
    const int64_t rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo_row = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t hi_row = lo_row + rows_per_cta < n ? lo_row + rows_per_cta : n;
    for (int64_t base = lo_row + r0; base < hi_row; base += (int64_t)per * 32) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int t = 0; t < 32; t++) {
        const int64_t i = base + (int64_t)t * per;
        if (i < hi_row) {
          const float4 v = __ldg(reinterpret_cast<const float4 *>(W + i * ld) + g);
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          if (Wlo) {
            const float4 l = __ldg(reinterpret_cast<const float4 *>(Wlo + i * ld) + g);
            a.x += l.x; a.y += l.y; a.z += l.z; a.w += l.w;
          }
        }
      }
      acc.x += (double)a.x; acc.y += (double)a.y; acc.z += (double)a.z; acc.w += (double)a.w;
    }
    part[r0 * G + g] = acc;
  }
  __syncthreads();
  if (r0 == 0) {
    for (int r = 1; r < per; r++) {
      const double4 o = part[r * G + g];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    atomicAdd(&out[4 * g], acc.x); atomicAdd(&out[4 * g + 1], acc.y);
    atomicAdd(&out[4 * g + 2], acc.z); atomicAdd(&out[4 * g + 3], acc.w);
  }
}

__global__ void split_kernel(const float *__restrict__ src, float *__restrict__ hi, float *__restrict__ lo,
                             int64_t rows, int64_t cols, int64_t ld) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    float v = src[r * ld + c];
    float h = tf32_hi(v);
    hi[r * ld + c] = h;
    if (lo) lo[r * ld + c] = v - h;
  }
}

template <typename S, typename D>
__global__ void convert_kernel(const S *__restrict__ src, const S *__restrict__ src_lo, int64_t src_ld,
                               D *__restrict__ dst, int64_t dst_ld, int64_t rows, int64_t cols, int transpose,
                               double scale, int scale_f32) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    D v = scale == 1.0 ? (D)src[r * src_ld + c]
                       : (scale_f32 ? (D)((float)src[r * src_ld + c] * (float)scale)           // numpy's float32 product
                                    : (D)((double)src[r * src_ld + c] * scale));             // float64 product, one rounding
    if (src_lo) v += (D)src_lo[r * src_ld + c];
    if (transpose) dst[c * dst_ld + r] = v; else dst[r * dst_ld + c] = v;
  }
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename T>
__global__ void fill_uniform_kernel(T *__restrict__ p, int64_t rows, int64_t cols, int64_t ld, uint64_t seed,
                                    int64_t row_offset) {
  const int64_t total = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    uint64_t h = mix64(seed ^ mix64((uint64_t)(r + row_offset) * 0x100000001B3ull + (uint64_t)c));
    // uniform in (0, 1]
    p[r * ld + c] = (T)(((double)(h >> 11) + 1.0) * (1.0 / 9007199254740992.0));
  }
}

template <typename T>
__global__ void check_kernel(const T *__restrict__ p, int64_t rows, int64_t cols, int64_t ld, int *flags) {
  const int64_t total = rows * cols;
  int neg = 0, bad = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / cols, c = i - r * cols;
    T v = p[r * ld + c];
    neg |= (v < (T)0);
    bad |= !isfinite((double)v);
  }
  if (neg) atomicOr(&flags[FL_NEG], 1);
  if (bad) atomicOr(&flags[FL_NONFINITE], 1);
}

// rowsumH[a] = sum_j H[a,j]  (dense layout: one block per row; transposed: strided)
template <typename T>
__global__ void __launch_bounds__(256) rowsum_kernel(const T *__restrict__ H, const T *__restrict__ Hlo,
                                                     int64_t k, int64_t f, int64_t ld, int transposed,
                                                     double *__restrict__ out) {
  __shared__ double red[32];
  const int64_t a = blockIdx.x;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < f; j += blockDim.x) {
    int64_t off = transposed ? j * ld + a : a * ld + j;
    T v = H[off];
    if (Hlo) v += Hlo[off];
    s += (double)v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[a] = s;
}

template <typename T>
__global__ void __launch_bounds__(256) sum_vals_kernel(const T *__restrict__ v, int64_t n, double *out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += (double)v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

inline int grid_for(klnmf_ctx *ctx, int64_t total, int threads) {
  int64_t g = ceil_div(total, threads);
  int64_t cap = (int64_t)ctx->sm_count * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int launch_dict_update(klnmf_ctx *ctx, const void *H_old, void *H_new, void *Hlo_new, const double *rowadd) {
  const int cur = ctx->hcur;
  if (ctx->es == 8) {
    dict_update_kernel<double><<<(unsigned)ctx->k, 512, 0, ctx->stream>>>(
        (const double *)H_old, nullptr, (const double *)ctx->num, (double *)H_new, nullptr, ctx->f, ctx->ldh,
        ctx->rowsumH, rowadd, ctx->flags + FL_STOP, ctx->num_cc, ctx->k);
  } else {
    dict_update_kernel<float><<<(unsigned)ctx->k, 512, 0, ctx->stream>>>(
        (const float *)H_old, (const float *)ctx->Hlo[cur], (const float *)ctx->num, (float *)H_new,
        (float *)Hlo_new, ctx->f, ctx->ldh, ctx->rowsumH, rowadd, ctx->flags + FL_STOP, ctx->num_cc, ctx->k);
  }
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_dict_update_t(klnmf_ctx *ctx, const void *Ht_old, void *Ht_new) {
  KL_CUDA(cudaMemsetAsync(ctx->hsum, 0, sizeof(double) * ctx->k, ctx->stream));
  const int rpb = 64;
  const unsigned g1 = (unsigned)ceil_div(ctx->f, rpb);
  const int g2 = grid_for(ctx, ctx->f * ctx->k, 256);
  if (ctx->es == 8) {
    dict_colsum_t_kernel<double><<<g1, 256, 0, ctx->stream>>>((const double *)Ht_old, (const double *)ctx->num,
                                                              ctx->f, ctx->k, ctx->ldh, ctx->hsum, rpb,
                                                              ctx->flags + FL_STOP);
    dict_scale_t_kernel<double><<<g2, 256, 0, ctx->stream>>>((const double *)Ht_old, (const double *)ctx->num,
                                                             (double *)Ht_new, ctx->f, ctx->k, ctx->ldh, ctx->hsum,
                                                             ctx->rowsumH, ctx->flags + FL_STOP, nullptr);
  } else {
    dict_colsum_t_kernel<float><<<g1, 256, 0, ctx->stream>>>((const float *)Ht_old, (const float *)ctx->num,
                                                             ctx->f, ctx->k, ctx->ldh, ctx->hsum, rpb,
                                                             ctx->flags + FL_STOP);
    dict_scale_t_kernel<float><<<g2, 256, 0, ctx->stream>>>((const float *)Ht_old, (const float *)ctx->num,
                                                            (float *)Ht_new, ctx->f, ctx->k, ctx->ldh, ctx->hsum,
                                                            ctx->rowsumH, ctx->flags + FL_STOP, nullptr);
  }
  ctx->n_launch += 2;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_mask_ratio(klnmf_ctx *ctx, void *Q, int64_t ldq, const void *X, int64_t ldx, int64_t rows, int64_t cols,
                      const int *stop, int round_q) {
  if (rows * cols == 0) return KLNMF_OK;
  const int g = grid_for(ctx, rows * cols, 256);
  if (ctx->es == 8) mask_ratio_kernel<double><<<g, 256, 0, ctx->stream>>>((double *)Q, ldq, (const double *)X, ldx, rows, cols, stop, 0);
  else mask_ratio_kernel<float><<<g, 256, 0, ctx->stream>>>((float *)Q, ldq, (const float *)X, ldx, rows, cols, stop, round_q);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

// The dictionary update of a hybrid stack (nmf.py:345-351): one normaliser per component over BOTH blocks.
// Sparse block: Ht (f_s x k, numerator ctx->num), dense block: Hd (k x fd, numerator Nd).  total = scratch of k doubles.
int launch_dict_update_hybrid(klnmf_ctx *ctx, const void *Ht_old, void *Ht_new, const void *Hd_old, void *Hd_new,
                              const void *Nd, int64_t fd, int64_t ldhd, double *total) {
  const int *stop = ctx->flags + FL_STOP;
  KL_CUDA(cudaMemsetAsync(ctx->hsum, 0, sizeof(double) * ctx->k, ctx->stream));
  const int rpb = 64;
  const unsigned g1 = (unsigned)ceil_div(ctx->f, rpb);
  const int g2 = grid_for(ctx, ctx->f * ctx->k, 256), g3 = grid_for(ctx, ctx->k * fd, 256);
  const unsigned gk = (unsigned)ceil_div(ctx->k, 256);
  if (ctx->es == 8) {
    hyb_rowsum_kernel<double><<<(unsigned)ctx->k, 256, 0, ctx->stream>>>((const double *)Hd_old, (const double *)Nd, fd, ldhd, total, stop);
    dict_colsum_t_kernel<double><<<g1, 256, 0, ctx->stream>>>((const double *)Ht_old, (const double *)ctx->num, ctx->f, ctx->k,
                                                              ctx->ldh, ctx->hsum, rpb, stop);
    hyb_total_kernel<<<gk, 256, 0, ctx->stream>>>(total, ctx->hsum, ctx->k, stop);
    dict_scale_t_kernel<double><<<g2, 256, 0, ctx->stream>>>((const double *)Ht_old, (const double *)ctx->num, (double *)Ht_new,
                                                             ctx->f, ctx->k, ctx->ldh, total, ctx->rowsumH, stop, ctx->hsum);
    hyb_scale_kernel<double><<<g3, 256, 0, ctx->stream>>>((const double *)Hd_old, (const double *)Nd, (double *)Hd_new, ctx->k, fd,
                                                          ldhd, total, stop);
  } else {
    hyb_rowsum_kernel<float><<<(unsigned)ctx->k, 256, 0, ctx->stream>>>((const float *)Hd_old, (const float *)Nd, fd, ldhd, total, stop);
    dict_colsum_t_kernel<float><<<g1, 256, 0, ctx->stream>>>((const float *)Ht_old, (const float *)ctx->num, ctx->f, ctx->k,
                                                             ctx->ldh, ctx->hsum, rpb, stop);
    hyb_total_kernel<<<gk, 256, 0, ctx->stream>>>(total, ctx->hsum, ctx->k, stop);
    dict_scale_t_kernel<float><<<g2, 256, 0, ctx->stream>>>((const float *)Ht_old, (const float *)ctx->num, (float *)Ht_new,
                                                            ctx->f, ctx->k, ctx->ldh, total, ctx->rowsumH, stop, ctx->hsum);
    hyb_scale_kernel<float><<<g3, 256, 0, ctx->stream>>>((const float *)Hd_old, (const float *)Nd, (float *)Hd_new, ctx->k, fd, ldhd,
                                                         total, stop);
  }
  ctx->n_launch += 5;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_decide(klnmf_ctx *ctx, int /*iter_index*/) {
  decide_kernel<<<1, 32, 0, ctx->stream>>>(ctx->dred, ctx->dscal, ctx->flags, ctx->errors_dev, ctx->errors_cap,
                                           ctx->rowsumH, ctx->k, ctx->sparse ? 1 : 0, ctx->colsumW);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_scale_block(klnmf_ctx *ctx, void *X, int64_t ld, int64_t rows, int64_t cols, double scale, int f32) {
  if (rows * cols == 0 || scale == 1.0) return KLNMF_OK;
  const int g = grid_for(ctx, rows * cols, 256);
  if (ctx->es == 8) scale_block_kernel<double><<<g, 256, 0, ctx->stream>>>((double *)X, rows, cols, ld, scale, f32);
  else scale_block_kernel<float><<<g, 256, 0, ctx->stream>>>((float *)X, rows, cols, ld, scale, f32);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_rsh32(klnmf_ctx *ctx) {
  const int64_t len = ctx->ldw + 32;
  rsh32_kernel<<<(unsigned)ceil_div(len, 256), 256, 0, ctx->stream>>>(ctx->rowsumH, ctx->rsh32, ctx->k, len);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_colsum_w(klnmf_ctx *ctx, const void *W, const void *Wlo, double *out) {
  if (ctx->n == 0) return KLNMF_OK;
  const int64_t want = ceil_div(ctx->n, 8 * 16);
  const int grid = (int)(want < (int64_t)ctx->sm_count * 4 ? (want > 0 ? want : 1) : (int64_t)ctx->sm_count * 4);
  if (ctx->es == 8)
    colsum_w_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double *)W, (const double *)Wlo, ctx->n, ctx->k, ctx->ldw, out,
                                                           ctx->flags + FL_STOP);
  else if (ctx->ldw <= 1024 && !(getenv("KLNMF_COLSUM_GENERIC") && atoi(getenv("KLNMF_COLSUM_GENERIC")) == 1)) {
    // (columns k .. ldw of W are zero padding: their sums land in the padding of `out`, which has ldw + 32 slots)
    const int G = (int)(ctx->ldw >> 2), per = 256 / G > 0 ? 256 / G : 1;
    const int64_t want2 = ceil_div(ctx->n, (int64_t)per * 32);
    const int grid2 = (int)(want2 < (int64_t)ctx->sm_count * 8 ? (want2 > 0 ? want2 : 1) : (int64_t)ctx->sm_count * 8);
    colsum_w_f32_kernel<<<grid2, 256, (size_t)per * G * sizeof(double4), ctx->stream>>>((const float *)W, (const float *)Wlo, ctx->n,
                                                                                      ctx->ldw, out, ctx->flags + FL_STOP);
  } else
    colsum_w_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float *)W, (const float *)Wlo, ctx->n, ctx->k, ctx->ldw, out,
                                                          ctx->flags + FL_STOP);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

// measurement support: sustained L2 -> SM read bandwidth (a buffer that fits the L2, read many times by every SM with
// 16-byte loads) -- the denominator of the sparse path's L2 roofline (bench.py), whose gathers never leave the L2
__global__ void l2_read_kernel(const float4 *__restrict__ p, int64_t n4, int iters, float *__restrict__ sink) {
  // Shaped like the gathers it is the roofline of: a warp reads 512 contiguous bytes (16 per lane) at EIGHT places of
  // the buffer before it consumes any of them -- eight independent 16-byte loads in flight per thread.  (One load per
  // thread and loop trip, the first version, measured the latency of a load, not the bandwidth of the L2: 12-16 TB/s
  // where the numerator pass of sparse.cu sustains 18-19 TB/s.)
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t n_seg = (uint32_t)(n4 / 32);        // 512-byte segments (the host keeps n_seg > 8 * n_warps)
  const uint32_t hop = n_seg / 8;                    // the eight places of one trip are an eighth of the buffer apart
  const uint32_t trips = (n_seg + n_warps * 8 - 1) / (n_warps * 8);
  for (int it = 0; it < iters; it++) {
    // a different starting segment per pass keeps the L1 out of it (every SM walks the whole buffer)
    uint32_t s0 = (warp + (uint32_t)it * 4099u) % n_seg;
    for (uint32_t t = 0; t < trips; t++) {
      float4 v[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        uint32_t seg = s0 + e * hop;
        seg = seg >= n_seg ? seg - n_seg : seg;
        asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[e].x), "=f"(v[e].y), "=f"(v[e].z), "=f"(v[e].w)
                     : "l"(p + (uint64_t)(seg * 32u + lane)));
      }
#pragma unroll
      for (int e = 0; e < 8; e++) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
      s0 += n_warps;
      s0 = s0 >= n_seg ? s0 - n_seg : s0;
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;      // never true: keeps the loads alive
}

int l2_read_bench(int device, int64_t bytes, int iters, double *gbps) {
  KL_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  KL_CUDA(cudaGetDeviceProperties(&prop, device));
  float4 *buf = nullptr;
  float *sink = nullptr;
  const int64_t n4 = bytes / 16;
  KL_CHECK(n4 > (int64_t)prop.multiProcessorCount * 8 * 256 && n4 < ((int64_t)1 << 31), KLNMF_EINVAL, "l2_read_bench: 1 MB <= buffer < 32 GB");
  KL_CUDA(cudaMalloc((void **)&buf, (size_t)n4 * 16));
  KL_CUDA(cudaMalloc((void **)&sink, 16));
  cudaMemset(buf, 0, (size_t)n4 * 16);
  const int grid = prop.multiProcessorCount * 4, block = 256;
  l2_read_kernel<<<grid, block>>>(buf, n4, 2, sink);                     // warm the L2
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  l2_read_kernel<<<grid, block>>>(buf, n4, iters, sink);
  cudaEventRecord(e1);
  cudaError_t se = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
  if (se != cudaSuccess) { set_error("l2_read_bench: %s", cudaGetErrorString(se)); return KLNMF_ECUDA; }
  // every pass makes ceil(n_seg / (8 n_warps)) trips of 8 x 512 bytes per warp
  const int64_t n_warps = (int64_t)grid * block / 32, n_seg = n4 / 32;
  const double read = (double)((n_seg + n_warps * 8 - 1) / (n_warps * 8)) * (double)(n_warps * 8) * 512.0 * iters;
  *gbps = read / (ms * 1e-3) / 1e9;
  return KLNMF_OK;
}

int launch_split(klnmf_ctx *ctx, const float *src, float *hi, float *lo, int64_t rows, int64_t cols, int64_t ld) {
  if (rows * cols == 0) return KLNMF_OK;
  split_kernel<<<grid_for(ctx, rows * cols, 256), 256, 0, ctx->stream>>>(src, hi, lo, rows, cols, ld);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_convert(klnmf_ctx *ctx, const void *src, const void *src_lo, int src_dtype, int64_t src_ld, void *dst,
                   int dst_es, int64_t dst_ld, int64_t rows, int64_t cols, bool transpose, double scale, int scale_f32) {
  if (rows * cols == 0) return KLNMF_OK;
  const int g = grid_for(ctx, rows * cols, 256);
  const int t = transpose ? 1 : 0;
  if (src_dtype == KLNMF_F32 && dst_es == 4)
    convert_kernel<float, float><<<g, 256, 0, ctx->stream>>>((const float *)src, (const float *)src_lo, src_ld, (float *)dst, dst_ld, rows, cols, t, scale, scale_f32);
  else if (src_dtype == KLNMF_F32 && dst_es == 8)
    convert_kernel<float, double><<<g, 256, 0, ctx->stream>>>((const float *)src, (const float *)src_lo, src_ld, (double *)dst, dst_ld, rows, cols, t, scale, scale_f32);
  else if (src_dtype == KLNMF_F64 && dst_es == 4)
    convert_kernel<double, float><<<g, 256, 0, ctx->stream>>>((const double *)src, (const double *)src_lo, src_ld, (float *)dst, dst_ld, rows, cols, t, scale, scale_f32);
  else
    convert_kernel<double, double><<<g, 256, 0, ctx->stream>>>((const double *)src, (const double *)src_lo, src_ld, (double *)dst, dst_ld, rows, cols, t, scale, scale_f32);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_zero(klnmf_ctx *ctx, void *p, int64_t bytes) {
  if (bytes > 0) KL_CUDA(cudaMemsetAsync(p, 0, (size_t)bytes, ctx->stream));
  return KLNMF_OK;
}

int launch_fill_uniform(klnmf_ctx *ctx, void *p, int es, int64_t rows, int64_t cols, int64_t ld, uint64_t seed) {
  if (rows * cols == 0) return KLNMF_OK;
  const int g = grid_for(ctx, rows * cols, 256);
  const int64_t row_offset = 0;
  if (es == 8) fill_uniform_kernel<double><<<g, 256, 0, ctx->stream>>>((double *)p, rows, cols, ld, seed, row_offset);
  else fill_uniform_kernel<float><<<g, 256, 0, ctx->stream>>>((float *)p, rows, cols, ld, seed, row_offset);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_check(klnmf_ctx *ctx, const void *p, int es, int64_t rows, int64_t cols, int64_t ld) {
  if (rows * cols == 0) return KLNMF_OK;
  const int g = grid_for(ctx, rows * cols, 256);
  if (es == 8) check_kernel<double><<<g, 256, 0, ctx->stream>>>((const double *)p, rows, cols, ld, ctx->flags);
  else check_kernel<float><<<g, 256, 0, ctx->stream>>>((const float *)p, rows, cols, ld, ctx->flags);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_rowsum_h(klnmf_ctx *ctx) {
  const int cur = ctx->hcur;
  const int t = ctx->sparse ? 1 : 0;
  if (ctx->es == 8)
    rowsum_kernel<double><<<(unsigned)ctx->k, 256, 0, ctx->stream>>>((const double *)ctx->H[cur], nullptr, ctx->k,
                                                                     ctx->f, ctx->ldh, t, ctx->rowsumH);
  else
    rowsum_kernel<float><<<(unsigned)ctx->k, 256, 0, ctx->stream>>>((const float *)ctx->H[cur],
                                                                    (const float *)ctx->Hlo[cur], ctx->k, ctx->f,
                                                                    ctx->ldh, t, ctx->rowsumH);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int launch_sum_vals(klnmf_ctx *ctx) {
  KL_CUDA(cudaMemsetAsync(ctx->dscal + DS_SUMX, 0, sizeof(double), ctx->stream));
  if (ctx->nnz > 0) {
    const int g = grid_for(ctx, ctx->nnz, 256);
    if (ctx->es == 8) sum_vals_kernel<double><<<g, 256, 0, ctx->stream>>>((const double *)ctx->vals, ctx->nnz, ctx->dscal + DS_SUMX);
    else sum_vals_kernel<float><<<g, 256, 0, ctx->stream>>>((const float *)ctx->vals, ctx->nnz, ctx->dscal + DS_SUMX);
    ctx->n_launch++;
  }
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

}  // namespace klnmf
