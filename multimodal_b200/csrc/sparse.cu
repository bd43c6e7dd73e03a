// Sparse (CSR bag-of-features) path of one reference iteration.
//   pass 1, one warp per sample row i (nmf.py:52-70, 297-310, 325-343):
//        s_p = W[i,:] . H[:,j_p]           SDDMM at the stored non-zeros
//        q_p = (x_p+eps)/(s_p+eps)         ratio, only where X is stored (structural zeros stay zero)
//        kl += x_p log q_p ;  G[i,:] += q_p H[:,j_p]^T ;  W'[i,:] = W[i,:] (.) G[i,:]
//   pass 2, one warp per row (nmf.py:345-349):  N[:,j_p] += q_p W'[i,:]   (vector red.add into L2)
// The dictionary is kept TRANSPOSED (Ht: f x k) so that each non-zero gathers / scatters one
// contiguous k-vector (coalesced 128-bit accesses).
#include <stdlib.h>

#include "common.cuh"

namespace klnmf {

namespace {

constexpr int WARPS = 8;

template <typename T> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

template <typename T>
__device__ __forceinline__ void ld4(const T *p, T v[4]) {
  if (sizeof(T) == 4) {
    float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = (T)t.x; v[1] = (T)t.y; v[2] = (T)t.z; v[3] = (T)t.w;
  } else {
    double2 t0 = *reinterpret_cast<const double2 *>(p);
    double2 t1 = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = (T)t0.x; v[1] = (T)t0.y; v[2] = (T)t1.x; v[3] = (T)t1.y;
  }
}
template <typename T>
__device__ __forceinline__ void st4(T *p, const T v[4]) {
  if (sizeof(T) == 4) {
    *reinterpret_cast<float4 *>(p) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
  } else {
    *reinterpret_cast<double2 *>(p) = make_double2((double)v[0], (double)v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2((double)v[2], (double)v[3]);
  }
}
__device__ __forceinline__ void red4(float *p, const float v[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3])
               : "memory");
}
__device__ __forceinline__ void red4(double *p, const double v[4]) {
#pragma unroll
  for (int e = 0; e < 4; e++) atomicAdd(p + e, v[e]);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__device__ __forceinline__ void ldg4(const T *p, T v[4]) {   // read-only path (the dictionary never changes inside a pass)
  if (sizeof(T) == 4) {
    float4 t = __ldg(reinterpret_cast<const float4 *>(p));
    v[0] = (T)t.x; v[1] = (T)t.y; v[2] = (T)t.z; v[3] = (T)t.w;
  } else {
    double2 t0 = __ldg(reinterpret_cast<const double2 *>(p));
    double2 t1 = __ldg(reinterpret_cast<const double2 *>(p + 2));
    v[0] = (T)t0.x; v[1] = (T)t0.y; v[2] = (T)t1.x; v[3] = (T)t1.y;
  }
}

// One halving step of the batched warp reduction: every lane holds CNT partial sums (one per non-zero of the
// batch); lanes whose OFF bit is set keep the upper half, the others the lower half, and each adds what its
// partner held of the half it keeps.  log2(B) such steps followed by plain butterflies leave the complete
// sum of non-zero b on the lanes with lane >> (5 - log2 B) == b: B - 1 + (5 - log2 B) shuffles per batch
// instead of 5 per non-zero.
template <typename T, int CNT, int OFF>
__device__ __forceinline__ void halve(T v[], int lane) {
  const bool up = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < CNT / 2; i++) {
    const T send = up ? v[i] : v[i + CNT / 2];
    const T keep = up ? v[i + CNT / 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
  (void)lane;
}
template <typename T, int B>
__device__ __forceinline__ T batch_reduce(T v[], int lane) {
  static_assert(B == 8 || B == 4 || B == 2 || B == 1, "batch of 1, 2, 4 or 8 non-zeros");
  if (B >= 2) halve<T, B, 16>(v, lane);
  if (B >= 4) halve<T, B / 2, 8>(v, lane);
  if (B >= 8) halve<T, B / 4, 4>(v, lane);
  T r = v[0];
  if (B < 2) r += __shfl_xor_sync(0xffffffffu, r, 16);
  if (B < 4) r += __shfl_xor_sync(0xffffffffu, r, 8);
  if (B < 8) r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
// batch size: the gathered columns of a batch stay in registers (<= 64 32-bit words per lane)
template <typename T, int VPL>
struct Batch {
  static constexpr int WORDS = VPL * 4 * (int)(sizeof(T) / 4);
  static constexpr int B = WORDS <= 8 ? 8 : (WORDS <= 16 ? 4 : (WORDS <= 32 ? 2 : 1));
  static constexpr int SH = B == 8 ? 2 : (B == 4 ? 3 : (B == 2 ? 4 : 5));   // lanes per entry after the reduction = 1 << SH
};

// lane owns the k-elements { 128*c + 4*lane + e : c < VPL, e < 4 } (zero padded up to ld)
// MODE 0: full pass 1;  MODE 1: objective only;  MODE 2: W0 = X . H0^T (nmf.py:156);
// MODE 3: SDDMM only, qnz <- (W.H) at the non-zeros (nmf.py:52-70)
// A row is walked in batches of B stored entries: B dictionary columns are gathered (one coalesced k-vector
// each), their dot products with the row of W are reduced together, and the lanes {b << SH} finish entry b
// (ratio, objective term) in parallel before the SAME gathered columns are reused for G += q_b H[:,j_b].
template <typename T, int VPL, int MODE>
__global__ void __launch_bounds__(WARPS * 32)
sparse_rows_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                   const T *__restrict__ vals, const T *__restrict__ W, int64_t ldw, const T *__restrict__ Ht,
                   int64_t ldh, T *__restrict__ Wn, T *__restrict__ qnz, int64_t n, double *__restrict__ dred,
                   const int *stop, const T *__restrict__ g0) {
  // g0 (MODE 0 and 2, may be null): n x ldw, what G starts from -- the dense block's share Q_d.H_d^T of a hybrid stack
  // (api.cu: HybridSide), so that W' = W (.) (G_dense + G_sparse) comes out of this pass
  constexpr int B = Batch<T, VPL>::B, SH = Batch<T, VPL>::SH;
  if (stop != nullptr && *stop != 0) return;
  __shared__ double red[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t n_warps = (int64_t)gridDim.x * WARPS;
  const int mine = lane >> SH;                      // the entry of a batch this lane finishes
  const bool leader = (lane & ((1 << SH) - 1)) == 0;
  bool act[VPL];
#pragma unroll
  for (int c = 0; c < VPL; c++) act[c] = (128 * c + 4 * lane) < ldh && (128 * c + 4 * lane) < ldw;

  double kl = 0.0;
  double cs[VPL][4];
#pragma unroll
  for (int c = 0; c < VPL; c++)
#pragma unroll
    for (int e = 0; e < 4; e++) cs[c][e] = 0.0;

  for (int64_t i = warp_global; i < n; i += n_warps) {
    T w[VPL][4], g[VPL][4];
#pragma unroll
    for (int c = 0; c < VPL; c++) {
#pragma unroll
      for (int e = 0; e < 4; e++) { w[c][e] = (T)0; g[c][e] = (T)0; }
      if (MODE != 2 && act[c]) ld4(W + i * ldw + 128 * c + 4 * lane, w[c]);
      if ((MODE == 0 || MODE == 2) && g0 != nullptr && act[c]) ld4(g0 + i * ldw + 128 * c + 4 * lane, g[c]);
      if (MODE != 2) {
#pragma unroll
        for (int e = 0; e < 4; e++) cs[c][e] += (double)w[c][e];
      }
    }
    const int64_t p0 = indptr[i], p1 = indptr[i + 1];
    float klf = 0.f;
    for (int64_t p = p0; p < p1; p += B) {
      const int nb = (int)(p1 - p < B ? p1 - p : B);        // entries of this batch
      const bool have = mine < nb;
      const T xm = have ? __ldg(vals + p + mine) : (T)0;    // the value of "my" entry
      T h[B][VPL][4];
      T s[B];
#pragma unroll
      for (int b = 0; b < B; b++) {
        // entries past the end of the row re-read the first column of the batch; their coefficient is forced to 0
        const int32_t j = __ldg(indices + (b < nb ? p + b : p));
        s[b] = (T)0;
#pragma unroll
        for (int c = 0; c < VPL; c++) {
#pragma unroll
          for (int e = 0; e < 4; e++) h[b][c][e] = (T)0;
          if (act[c]) ldg4(Ht + (int64_t)j * ldh + 128 * c + 4 * lane, h[b][c]);
        }
      }
      T coef;                                               // what multiplies column b in the SpMM, on lanes of entry b
      if (MODE == 2) {
        coef = xm;
      } else {
#pragma unroll
        for (int b = 0; b < B; b++)
#pragma unroll
          for (int c = 0; c < VPL; c++)
#pragma unroll
            for (int e = 0; e < 4; e++) s[b] += w[c][e] * h[b][c][e];
        const T sm = batch_reduce<T, B>(s, lane);          // complete dot product of entry `mine`
        const T q = have ? (xm + (T)KL_EPS) / (sm + (T)KL_EPS) : (T)0;
        coef = q;
        if (have && leader) {
          if (sizeof(T) == 4) klf += (float)xm * logf((float)q);
          else kl += (double)xm * log((double)q);
          // (Writing q where the numerator pass reads it -- blocked-CSC order through a CSR -> blocked-CSC map, so that
          // pass streams it instead of gathering 4 bytes per 32-byte sector -- was measured: the scattered 4-byte stores
          // cost far more than the gathers they replace, rows pass 8.8 -> 20.0 ms at n = 500 000, numerator 13.3 -> 11.0;
          // profiles/r2_run13_sparse_q_order.log.)
          if (MODE == 0 && qnz != nullptr) qnz[p + mine] = q;
          if (MODE == 3) qnz[p + mine] = sm;
        }
      }
      if (MODE == 0 || MODE == 2) {
#pragma unroll
        for (int b = 0; b < B; b++) {
          const T cb = __shfl_sync(0xffffffffu, coef, b << SH);
#pragma unroll
          for (int c = 0; c < VPL; c++)
#pragma unroll
            for (int e = 0; e < 4; e++) g[c][e] += cb * h[b][c][e];
        }
      }
    }
    kl += (double)klf;                                       // <= a row of terms in FP32, FP64 across rows
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int c = 0; c < VPL; c++) {
        if (MODE == 0) {
#pragma unroll
          for (int e = 0; e < 4; e++) g[c][e] *= w[c][e];
        }
        if (act[c]) st4(Wn + i * ldw + 128 * c + 4 * lane, g[c]);
      }
    }
  }
  if (MODE == 2) return;
  // colsum(W) for the sparse objective's sum_k colsum(W)_k rowsum(H)_k term (nmf.py:304)
#pragma unroll
  for (int c = 0; c < VPL; c++)
    if (act[c]) {
#pragma unroll
      for (int e = 0; e < 4; e++) atomicAdd(&dred[2 + 128 * c + 4 * lane + e], cs[c][e]);
    }
  kl = warp_sum(kl);
  if (lane == 0) red[warp] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < WARPS; w++) t += red[w];
    atomicAdd(&dred[0], t);
  }
}


// ---------------------------------------------------------------------------------------------------------
// The rows pass for FP32 state whose padded width is exactly 128 * VPL (k = 128, 256, 512, 1024: cfg4 and the
// reference's k = 200 fall here after padding only when k itself is a multiple of 128; everything else takes the
// generic kernel above).  Same arithmetic, fewer instructions: the generic kernel issued 52 per stored entry at
// k = 256 and kept the schedulers 69 % busy (ncu, n = 10^6: `not_selected` + `selected` = 35 % of the stall samples) --
// it was bound by its issue rate, not by the 1 KB it gathers per entry.  Here
//   * no predicates and no zero-initialisation of the gather registers (32 CS2R per batch of 8 entries),
//   * the column indices of a batch are fetched by lanes 0..B-1 and handed out by shuffles, offsets inside the
//     dictionary are 32-bit (f * ld < 2^32): one IMAD.WIDE per gather instead of a 64-bit multiply-add chain,
//   * the dot products and the G update work on pairs (fma.rn.f32x2 -> FFMA2): 64 instead of 128 FMAs per batch.
// ---------------------------------------------------------------------------------------------------------
typedef unsigned long long f2_t;   // two packed FP32 values
__device__ __forceinline__ f2_t pk2(float a, float b) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f2_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2_t fma2(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2_t mul2(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int VPL, int MODE>   // MODE 0: full pass, 1: objective only
__global__ void __launch_bounds__(WARPS * 32, VPL <= 2 ? 2 : 1)
sparse_rows_full_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                        const float *__restrict__ vals, const float *__restrict__ W, const float *__restrict__ Ht,
                        uint32_t ld, float *__restrict__ Wn, float *__restrict__ qnz, int64_t n,
                        double *__restrict__ dred, const int *stop, const float *__restrict__ g0) {
  constexpr int B = Batch<float, VPL>::B, SH = Batch<float, VPL>::SH;
  if (stop != nullptr && *stop != 0) return;
  __shared__ double red[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t n_warps = (int64_t)gridDim.x * WARPS;
  const int mine = lane >> SH;
  const bool leader = (lane & ((1 << SH) - 1)) == 0;

  double kl = 0.0;
  double cs[VPL][4];
#pragma unroll
  for (int c = 0; c < VPL; c++)
#pragma unroll
    for (int e = 0; e < 4; e++) cs[c][e] = 0.0;

  // column indices of the batch starting at p, on lanes 0..B-1 (entries past the end of the row re-read the batch's
  // first column; their coefficient is forced to 0)
  auto batch_indices = [&](uint32_t p, uint32_t pend) -> int32_t {
    const int nb = (int)(pend - p < B ? pend - p : B);
    return lane < B ? __ldg(indices + (lane < nb ? p + lane : p)) : 0;
  };
  // What bounds the pass is the rate at which the L2 answers 1 KB gathers spread over the whole 51 MB dictionary
  // (cfg4): 512 GB in 33.5 ms = 15.3 TB/s, 0.94 of what klnmf_l2_read_bench gets out of a 48 MB buffer.  Measured and
  // dropped (profiles/r2_run26..29_*.log): FP16 gather copies (half the bytes: -5 %, at 100x the error), register
  // double-buffering of half batches (+34 %), L1 prefetches of the next batch (+19 %; +40 % with no-allocate loads),
  // half batches at three CTAs per SM (80 registers: +20 %).
  // (row and entry positions fit 32 bits: the host checks n, nnz < 2^32)
  uint32_t i = (uint32_t)warp_global;
  const uint32_t nn = (uint32_t)n, step = (uint32_t)n_warps;
  uint32_t p0 = 0, p1 = 0;
  int32_t jl = 0;
  if (warp_global < n) {
    p0 = (uint32_t)indptr[i]; p1 = (uint32_t)indptr[i + 1];
    if (p0 < p1) jl = batch_indices(p0, p1);
  }
  for (; warp_global < n && i < nn;) {
    const uint32_t in = i + step;
    uint32_t pn0 = 0, pn1 = 0;
    if (in < nn) { pn0 = (uint32_t)indptr[in]; pn1 = (uint32_t)indptr[in + 1]; }
    f2_t w[VPL][2], g[VPL][2];
    const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(W + (uint64_t)i * ld) + lane;
#pragma unroll
    for (int c = 0; c < VPL; c++) {
      const ulonglong2 t = wp[32 * c];
      w[c][0] = t.x; w[c][1] = t.y;
      float a0, a1, a2, a3;
      upk2(t.x, a0, a1); upk2(t.y, a2, a3);
      cs[c][0] += (double)a0; cs[c][1] += (double)a1; cs[c][2] += (double)a2; cs[c][3] += (double)a3;
      g[c][0] = pk2(0.f, 0.f); g[c][1] = pk2(0.f, 0.f);
      if (MODE == 0 && g0 != nullptr) {           // hybrid stack: G starts from the dense block's share
        const ulonglong2 t0 = (reinterpret_cast<const ulonglong2 *>(g0 + (uint64_t)i * ld) + lane)[32 * c];
        g[c][0] = t0.x; g[c][1] = t0.y;
      }
    }
    float klf = 0.f;
    int32_t jl_next = 0;
    if (p0 == p1 && pn0 < pn1) jl_next = batch_indices(pn0, pn1);      // an empty row hands over to the next one
    for (uint32_t p = p0; p < p1; p += B) {
      const int nb = (int)(p1 - p < B ? p1 - p : B);
      const bool have = mine < nb;
      const float xm = have ? __ldg(vals + p + mine) : 0.f;
      // the batch after this one: the next of this row, or the first of the warp's next row
      const bool more = p + B < p1;
      if (more) jl_next = batch_indices(p + B, p1);
      else if (pn0 < pn1) jl_next = batch_indices(pn0, pn1);
      f2_t h[B][VPL][2];
      float s[B];
#pragma unroll
      for (int b = 0; b < B; b++) {
        const uint32_t j = (uint32_t)__shfl_sync(0xffffffffu, jl, b);
        const ulonglong2 *col = reinterpret_cast<const ulonglong2 *>(Ht + (uint64_t)(j * ld)) + lane;
#pragma unroll
        for (int c = 0; c < VPL; c++) {
          const ulonglong2 t = __ldg(col + 32 * c);
          h[b][c][0] = t.x; h[b][c][1] = t.y;
        }
      }
#pragma unroll
      for (int b = 0; b < B; b++) {
        f2_t a2 = mul2(w[0][0], h[b][0][0]);
        a2 = fma2(w[0][1], h[b][0][1], a2);
#pragma unroll
        for (int c = 1; c < VPL; c++) { a2 = fma2(w[c][0], h[b][c][0], a2); a2 = fma2(w[c][1], h[b][c][1], a2); }
        float lo, hi;
        upk2(a2, lo, hi);
        s[b] = lo + hi;
      }
      const float sm = batch_reduce<float, B>(s, lane);
      const float q = have ? (xm + (float)KL_EPS) / (sm + (float)KL_EPS) : 0.f;
      if (have && leader) {
        klf += xm * logf(q);
        if (MODE == 0 && qnz != nullptr) qnz[p + mine] = q;
      }
      if (MODE == 0) {
#pragma unroll
        for (int b = 0; b < B; b++) {
          const float cb = __shfl_sync(0xffffffffu, q, b << SH);
          const f2_t cb2 = pk2(cb, cb);
#pragma unroll
          for (int c = 0; c < VPL; c++) { g[c][0] = fma2(cb2, h[b][c][0], g[c][0]); g[c][1] = fma2(cb2, h[b][c][1], g[c][1]); }
        }
      }
      jl = jl_next;
    }
    if (p0 == p1) jl = jl_next;
    kl += (double)klf;
    if (MODE == 0) {
      ulonglong2 *op = reinterpret_cast<ulonglong2 *>(Wn + (uint64_t)i * ld) + lane;
#pragma unroll
      for (int c = 0; c < VPL; c++) {
        ulonglong2 o;
        o.x = mul2(w[c][0], g[c][0]); o.y = mul2(w[c][1], g[c][1]);
        op[32 * c] = o;
      }
    }
    i = in; p0 = pn0; p1 = pn1;
  }
#pragma unroll
  for (int c = 0; c < VPL; c++)
#pragma unroll
    for (int e = 0; e < 4; e++) atomicAdd(&dred[2 + 128 * c + 4 * lane + e], cs[c][e]);
  kl = warp_sum(kl);
  if (lane == 0) red[warp] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < WARPS; w++) t += red[w];
    atomicAdd(&dred[0], t);
  }
}

template <typename T, int VPL>
__global__ void __launch_bounds__(WARPS * 32)
sparse_scatter_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                      const T *__restrict__ qnz, const T *__restrict__ Wn, int64_t ldw, T *__restrict__ Nt,
                      int64_t ldh, int64_t n, const int *stop) {
  if (*stop != 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t n_warps = (int64_t)gridDim.x * WARPS;
  bool act[VPL];
#pragma unroll
  for (int c = 0; c < VPL; c++) act[c] = (128 * c + 4 * lane) < ldh && (128 * c + 4 * lane) < ldw;
  for (int64_t i = warp_global; i < n; i += n_warps) {
    T w[VPL][4];
#pragma unroll
    for (int c = 0; c < VPL; c++) {
#pragma unroll
      for (int e = 0; e < 4; e++) w[c][e] = (T)0;
      if (act[c]) ld4(Wn + i * ldw + 128 * c + 4 * lane, w[c]);
    }
    const int64_t p0 = indptr[i], p1 = indptr[i + 1];
    for (int64_t p = p0; p < p1; p++) {
      const int32_t j = indices[p];
      const T q = qnz[p];
#pragma unroll
      for (int c = 0; c < VPL; c++)
        if (act[c]) {
          T v[4];
#pragma unroll
          for (int e = 0; e < 4; e++) v[e] = q * w[c][e];
          red4(Nt + (int64_t)j * ldh + 128 * c + 4 * lane, v);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Dictionary numerator without per-entry atomics: a blocked-CSC copy of the sparsity pattern.
// Rows are cut into blocks of R samples whose W' rows (R x k) stay L2-resident; inside a block the stored
// entries are regrouped by column.  One warp then owns a (block, column) unit: it GATHERS the W' rows of the
// unit's entries (reads cross the L2 2.3x faster than red.add does), accumulates q_ij W'[i,:] in
// registers and issues ONE vector red.add per unit instead of one per entry.  The copy costs 8 bytes per stored
// entry and is built once per data set on the device (count -> scan -> fill).
// ---------------------------------------------------------------------------------------------------------
struct BlockedCsc {
  int64_t R = 0;
  int64_t n_blocks = 0;
  int32_t *colptr = nullptr;   // n_blocks x (f + 1): offsets of the columns inside the block's entry range
  int32_t *rowidx = nullptr;   // nnz: row inside the block
  int32_t *src = nullptr;      // nnz: CSR position inside the block's entry range (where q_ij lives)
  unsigned long long *ticket = nullptr;   // next unit to hand out (zeroed before every launch)
};

__global__ void bcsc_count_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, int64_t n,
                                  int64_t f, int64_t R, int32_t *__restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp_global; i < n; i += n_warps) {
    int32_t *c = cnt + (i / R) * (f + 1);
    for (int64_t p = indptr[i] + lane; p < indptr[i + 1]; p += 32) atomicAdd(c + indices[p], 1);
  }
}
// exclusive scan of every block's f + 1 counters, one CTA per block
__global__ void __launch_bounds__(1024) bcsc_scan_kernel(int32_t *__restrict__ cnt, int64_t f) {
  __shared__ int32_t part[1024];
  int32_t *c = cnt + (int64_t)blockIdx.x * (f + 1);
  const int64_t len = f + 1, per = (len + 1023) / 1024;
  const int64_t lo = threadIdx.x * per, hi = lo + per < len ? lo + per : len;
  int32_t sum = 0;
  for (int64_t t = lo; t < hi; t++) sum += c[t];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int32_t v = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int32_t run = part[threadIdx.x] - sum;
  for (int64_t t = lo; t < hi; t++) { const int32_t v = c[t]; c[t] = run; run += v; }
}
__global__ void bcsc_fill_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices, int64_t n,
                                 int64_t f, int64_t R, int32_t *__restrict__ cursor, int32_t *__restrict__ rowidx,
                                 int32_t *__restrict__ src) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp_global; i < n; i += n_warps) {
    const int64_t b = i / R, base = indptr[b * R];
    int32_t *c = cursor + b * (f + 1);
    for (int64_t p = indptr[i] + lane; p < indptr[i + 1]; p += 32) {
      const int32_t pos = atomicAdd(c + indices[p], 1);
      rowidx[base + pos] = (int32_t)(i - b * R);
      src[base + pos] = (int32_t)(p - base);
    }
  }
}

template <typename T, int VPL>
__global__ void __launch_bounds__(WARPS * 32)
sparse_numerator_bcsc_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ colptr,
                             const int32_t *__restrict__ rowidx, const int32_t *__restrict__ src,
                             const T *__restrict__ qnz, const T *__restrict__ Wn, int64_t ldw, T *__restrict__ Nt,
                             int64_t ldh, int64_t n, int64_t f, int64_t R, int64_t n_blocks, const int *stop,
                             unsigned long long *__restrict__ ticket) {
  if (*stop != 0) return;
  const int lane = threadIdx.x & 31;
  bool act[VPL];
#pragma unroll
  for (int c = 0; c < VPL; c++) act[c] = (128 * c + 4 * lane) < ldh && (128 * c + 4 * lane) < ldw;
  const int64_t units = n_blocks * f;
  // Units are handed out IN ORDER by a ticket counter, one per warp and grab: every resident warp then works within a
  // few thousand units of the frontier, i.e. inside ONE row block, whose W' rows stay in the L2.  (A static grid-stride
  // assignment lets the warps of faster SMs run ahead by whole blocks over a long launch: at n = 10^6 the pass read
  // 88.7 GB from DRAM for 4 GB of compulsory traffic, L2 hit rate 54 %, and its time per row grew with n -- 19.5 ns at
  // n = 131 072, 31.9 ns at 2 * 10^6; profiles/r2_run24_sparse_scaling.log.)
  // Software pipeline over the dependent loads (ticket -> column range -> row index / ratio position -> ratio ->
  // coefficient rows): the ticket and the column range of the NEXT unit and the (row, ratio) pairs of the NEXT batch of
  // eight entries are fetched before the eight row gathers of the current batch are issued, so the L2 round trips of
  // the chain overlap the gathers instead of preceding them (ncu: long_scoreboard was 75 % of all stall samples).
  auto grab = [&]() -> int64_t {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(ticket, 1ull);
    return (int64_t)__shfl_sync(0xffffffffu, t, 0);
  };
  int64_t u = grab();
  int32_t c0 = 0, c1 = 0;
  if (u < units) {
    const int32_t *cp = colptr + (u / f) * (f + 1) + (u % f);
    c0 = __ldg(cp); c1 = __ldg(cp + 1);
  }
  int64_t un = u;
  for (; u < units; u = un) {   // block-major: the whole grid sweeps one row block at a time
    const int64_t b = u / f, j = u - b * f;
    const int32_t t0 = c0, t1 = c1;
    {
      un = grab();
      if (un < units) {
        const int32_t *cp = colptr + (un / f) * (f + 1) + (un % f);
        c0 = __ldg(cp); c1 = __ldg(cp + 1);
      }
    }
    if (t0 == t1) continue;
    const int64_t base = indptr[b * R];
    const T *Wb = Wn + b * R * ldw;
    T acc[VPL][4];
#pragma unroll
    for (int c = 0; c < VPL; c++)
#pragma unroll
      for (int e = 0; e < 4; e++) acc[c][e] = (T)0;
    int32_t ri = 0;
    T q = (T)0;
    if (lane < 8 && t0 + lane < t1) {
      ri = __ldg(rowidx + base + t0 + lane);
      q = qnz[base + __ldg(src + base + t0 + lane)];
    }
    for (int32_t t = t0; t < t1; t += 8) {
      int32_t ri_n = 0;
      T q_n = (T)0;
      if (lane < 8 && t + 8 + lane < t1) {
        ri_n = __ldg(rowidx + base + t + 8 + lane);
        q_n = qnz[base + __ldg(src + base + t + 8 + lane)];
      }
#pragma unroll
      for (int e8 = 0; e8 < 8; e8++) {
        const int32_t i = __shfl_sync(0xffffffffu, ri, e8);    // lanes past the end hold row 0 with q = 0
        const T qq = __shfl_sync(0xffffffffu, q, e8);
#pragma unroll
        for (int c = 0; c < VPL; c++)
          if (act[c]) {
            T w[4];
            ldg4(Wb + (int64_t)i * ldw + 128 * c + 4 * lane, w);
#pragma unroll
            for (int e = 0; e < 4; e++) acc[c][e] += qq * w[e];
          }
      }
      ri = ri_n;
      q = q_n;
    }
#pragma unroll
    for (int c = 0; c < VPL; c++)
      if (act[c]) red4(Nt + j * ldh + 128 * c + 4 * lane, acc[c]);
  }
}

template <typename T>
__global__ void fill_csr_kernel(int64_t *__restrict__ indptr, int32_t *__restrict__ indices, T *__restrict__ vals,
                                int64_t n, int64_t f, int64_t m, uint64_t seed) {
  const int64_t total = n * m;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / m, t = i - r * m;
    uint64_t z = seed ^ ((uint64_t)i * 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    // stratified: the t-th non-zero of a row falls in the t-th of m equal column strata -> sorted, distinct
    int64_t lo = t * f / m, hi = (t + 1) * f / m;
    indices[i] = (int32_t)(lo + (int64_t)(z % (uint64_t)(hi - lo)));
    vals[i] = (T)(((double)(z >> 11) + 1.0) * (1.0 / 9007199254740992.0));
    if (t == 0) indptr[r] = r * m;
    if (i == total - 1) indptr[n] = total;
  }
}

int bcsc_build(klnmf_ctx *ctx);
// the blocked-CSC numerator (default) indexes a block's entries with int32 offsets: fine while nnz < 2^31 (cfg4: 5e8);
// KLNMF_SPARSE_ATOMICS=1 selects the first version (one red.add per stored entry, CSR order)
inline bool use_bcsc(const klnmf_ctx *ctx) {
  static const bool atomics = getenv("KLNMF_SPARSE_ATOMICS") && atoi(getenv("KLNMF_SPARSE_ATOMICS")) == 1;
  return !atomics && ctx->nnz > 0 && ctx->nnz < (((int64_t)1 << 31) - 1);
}

// q_order (mode 0 only): 0 = the ratio is not kept (transform: 4 bytes per stored entry less to write), 1 / 2 = kept in
// CSR order, for the _Q hook / for the numerator pass
template <typename T, int VPL>
int run_rows(klnmf_ctx *ctx, int mode, const T *W, T *Wn, int q_order, const T *g0, int64_t r0, int64_t n) {
  // the rows [r0, r0 + n) of the data (hybrid stacks walk the samples in panels): stored-entry positions are absolute,
  // so a row range is the same kernels on offset row pointers
  const int64_t *indptr = ctx->indptr + r0;
  if (W) W += r0 * ctx->ldw;
  if (Wn) Wn += r0 * ctx->ldw;
  if (g0) g0 += r0 * ctx->ldw;
  const int grid = (int)(ceil_div(n, WARPS) < (int64_t)ctx->sm_count * 8 ? ceil_div(n, WARPS) : (int64_t)ctx->sm_count * 8);
  const int *stop = ctx->flags + FL_STOP;
  const T *Ht = (const T *)ctx->H[ctx->hcur];
  static const bool generic_only = getenv("KLNMF_SPARSE_GENERIC") && atoi(getenv("KLNMF_SPARSE_GENERIC")) == 1;
  if (sizeof(T) == 4 && (mode == 0 || mode == 1) && !generic_only && ctx->ldw == 128 * VPL && ctx->ldh == ctx->ldw &&
      ctx->f * ctx->ldh < ((int64_t)1 << 32) && ctx->nnz < ((int64_t)1 << 32) - 64 &&
      n + (int64_t)grid * WARPS < ((int64_t)1 << 32)) {
    if (mode == 0)
      sparse_rows_full_kernel<VPL, 0><<<grid, WARPS * 32, 0, ctx->stream>>>(
          indptr, ctx->indices, (const float *)ctx->vals, (const float *)W, (const float *)Ht, (uint32_t)ctx->ldh,
          (float *)Wn, q_order == 0 ? nullptr : (float *)ctx->qnz, n, ctx->dred, stop, (const float *)g0);
    else
      sparse_rows_full_kernel<VPL, 1><<<grid, WARPS * 32, 0, ctx->stream>>>(
          indptr, ctx->indices, (const float *)ctx->vals, (const float *)W, (const float *)Ht, (uint32_t)ctx->ldh,
          nullptr, nullptr, n, ctx->dred, nullptr, nullptr);
    ctx->n_launch++;
    KL_CUDA(cudaGetLastError());
    return KLNMF_OK;
  }
  if (mode == 0)
    sparse_rows_kernel<T, VPL, 0><<<grid, WARPS * 32, 0, ctx->stream>>>(indptr, ctx->indices, (const T *)ctx->vals, W,
                                                                        ctx->ldw, Ht, ctx->ldh, Wn,
                                                                        q_order == 0 ? nullptr : (T *)ctx->qnz, n,
                                                                        ctx->dred, stop, g0);
  else if (mode == 1)
    sparse_rows_kernel<T, VPL, 1><<<grid, WARPS * 32, 0, ctx->stream>>>(indptr, ctx->indices, (const T *)ctx->vals, W,
                                                                        ctx->ldw, Ht, ctx->ldh, nullptr, nullptr, n,
                                                                        ctx->dred, nullptr, nullptr);
  else if (mode == 3)
    sparse_rows_kernel<T, VPL, 3><<<grid, WARPS * 32, 0, ctx->stream>>>(indptr, ctx->indices, (const T *)ctx->vals, W,
                                                                        ctx->ldw, Ht, ctx->ldh, nullptr, (T *)ctx->qnz, n,
                                                                        ctx->dred, nullptr, nullptr);
  else
    sparse_rows_kernel<T, VPL, 2><<<grid, WARPS * 32, 0, ctx->stream>>>(indptr, ctx->indices, (const T *)ctx->vals,
                                                                        nullptr, ctx->ldw, Ht, ctx->ldh, Wn, nullptr, n,
                                                                        nullptr, nullptr, g0);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

int bcsc_build(klnmf_ctx *ctx) {
  BlockedCsc *bc = new BlockedCsc();
  const int64_t row_bytes = ctx->ldw * (int64_t)ctx->es;
  // W' block of about 24 MB: it stays L2-resident next to the k x f numerator the units add into (KLNMF_SPARSE_BLOCK_MB)
  static const int64_t block_mb = getenv("KLNMF_SPARSE_BLOCK_MB") ? atoi(getenv("KLNMF_SPARSE_BLOCK_MB")) : 24;
  int64_t R = ((block_mb > 0 ? block_mb : 24) << 20) / row_bytes / 1024 * 1024;
  if (R < 1024) R = 1024;
  if (R > ctx->n) R = ctx->n;
  bc->R = R;
  bc->n_blocks = ceil_div(ctx->n, R);
  const int64_t cells = bc->n_blocks * (ctx->f + 1);
  int32_t *cursor = nullptr;
  auto fail = [&](const char *what) {
    set_error("blocked-CSC build: %s", what);
    if (bc->colptr) cudaFree(bc->colptr);
    if (bc->rowidx) cudaFree(bc->rowidx);
    if (bc->src) cudaFree(bc->src);
    if (bc->ticket) cudaFree(bc->ticket);
    if (cursor) cudaFree(cursor);
    delete bc;
    return KLNMF_ENOMEM;
  };
  const int64_t nz = ctx->nnz > 0 ? ctx->nnz : 1;
  if (cudaMalloc((void **)&bc->colptr, cells * 4) != cudaSuccess) return fail("cudaMalloc colptr");
  if (cudaMalloc((void **)&cursor, cells * 4) != cudaSuccess) return fail("cudaMalloc cursor");
  if (cudaMalloc((void **)&bc->rowidx, nz * 4) != cudaSuccess) return fail("cudaMalloc rowidx");
  if (cudaMalloc((void **)&bc->src, nz * 4) != cudaSuccess) return fail("cudaMalloc src");
  if (cudaMalloc((void **)&bc->ticket, 8) != cudaSuccess) return fail("cudaMalloc ticket");
  cudaMemsetAsync(bc->colptr, 0, cells * 4, ctx->stream);
  const int grid = ctx->sm_count * 8;
  bcsc_count_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->indptr, ctx->indices, ctx->n, ctx->f, R, bc->colptr);
  bcsc_scan_kernel<<<(unsigned)bc->n_blocks, 1024, 0, ctx->stream>>>(bc->colptr, ctx->f);
  cudaMemcpyAsync(cursor, bc->colptr, cells * 4, cudaMemcpyDeviceToDevice, ctx->stream);
  bcsc_fill_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->indptr, ctx->indices, ctx->n, ctx->f, R, cursor, bc->rowidx, bc->src);
  ctx->n_launch += 3;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return fail("kernels failed");
  cudaFree(cursor);
  ctx->bcsc = bc;
  return KLNMF_OK;
}

template <typename T, int VPL>
int run_scatter(klnmf_ctx *ctx, const T *Wn) {
  const int grid = (int)(ceil_div(ctx->n, WARPS) < (int64_t)ctx->sm_count * 8 ? ceil_div(ctx->n, WARPS)
                                                                               : (int64_t)ctx->sm_count * 8);
  if (use_bcsc(ctx)) {
    if (!ctx->bcsc) KL_TRY(bcsc_build(ctx));
    const BlockedCsc *bc = (const BlockedCsc *)ctx->bcsc;
    // exactly the CTAs that are resident at once: with tickets a second wave would only find the counter exhausted
    static int per_sm[64] = {};
    if (!per_sm[ctx->device & 63]) {
      int b = 0;
      KL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sparse_numerator_bcsc_kernel<T, VPL>, WARPS * 32, 0));
      per_sm[ctx->device & 63] = b > 0 ? b : 1;
    }
    KL_CUDA(cudaMemsetAsync(bc->ticket, 0, 8, ctx->stream));
    sparse_numerator_bcsc_kernel<T, VPL><<<ctx->sm_count * per_sm[ctx->device & 63], WARPS * 32, 0, ctx->stream>>>(
        ctx->indptr, bc->colptr, bc->rowidx, bc->src, (const T *)ctx->qnz, Wn, ctx->ldw, (T *)ctx->num, ctx->ldh, ctx->n,
        ctx->f, bc->R, bc->n_blocks, ctx->flags + FL_STOP, bc->ticket);
    ctx->n_launch++;
    KL_CUDA(cudaGetLastError());
    return KLNMF_OK;
  }
  sparse_scatter_kernel<T, VPL><<<grid, WARPS * 32, 0, ctx->stream>>>(ctx->indptr, ctx->indices, (const T *)ctx->qnz, Wn,
                                                                      ctx->ldw, (T *)ctx->num, ctx->ldh, ctx->n,
                                                                      ctx->flags + FL_STOP);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

template <typename T>
int dispatch_rows(klnmf_ctx *ctx, int mode, const T *W, T *Wn, int q_order, const T *g0, int64_t r0, int64_t rows) {
  const int64_t kp = ctx->ldw;
  if (kp <= 128) return run_rows<T, 1>(ctx, mode, W, Wn, q_order, g0, r0, rows);
  if (kp <= 256) return run_rows<T, 2>(ctx, mode, W, Wn, q_order, g0, r0, rows);
  if (kp <= 512) return run_rows<T, 4>(ctx, mode, W, Wn, q_order, g0, r0, rows);
  if (kp <= 1024) return run_rows<T, 8>(ctx, mode, W, Wn, q_order, g0, r0, rows);
  set_error("sparse path supports n_components <= 1024 (got %lld)", (long long)ctx->k);
  return KLNMF_EINVAL;
}
template <typename T>
int dispatch_scatter(klnmf_ctx *ctx, const T *Wn) {
  const int64_t kp = ctx->ldw;
  if (kp <= 128) return run_scatter<T, 1>(ctx, Wn);
  if (kp <= 256) return run_scatter<T, 2>(ctx, Wn);
  if (kp <= 512) return run_scatter<T, 4>(ctx, Wn);
  if (kp <= 1024) return run_scatter<T, 8>(ctx, Wn);
  set_error("sparse path supports n_components <= 1024 (got %lld)", (long long)ctx->k);
  return KLNMF_EINVAL;
}

}  // namespace

// Pass 1 of the sparse iteration (SDDMM -> ratio -> objective -> SpMM -> W update), or with
// only_error just the objective terms.
int sparse_rows(klnmf_ctx *ctx, int mode, int q_order, const void *g0, int64_t r0, int64_t rows) {
  const int cur = ctx->cur;
  if (rows < 0) { r0 = 0; rows = ctx->n; }
  if (rows == 0) return KLNMF_OK;
  return ctx->es == 8 ? dispatch_rows<double>(ctx, mode, (const double *)ctx->W[cur], (double *)ctx->W[cur ^ 1], q_order,
                                              (const double *)g0, r0, rows)
                      : dispatch_rows<float>(ctx, mode, (const float *)ctx->W[cur], (float *)ctx->W[cur ^ 1], q_order,
                                             (const float *)g0, r0, rows);
}

// Pass 2: dictionary numerator N^T[j,:] += q W'[i,:] over the stored non-zeros.
int sparse_scatter(klnmf_ctx *ctx, bool use_current_w) {
  const int w = use_current_w ? ctx->cur : ctx->cur ^ 1;
  return ctx->es == 8 ? dispatch_scatter<double>(ctx, (const double *)ctx->W[w])
                      : dispatch_scatter<float>(ctx, (const float *)ctx->W[w]);
}

int sparse_init_w(klnmf_ctx *ctx, const void *g0, int64_t r0, int64_t rows) {
  if (rows < 0) { r0 = 0; rows = ctx->n; }
  if (rows == 0) return KLNMF_OK;
  return ctx->es == 8 ? dispatch_rows<double>(ctx, 2, nullptr, (double *)ctx->W[ctx->cur], 0, (const double *)g0, r0, rows)
                      : dispatch_rows<float>(ctx, 2, nullptr, (float *)ctx->W[ctx->cur], 0, (const float *)g0, r0, rows);
}

void sparse_release_pattern(klnmf_ctx *ctx) {
  BlockedCsc *bc = (BlockedCsc *)ctx->bcsc;
  if (!bc) return;
  if (bc->colptr) cudaFree(bc->colptr);
  if (bc->rowidx) cudaFree(bc->rowidx);
  if (bc->src) cudaFree(bc->src);
  if (bc->ticket) cudaFree(bc->ticket);
  delete bc;
  ctx->bcsc = nullptr;
}

int sparse_fill_synthetic(klnmf_ctx *ctx, int64_t m, uint64_t seed) {
  sparse_release_pattern(ctx);
  const int64_t total = ctx->n * m;
  int64_t g = ceil_div(total, 256);
  if (g > (int64_t)ctx->sm_count * 32) g = (int64_t)ctx->sm_count * 32;
  if (ctx->es == 8)
    fill_csr_kernel<double><<<(unsigned)g, 256, 0, ctx->stream>>>(ctx->indptr, ctx->indices, (double *)ctx->vals, ctx->n,
                                                                  ctx->f, m, seed);
  else
    fill_csr_kernel<float><<<(unsigned)g, 256, 0, ctx->stream>>>(ctx->indptr, ctx->indices, (float *)ctx->vals, ctx->n,
                                                                 ctx->f, m, seed);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

}  // namespace klnmf
