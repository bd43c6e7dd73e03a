// Generic strided contraction with fused KL-NMF epilogues.
//   T = double : FP64 mode, tensor cores through DMMA (mma.sync.m8n8k4.f64)
//   T = float  : FP32 FMA bring-up engine (KLNMF_DEBUG_ENGINE=simt), never the default
// The three contractions of one reference iteration (nmf.py:325-351) all map onto it:
//   ratio      Q  = (X+eps)/(W.H+eps)          A=W  (K-contig)   B=H  (N-contig)   EPI_RATIO
//   coefficient W' = W (.) (Q.H^T)             A=Q  (K-contig)   B=H^T(K-contig)   EPI_MULW
//   numerator  N += W'^T.Q                     A=W'^T(M-contig)  B=Q  (N-contig)   EPI_ACC
#include "common.cuh"

namespace klnmf {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 8, NT = 256;

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <typename T>
struct EpiArgs {
  T *out; int64_t ldo;
  const T *aux; int64_t ldaux;
  int64_t M, N;
  int only_kl;
};

template <typename T, int EPI>
__device__ __forceinline__ void epi_elem(const EpiArgs<T> &e, int64_t row, int64_t col, T c, double &kl) {
  if (row >= e.M || col >= e.N) return;
  if (EPI == EPI_STORE) {
    e.out[row * e.ldo + col] = c;
  } else if (EPI == EPI_RATIO) {
    T x = e.aux[row * e.ldaux + col];
    T q = (x + (T)KL_EPS) / (c + (T)KL_EPS);
    if (!e.only_kl) e.out[row * e.ldo + col] = q;
    kl += (double)x * log((double)q) - (double)x + (double)c;
  } else if (EPI == EPI_MULW) {
    e.out[row * e.ldo + col] = e.aux[row * e.ldaux + col] * c;
  } else {
    atomicAdd(&e.out[row * e.ldo + col], c);
  }
}

template <typename T, int EPI>
__global__ void __launch_bounds__(NT) generic_gemm_kernel(GemmDesc d) {
  if (d.stop != nullptr && *d.stop != 0) return;
  __shared__ T As[BK][BM + PAD];
  __shared__ T Bs[BK][BN + PAD];
  __shared__ double red[NT / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const T *A = (const T *)d.A;
  const T *B = (const T *)d.B;

  int64_t k_begin = 0, k_end = d.K;
  if (EPI == EPI_ACC && d.splitk > 1) {
    int64_t per = ((d.K + d.splitk - 1) / d.splitk + BK - 1) / BK * BK;
    k_begin = (int64_t)blockIdx.z * per;
    k_end = k_begin + per < d.K ? k_begin + per : d.K;
    if (k_begin >= k_end) return;
  }

  constexpr bool kDmma = sizeof(T) == 8;
  // DMMA: warp grid 2 (M) x 4 (N), warp tile 32 x 16 = 4 x 2 m8n8 tiles
  // FMA : thread grid 16 x 16, thread tile 4 x 4
  const int wm = warp >> 2, wn = warp & 3;
  const int tx = tid & 15, ty = tid >> 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = (T)0;

  const bool a_kcontig = (d.a_sk == 1);
  const bool b_ncontig = (d.b_sn == 1);

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int it = 0; it < BM * BK / NT; it++) {
      int idx = tid + it * NT;
      int kk, m;
      if (a_kcontig) { kk = idx % BK; m = idx / BK; } else { m = idx % BM; kk = idx / BM; }
      int64_t gm = m0 + m, gk = k0 + kk;
      As[kk][m] = (gm < d.M && gk < k_end) ? A[gm * d.a_sm + gk * d.a_sk] : (T)0;
    }
#pragma unroll
    for (int it = 0; it < BN * BK / NT; it++) {
      int idx = tid + it * NT;
      int kk, n;
      if (b_ncontig) { n = idx % BN; kk = idx / BN; } else { kk = idx % BK; n = idx / BK; }
      int64_t gn = n0 + n, gk = k0 + kk;
      Bs[kk][n] = (gn < d.N && gk < k_end) ? B[gk * d.b_sk + gn * d.b_sn] : (T)0;
    }
    __syncthreads();
    if (kDmma) {
#pragma unroll
      for (int ks = 0; ks < BK; ks += 4) {
        double af[4], bf[2];
#pragma unroll
        for (int i = 0; i < 4; i++) af[i] = (double)As[ks + (lane & 3)][wm * 32 + i * 8 + (lane >> 2)];
#pragma unroll
        for (int j = 0; j < 2; j++) bf[j] = (double)Bs[ks + (lane & 3)][wn * 16 + j * 8 + (lane >> 2)];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 2; j++) {
            double c0 = (double)acc[i][2 * j], c1 = (double)acc[i][2 * j + 1];
            dmma_m8n8k4(c0, c1, af[i], bf[j]);
            acc[i][2 * j] = (T)c0;
            acc[i][2 * j + 1] = (T)c1;
          }
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < BK; ks++) {
        T a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = As[ks][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = Bs[ks][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
      }
    }
    __syncthreads();
  }

  EpiArgs<T> e{(T *)d.out, d.ldo, (const T *)d.aux, d.ldaux, d.M, d.N, d.only_kl};
  double kl = 0.0;
  if (kDmma) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        int64_t row = m0 + wm * 32 + i * 8 + (lane >> 2);
        int64_t col = n0 + wn * 16 + j * 8 + 2 * (lane & 3);
        epi_elem<T, EPI>(e, row, col, acc[i][2 * j], kl);
        epi_elem<T, EPI>(e, row, col + 1, acc[i][2 * j + 1], kl);
      }
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) epi_elem<T, EPI>(e, m0 + ty * 4 + i, n0 + tx * 4 + j, acc[i][j], kl);
  }
  if (EPI == EPI_RATIO && d.kl != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
    if (lane == 0) red[warp] = kl;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < NT / 32; w++) s += red[w];
      atomicAdd(d.kl, s);
    }
  }
}

template <typename T>
int launch_t(klnmf_ctx *ctx, int epi, const GemmDesc &d) {
  dim3 grid((unsigned)ceil_div(d.N, BN), (unsigned)ceil_div(d.M, BM), 1);
  KL_CHECK(grid.y <= 65535u * 64u, KLNMF_EINVAL, "generic_gemm: M too large");
  GemmDesc dd = d;
  if (epi == EPI_ACC) {
    int64_t tiles = (int64_t)grid.x * grid.y;
    int64_t want = ceil_div((int64_t)ctx->sm_count * 8, tiles);
    int64_t maxs = ceil_div(d.K, 512);
    int64_t s = want < maxs ? want : maxs;
    if (s < 1) s = 1;
    if (s > 65535) s = 65535;
    dd.splitk = (int)s;
    grid.z = (unsigned)s;
  }
  // grid.y is limited to 65535: fold large M into several launches
  const int64_t max_rows = (int64_t)65535 * BM;
  for (int64_t r0 = 0; r0 < d.M; r0 += max_rows) {
    GemmDesc c = dd;
    int64_t rows = d.M - r0 < max_rows ? d.M - r0 : max_rows;
    c.M = rows;
    c.A = (const char *)d.A + r0 * d.a_sm * (int64_t)sizeof(T);
    c.out = (char *)d.out + r0 * d.ldo * (int64_t)sizeof(T);
    if (d.aux) c.aux = (const char *)d.aux + r0 * d.ldaux * (int64_t)sizeof(T);
    grid.y = (unsigned)ceil_div(rows, BM);
    switch (epi) {
      case EPI_STORE: generic_gemm_kernel<T, EPI_STORE><<<grid, NT, 0, ctx->stream>>>(c); break;
      case EPI_RATIO: generic_gemm_kernel<T, EPI_RATIO><<<grid, NT, 0, ctx->stream>>>(c); break;
      case EPI_MULW: generic_gemm_kernel<T, EPI_MULW><<<grid, NT, 0, ctx->stream>>>(c); break;
      default: generic_gemm_kernel<T, EPI_ACC><<<grid, NT, 0, ctx->stream>>>(c); break;
    }
    ctx->n_launch++;
  }
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

}  // namespace

int generic_gemm(klnmf_ctx *ctx, int es, int epi, const GemmDesc &d) {
  KL_CHECK(d.qshift == 0.f && d.colbias == nullptr, KLNMF_EINVAL, "generic_gemm: the centered ratio is a tcgen05-engine feature");
  if (d.M <= 0 || d.N <= 0) return KLNMF_OK;
  return es == 8 ? launch_t<double>(ctx, epi, d) : launch_t<float>(ctx, epi, d);
}

}  // namespace klnmf
