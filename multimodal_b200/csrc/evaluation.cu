// Pairwise distances + nearest-example search on the device: the step right after the hot path
// (SURVEY 8f-3).  Replaces evaluation.all_distances / classify_NN (evaluation.py:103-116), which broadcast
// reco[:, None, :] against ex[None, :, :] (an n_test x n_ex x d temporary) before reducing it with one of
// metrics.kl_div / rev_kl_div / sym_kl_div / frobenius / cosine_diff (metrics.py:58-86).
//
// One CTA owns 16 test rows and walks the examples in chunks of 16 and the features in chunks of 32 through
// shared memory; thread (tx, ty) accumulates the FP64 distance of test row ty to example tx.  The distance
// matrix is written only when asked for; the running (min, first argmin) of np.argmin (evaluation.py:74)
// is kept per thread and reduced across the 16 example lanes at the end, so classification never
// materialises the n_test x n_ex matrix.  HBM-bound in principle (A and B read once per tile pair), FP64
// log / divide bound in practice for the KL measures: n_test x n_ex x d transcendental terms.
#include "common.cuh"

namespace klnmf {
namespace {

constexpr int NT = 16;    // test rows / examples per tile
constexpr int ND = 32;    // features per smem chunk

template <int MEASURE>
__device__ __forceinline__ void accumulate(double a, double b, double &s0, double &s1, double &s2) {
  if (MEASURE == KLNMF_MEASURE_KL) {            // generalized_KL(a, b) (metrics.py:18-20)
    s0 += a * log((a + KL_EPS) / (b + KL_EPS)) - a + b;
  } else if (MEASURE == KLNMF_MEASURE_REV_KL) { // kl_div(b, a)
    s0 += b * log((b + KL_EPS) / (a + KL_EPS)) - b + a;
  } else if (MEASURE == KLNMF_MEASURE_SYM_KL) { // .5 * (kl + rev_kl): the two sums are formed separately, as the reference does
    s0 += a * log((a + KL_EPS) / (b + KL_EPS)) - a + b;
    s1 += b * log((b + KL_EPS) / (a + KL_EPS)) - b + a;
  } else if (MEASURE == KLNMF_MEASURE_FROBENIUS) {
    const double d = a - b;
    s0 += d * d;
  } else {                                      // cosine_diff
    s0 += a * b; s1 += a * a; s2 += b * b;
  }
}
template <int MEASURE>
__device__ __forceinline__ double finish(double s0, double s1, double s2) {
  if (MEASURE == KLNMF_MEASURE_SYM_KL) return .5 * (s0 + s1);
  if (MEASURE == KLNMF_MEASURE_FROBENIUS) return sqrt(s0);
  if (MEASURE == KLNMF_MEASURE_COSINE_DIFF) return -(s0 / (sqrt(s1 * s2) + (s0 == 0.0 ? 1.0 : 0.0)));
  return s0;
}

template <int MEASURE>
__global__ void __launch_bounds__(NT * NT)
pairwise_kernel(const double *__restrict__ A, int64_t lda, const double *__restrict__ B, int64_t ldb, int64_t n_test,
                int64_t n_ex, int64_t d, double *__restrict__ D, int64_t ldd, int32_t *__restrict__ argmin,
                double *__restrict__ minval) {
  __shared__ double sa[NT][ND + 1], sb[NT][ND + 1];
  __shared__ double rmin[NT][NT];
  __shared__ int32_t ridx[NT][NT];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * NT + tx;
  const int64_t i = (int64_t)blockIdx.x * NT + ty;
  double best = INFINITY;
  int32_t best_j = 0;
  bool seen_nan = false;
  for (int64_t j0 = 0; j0 < n_ex; j0 += NT) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int64_t d0 = 0; d0 < d; d0 += ND) {
      __syncthreads();
      for (int t = tid; t < NT * ND; t += NT * NT) {
        const int rr = t / ND, cc = t % ND;
        const int64_t ia = (int64_t)blockIdx.x * NT + rr, jb = j0 + rr, dd = d0 + cc;
        sa[rr][cc] = (ia < n_test && dd < d) ? A[ia * lda + dd] : 0.0;
        sb[rr][cc] = (jb < n_ex && dd < d) ? B[jb * ldb + dd] : 0.0;
      }
      __syncthreads();
      const int lim = (int)(d - d0 < ND ? d - d0 : ND);
      for (int c = 0; c < lim; c++) accumulate<MEASURE>(sa[ty][c], sb[tx][c], s0, s1, s2);
    }
    const int64_t j = j0 + tx;
    if (i < n_test && j < n_ex) {
      const double v = finish<MEASURE>(s0, s1, s2);
      if (D != nullptr) D[i * ldd + j] = v;
      // np.argmin: first index of the minimum; a NaN, once met, wins (numpy propagates the first NaN)
      if (!seen_nan) {
        if (v != v) { seen_nan = true; best = v; best_j = (int32_t)j; }
        else if (v < best) { best = v; best_j = (int32_t)j; }
      }
    }
  }
  if (argmin == nullptr && minval == nullptr) return;
  rmin[ty][tx] = best;
  ridx[ty][tx] = (i < n_test && tx < n_ex) ? best_j : INT32_MAX;
  __syncthreads();
  if (tx == 0 && i < n_test) {
    double b = rmin[ty][0];
    int32_t bj = ridx[ty][0];
    bool nan = b != b;
    for (int t = 1; t < NT; t++) {
      const double v = rmin[ty][t];
      const int32_t vj = ridx[ty][t];
      if (vj == INT32_MAX) continue;
      const bool vnan = v != v;
      if (vnan) { if (!nan || vj < bj) { nan = true; b = v; bj = vj; } }
      else if (!nan && (v < b || (v == b && vj < bj))) { b = v; bj = vj; }
    }
    if (argmin) argmin[i] = bj;
    if (minval) minval[i] = b;
  }
}

}  // namespace
}  // namespace klnmf

using namespace klnmf;

extern "C" int klnmf_pairwise_host(int device, int measure, int64_t n_test, int64_t n_ex, int64_t d, const double *A,
                                   int64_t lda, const double *B, int64_t ldb, double *dists, int64_t ldd,
                                   int32_t *argmin, double *minval) {
  KL_CHECK(n_test >= 0 && n_ex >= 0 && d >= 0 && A && B, KLNMF_EINVAL, "klnmf_pairwise_host: bad arguments");
  KL_CHECK(measure >= KLNMF_MEASURE_KL && measure <= KLNMF_MEASURE_COSINE_DIFF, KLNMF_EINVAL,
           "klnmf_pairwise_host: unknown measure %d", measure);
  KL_CHECK(lda >= d && ldb >= d && (!dists || ldd >= n_ex), KLNMF_EINVAL, "klnmf_pairwise_host: leading dimension too small");
  KL_CHECK(!argmin || n_ex > 0, KLNMF_EINVAL, "klnmf_pairwise_host: argmin of an empty example set (numpy raises ValueError)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
    cudaGetLastError();
    set_error("klnmf_pairwise_host: no CUDA device %d; the library has no CPU path", device);
    return KLNMF_ENODEVICE;
  }
  if (n_test == 0) return KLNMF_OK;
  KL_CUDA(cudaSetDevice(device));
  double *dA = nullptr, *dB = nullptr, *dD = nullptr, *dM = nullptr;
  int32_t *dI = nullptr;
  int rc = KLNMF_OK;
  auto fail = [&](const char *what, cudaError_t e) {
    set_error("klnmf_pairwise_host: %s -> %s", what, cudaGetErrorString(e));
    rc = KLNMF_ECUDA;
  };
  cudaError_t e;
  const size_t ca = (size_t)(d > 0 ? d : 1);
  if ((e = cudaMalloc(&dA, (size_t)n_test * ca * 8)) != cudaSuccess) fail("cudaMalloc A", e);
  if (rc == KLNMF_OK && (e = cudaMalloc(&dB, (size_t)(n_ex > 0 ? n_ex : 1) * ca * 8)) != cudaSuccess) fail("cudaMalloc B", e);
  if (rc == KLNMF_OK && dists && n_ex > 0 && (e = cudaMalloc(&dD, (size_t)n_test * n_ex * 8)) != cudaSuccess) fail("cudaMalloc D", e);
  if (rc == KLNMF_OK && argmin && (e = cudaMalloc(&dI, (size_t)n_test * 4)) != cudaSuccess) fail("cudaMalloc argmin", e);
  if (rc == KLNMF_OK && minval && (e = cudaMalloc(&dM, (size_t)n_test * 8)) != cudaSuccess) fail("cudaMalloc min", e);
  if (rc == KLNMF_OK && d > 0) {
    if ((e = cudaMemcpy2D(dA, d * 8, A, lda * 8, d * 8, n_test, cudaMemcpyHostToDevice)) != cudaSuccess) fail("H2D A", e);
    if (rc == KLNMF_OK && n_ex > 0 &&
        (e = cudaMemcpy2D(dB, d * 8, B, ldb * 8, d * 8, n_ex, cudaMemcpyHostToDevice)) != cudaSuccess)
      fail("H2D B", e);
  }
  if (rc == KLNMF_OK && n_ex > 0) {
    dim3 block(NT, NT), grid((unsigned)ceil_div(n_test, NT));
    switch (measure) {
      case KLNMF_MEASURE_KL: pairwise_kernel<KLNMF_MEASURE_KL><<<grid, block>>>(dA, d, dB, d, n_test, n_ex, d, dD, n_ex, dI, dM); break;
      case KLNMF_MEASURE_REV_KL: pairwise_kernel<KLNMF_MEASURE_REV_KL><<<grid, block>>>(dA, d, dB, d, n_test, n_ex, d, dD, n_ex, dI, dM); break;
      case KLNMF_MEASURE_SYM_KL: pairwise_kernel<KLNMF_MEASURE_SYM_KL><<<grid, block>>>(dA, d, dB, d, n_test, n_ex, d, dD, n_ex, dI, dM); break;
      case KLNMF_MEASURE_FROBENIUS: pairwise_kernel<KLNMF_MEASURE_FROBENIUS><<<grid, block>>>(dA, d, dB, d, n_test, n_ex, d, dD, n_ex, dI, dM); break;
      default: pairwise_kernel<KLNMF_MEASURE_COSINE_DIFF><<<grid, block>>>(dA, d, dB, d, n_test, n_ex, d, dD, n_ex, dI, dM); break;
    }
    if ((e = cudaGetLastError()) != cudaSuccess) fail("launch", e);
    if (rc == KLNMF_OK && (e = cudaDeviceSynchronize()) != cudaSuccess) fail("kernel", e);
    if (rc == KLNMF_OK && dD &&
        (e = cudaMemcpy2D(dists, ldd * 8, dD, n_ex * 8, n_ex * 8, n_test, cudaMemcpyDeviceToHost)) != cudaSuccess)
      fail("D2H D", e);
    if (rc == KLNMF_OK && dI && (e = cudaMemcpy(argmin, dI, (size_t)n_test * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) fail("D2H argmin", e);
    if (rc == KLNMF_OK && dM && (e = cudaMemcpy(minval, dM, (size_t)n_test * 8, cudaMemcpyDeviceToHost)) != cudaSuccess) fail("D2H min", e);
  }
  if (dA) cudaFree(dA);
  if (dB) cudaFree(dB);
  if (dD) cudaFree(dD);
  if (dI) cudaFree(dI);
  if (dM) cudaFree(dM);
  return rc;
}
