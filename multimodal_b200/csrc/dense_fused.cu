// Fused coefficient half-step of KL-NMF for k <= 128 (tcgen05 / TMEM / TMA, sm_100a):
//
//     S = W.H            (nmf.py:336)      first contraction, accumulator in TMEM
//     Q = (X+eps)/(S+eps), KL += X*log Q - X + S   (nmf.py:325-336, metrics.py:18-20) in the TMEM epilogue
//     G += Q.H^T         (nmf.py:342)      second contraction, fed from shared memory -- Q never reaches HBM
//     W' = W (.) G       (nmf.py:343)      final epilogue of the row block
//
// One persistent CTA per SM owns row blocks of 128 samples.  The 128 x KP block of W stays in shared memory as
// the A operand of the first contraction for the whole sweep over the features; the dictionary streams past
// in steps of 32 features, once as H^T rows (B of S = W.H, K-major over k) and once as H rows (B of
// G = Q.H^T, K-major over f).  TMEM holds G (KP columns, lives for the whole row block) and two 32-column
// S buffers, so the ratio epilogue of step j overlaps the first contraction of step j+1 and the second
// contraction of step j-1.  HBM traffic per iteration: X once + W read + W' written -- the compulsory bytes of
// SURVEY 8d -- instead of X + 3 x Q of the unfused form.  WRITE_Q (fit): the ratio tile additionally leaves
// through a TMA store for the dictionary numerator N += W'^T.Q (nmf.py:349), which needs the finished W'.
//
//   warp 0      TMA producer : W block per row block; per step the two dictionary tiles
//   warp 1      MMA issuer   : whole warp walks the loop, one elected lane issues tcgen05.mma / commit
//   warps 2..9  epilogue     : two groups on alternate steps: tcgen05.ld 32x32b.x32 -> ratio / objective -> Q tile
//                              (128B-swizzled A operand of the second contraction)
//   warp 10     X loader     : 128 x 32 chunks of X by TMA into a swizzled ring
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace klnmf {

namespace {

constexpr int FBM = 128;                 // samples per row block (UMMA M)
constexpr int FBN = 32;                  // features per step = one 128-byte swizzle span
constexpr int F_THREADS = 352;
constexpr int F_EPI_WARPS = 8;
constexpr int CHUNK_BYTES = FBM * FBN * 4;   // a 128 x 32 fp32 tile: X chunk, Q tile, one K block of W

// TSW: the W block lives in TMEM (A operand of the first contraction read from tensor memory, "TS" form of
// tcgen05.mma) instead of shared memory: no 4 KB shared-memory read of A per N = 32 instruction, and its
// 64 KB go to dictionary stages.
template <int KP, bool TSW, int V = 0>
struct FCfg {
  static constexpr int KB = KP / 32;                       // K blocks of the first contraction
  static constexpr int W_BYTES = TSW ? 0 : KB * CHUNK_BYTES;   // resident A operand
  static constexpr int H1_BYTES = KB * FBN * 32 * 4;       // H^T tile: KB blocks of 32 feature rows x 32 k
  static constexpr int H2_BYTES = KP * FBN * 4;            // H tile: KP rows x 32 features
  static constexpr int HSTAGE_BYTES = H1_BYTES + H2_BYTES;
  // the X ring has to cover the HBM latency by itself: 44 GB/s per SM x ~2 us = 88 KB in flight at the HBM
  // roofline, so every byte the operands do not need goes to X chunks (16 KB each)
  // V = 1 (k > 64, W in TMEM): five dictionary stages and a two-chunk X ring instead of four and four
  static constexpr int HS = KP >= 128 ? (TSW ? (V ? 5 : 4) : 3) : (TSW ? 6 : 4);
  static constexpr int XB = KP >= 128 ? (TSW ? (V ? 2 : 4) : 2) : (TSW ? 6 : 6);
  static constexpr int SMEM_BYTES = W_BYTES + HS * HSTAGE_BYTES + 2 * CHUNK_BYTES + XB * CHUNK_BYTES + 1024 + 256;
  static constexpr int TMEM_USED = KP + 64 + (TSW ? KP : 0);
  static constexpr int TMEM_COLS = TMEM_USED <= 128 ? 128 : (TMEM_USED <= 256 ? 256 : 512);
  static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");
};

struct FusedParams {
  int64_t M, F;
  int n_blocks, n_steps;
  const float *W;          // current coefficients (multiplied into G in the final epilogue)
  int64_t ldw;
  float *Wout;
  int64_t ldwo, w_cols;    // columns < w_cols are written (multiple of 16)
  double *kl;
  const int *stop;
  int *err;
  int only_kl;
  float qshift;            // the ratio tile holds q - qshift (centered ratio, api.cu)
  const float *colbias;    // W' = W (.) (G + colbias[component])
  int lookahead;           // steps the first contraction runs ahead of the second one (1 or 2)
  const float *Wlo;        // TF32R: low parts of W (W itself is the TF32-exact high part); nullptr otherwise
  float *Wout_lo;          // TF32R: low parts of W'
  int accurate;            // TF32R: cancellation-free objective, centered ratio stored as u, tile rounded to TF32
};

__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t v[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (M lanes x 8 columns of 32-bit K elements) comes from TMEM
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int KP, bool WRITE_Q, bool TSW, int V>
__global__ void __launch_bounds__(F_THREADS, 1)
fused_coef_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmHt,
                  const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmQ, const FusedParams p) {
  using C = FCfg<KP, TSW, V>;
  constexpr int KB = C::KB, HS = C::HS, XB = C::XB;
  if (p.stop != nullptr && *p.stop != 0) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t w_s = smem_base;
  const uint32_t h_s = w_s + C::W_BYTES;
  const uint32_t q_s = h_s + HS * C::HSTAGE_BYTES;
  const uint32_t x_s = q_s + 2 * CHUNK_BYTES;
  const uint32_t bar_base = x_s + XB * CHUNK_BYTES;
  uint8_t *q_gen = smem_gen + (q_s - smem_base);
  uint8_t *x_gen = smem_gen + (x_s - smem_base);
  // barriers: w_full | w_empty | g_full | g_empty | h_full[HS] | h_empty[HS] | s_full[2] | s_empty[2] |
  //           q_full[2] | q_empty[2] | x_full[XB] | x_empty[XB] | tmem_ptr
  const uint32_t w_full = bar_base, w_empty = bar_base + 8, g_full = bar_base + 16, g_empty = bar_base + 24;
  auto h_full = [&](int s) { return bar_base + 32u + 8u * s; };
  auto h_empty = [&](int s) { return bar_base + 32u + 8u * (HS + s); };
  const uint32_t b2 = bar_base + 32u + 16u * HS;
  auto s_full = [&](int a) { return b2 + 8u * a; };
  auto s_empty = [&](int a) { return b2 + 16u + 8u * a; };
  auto q_full = [&](int a) { return b2 + 32u + 8u * a; };
  auto q_empty = [&](int a) { return b2 + 48u + 8u * a; };
  auto x_full = [&](int b) { return b2 + 64u + 8u * b; };
  auto x_empty = [&](int b) { return b2 + 64u + 8u * (XB + b); };
  const uint32_t tmem_ptr_addr = b2 + 64u + 16u * XB;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_ptr_addr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmHt); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmX);
    if (WRITE_Q) tma_prefetch_desc(&tmQ);
    mbar_init(w_full, TSW ? F_EPI_WARPS : 1); mbar_init(w_empty, 1); mbar_init(g_full, 1); mbar_init(g_empty, F_EPI_WARPS);
    for (int s = 0; s < HS; s++) { mbar_init(h_full(s), 1); mbar_init(h_empty(s), 1); }
    for (int a = 0; a < 2; a++) {
      mbar_init(s_full(a), 1); mbar_init(s_empty(a), F_EPI_WARPS / 2);
      mbar_init(q_full(a), WRITE_Q ? 1 : F_EPI_WARPS / 2); mbar_init(q_empty(a), 1);
    }
    for (int b = 0; b < XB; b++) { mbar_init(x_full(b), 1); mbar_init(x_empty(b), F_EPI_WARPS / 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_addr, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t g_tmem = tmem_base;                 // columns [0, KP)
  const uint32_t s_tmem = tmem_base + KP;            // two 32-column S buffers
  const uint32_t w_tmem = tmem_base + KP + 64;       // TSW: the W block, KP columns

  const int nsteps = p.n_steps;

  if (warp == 0) {
    // =============================== TMA producer: W block, dictionary tiles ===============================
    if (lane == 0) {
      uint32_t hc = 0, rbc = 0;
      for (int rb = blockIdx.x; rb < p.n_blocks; rb += gridDim.x, rbc++) {
        if (!TSW) {
          mbar_wait(w_empty, (rbc & 1u) ^ 1u, p.err, 1);
          mbar_expect_tx(w_full, C::W_BYTES);
          for (int kb = 0; kb < KB; kb++) tma_load_2d(w_s + kb * CHUNK_BYTES, &tmW, w_full, kb * 32, rb * FBM);
        }
        for (int j = 0; j < nsteps; j++, hc++) {
          const uint32_t s = hc % HS, ph = (hc / HS) & 1u;
          mbar_wait(h_empty(s), ph ^ 1u, p.err, 2);
          mbar_expect_tx(h_full(s), C::HSTAGE_BYTES);
          const uint32_t dst = h_s + s * C::HSTAGE_BYTES;
          for (int kb = 0; kb < KB; kb++) tma_load_2d(dst + kb * 4096, &tmHt, h_full(s), kb * 32, j * FBN);
          tma_load_2d(dst + C::H1_BYTES, &tmH, h_full(s), j * FBN, 0);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // instruction descriptors: D=f32, A=B=tf32, both K-major, N>>3, M>>4
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FBN >> 3) << 17) | ((uint32_t)(FBM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(FBM >> 4) << 24);
    const uint32_t hi = desc_hi(1024u, 2u);          // K-major SWIZZLE_128B, 8-row groups 1024 B apart
    const uint32_t w_lo = desc_lo(w_s, 16u), h_lo = desc_lo(h_s, 16u), q_lo = desc_lo(q_s, 16u);
    uint32_t c1 = 0, c2 = 0, rbc = 0;                // first / second contractions issued, row blocks done
    for (int rb = blockIdx.x; rb < p.n_blocks; rb += gridDim.x, rbc++) {
      mbar_wait(w_full, rbc & 1u, p.err, 3);
      if (TSW) tc_fence_after();
      // the first contraction runs LA steps ahead of the second one: with LA = 2, S(j+2) only needs the S buffer
      // back (released as soon as epilogue j has read it), not the end of epilogue j -- at the price of one more
      // dictionary stage in flight
      const int LA = p.lookahead;
      for (int j = 0; j < nsteps + LA; j++) {
        if (j < nsteps) {
          // ---- S[a] = W . H^T-tile(j) ----
          const uint32_t s = c1 % HS, a = c1 & 1u;
          mbar_wait(h_full(s), (c1 / HS) & 1u, p.err, 4);
          mbar_wait(s_empty(a), ((c1 >> 1) & 1u) ^ 1u, p.err, 5);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bl = h_lo + ((s * C::HSTAGE_BYTES) >> 4);
#pragma unroll
            for (int kb = 0; kb < KB; kb++)
#pragma unroll
              for (int kk = 0; kk < 4; kk++) {
                const uint64_t db = desc_pack(bl + ((kb * 4096 + kk * 32) >> 4), hi);
                if (TSW) umma_tf32_ts(s_tmem + a * FBN, w_tmem + kb * 32 + kk * 8, db, idesc1, (kb | kk) ? 1u : 0u);
                else umma_tf32<1>(s_tmem + a * FBN, desc_pack(w_lo + ((kb * CHUNK_BYTES + kk * 32) >> 4), hi), db, idesc1,
                                  (kb | kk) ? 1u : 0u);
              }
            umma_commit<1>(s_full(a));
          }
          __syncwarp();
          c1++;
        }
        if (j >= LA) {
          // ---- G += Q(j-LA) . H-tile(j-LA)^T ----
          const uint32_t s = c2 % HS, b = c2 & 1u;
          mbar_wait(q_full(b), (c2 >> 1) & 1u, p.err, 6);
          if (j == LA) mbar_wait(g_empty, (rbc & 1u) ^ 1u, p.err, 7);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t al = q_lo + ((b * CHUNK_BYTES) >> 4);
            const uint32_t bl = h_lo + ((s * C::HSTAGE_BYTES + C::H1_BYTES) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
              umma_tf32<1>(g_tmem, desc_pack(al + ((kk * 32) >> 4), hi), desc_pack(bl + ((kk * 32) >> 4), hi), idesc2,
                           (j > LA || kk > 0) ? 1u : 0u);
            umma_commit<1>(h_empty(s));        // dictionary stage free once both of its contractions retired
            umma_commit<1>(q_empty(b));
            if (j == nsteps + LA - 1) { umma_commit<1>(g_full); umma_commit<1>(w_empty); }
          }
          __syncwarp();
          c2++;
        }
      }
    }
  } else if (warp == 10) {
    // =============================== X loader ===============================
    if (lane == 0) {
      uint32_t g = 0;
      for (int rb = blockIdx.x; rb < p.n_blocks; rb += gridDim.x)
        for (int j = 0; j < nsteps; j++, g++) {
          const uint32_t b = g % XB, ph = (g / XB) & 1u;
          mbar_wait(x_empty(b), ph ^ 1u, p.err, 8);
          mbar_expect_tx(x_full(b), CHUNK_BYTES);
          tma_load_2d(x_s + b * CHUNK_BYTES, &tmX, x_full(b), j * FBN, rb * FBM);
        }
    }
  } else {
    // =============================== epilogue ===============================
    // Two groups of four warps (one warp per TMEM lane quarter) take alternate steps, so that the MUFU-bound
    // arithmetic of one group overlaps the TMEM / barrier latencies of the other.  Group g owns S buffer g and
    // Q buffer g.
    const int e = warp - 2;                 // 0..7
    const int quarter = warp & 3;           // TMEM lanes this warp may touch: [32*quarter, +32)
    const int grp = e >> 2;
    const int r = quarter * 32 + lane;      // row inside the block = TMEM lane
    const uint32_t sw = (uint32_t)(r & 7);
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    uint32_t c = 0, rbc = 0;
    double kl = 0.0;
    for (int rb = blockIdx.x; rb < p.n_blocks; rb += gridDim.x, rbc++) {
      float kl_blk = 0.f;
      if (TSW) {
        // ---- the W block of this row block goes to TMEM: lane = row, column = component ----
        mbar_wait(w_empty, (rbc & 1u) ^ 1u, p.err, 13);     // first contractions of the previous block retired
        tc_fence_after();
        const int64_t row = (int64_t)rb * FBM + r;
#pragma unroll 1
        for (int cc = 0; cc < KP / 32; cc++) {
          const int col0 = grp * (KP / 2) + cc * 16;
          uint32_t wv[16];
          if (row < p.M && col0 < p.ldw) {
            const float *wi = p.W + row * p.ldw + col0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float4 w = __ldg(reinterpret_cast<const float4 *>(wi + 4 * i));
              wv[4 * i] = __float_as_uint(w.x); wv[4 * i + 1] = __float_as_uint(w.y);
              wv[4 * i + 2] = __float_as_uint(w.z); wv[4 * i + 3] = __float_as_uint(w.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; i++) wv[i] = 0u;
          }
          tmem_st16(w_tmem + lane_addr + col0, wv);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(w_full);
      }
      const int j_pf = nsteps > 8 ? nsteps - 8 : 0;
#pragma unroll 1
      for (int j = 0; j < nsteps; j++, c++) {
        if (TSW && j == j_pf) {
          // the next row block's coefficients are pulled into L2 a few steps before this warp copies them to TMEM
          const int64_t prow = (int64_t)(rb + gridDim.x) * FBM + r;
          if (rb + (int)gridDim.x < p.n_blocks && prow < p.M) {
#pragma unroll
            for (int i = 0; i < KP / 64; i++)
              if (grp * (KP / 2) + 32 * i < p.ldw)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + prow * p.ldw + grp * (KP / 2) + 32 * i));
          }
        }
        if ((int)(c & 1u) != grp) continue;
        const uint32_t ph2 = (c >> 1) & 1u, xb = c % XB, phx = (c / XB) & 1u;
        mbar_wait(s_full(grp), ph2, p.err, 9);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_issue(s_tmem + lane_addr + grp * FBN, v);
        mbar_wait(x_full(xb), phx, p.err, 10);
        float x[32];
        const uint8_t *xrow = x_gen + xb * CHUNK_BYTES + r * 128;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float4 t = *reinterpret_cast<const float4 *>(xrow + (((uint32_t)i ^ sw) << 4));
          x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
        }
        tmem_ld32_wait(v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive_relaxed(s_empty(grp)); mbar_arrive(x_empty(xb)); }
        // rows >= M and columns >= F hold x = 0, s = 0 (TMA zero fill): q = 1, the term is exactly 0 and the
        // zero-filled dictionary columns keep it out of G
        kl_blk += ratio_chunk32_dispatch(x, v, p.qshift, p.accurate);
        if (!p.only_kl) {
          mbar_wait(q_empty(grp), ph2 ^ 1u, p.err, 11);
          if (WRITE_Q) {
            // the TMA store of this group's previous tile must have finished reading the buffer as well
            if (quarter == 0 && lane == 0) bulk_wait_read0();
            named_bar(1 + grp, 128);
          }
          uint8_t *qrow = q_gen + grp * CHUNK_BYTES + r * 128;
#pragma unroll
          for (int i = 0; i < 8; i++)
            *reinterpret_cast<float4 *>(qrow + (((uint32_t)i ^ sw) << 4)) =
                make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
          fence_proxy_async();
        }
        if (WRITE_Q) {
          // fit: the finished 128 x 32 ratio tile also leaves for HBM (one TMA store per step), the dictionary
          // numerator N += W'^T.Q (nmf.py:349) reads it once W' of the whole panel exists
          named_bar(1 + grp, 128);
          if (quarter == 0 && lane == 0) {
            if (!p.only_kl) tma_store_2d(&tmQ, q_s + grp * CHUNK_BYTES, j * FBN, rb * FBM);
            mbar_arrive(q_full(grp));
          }
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(q_full(grp));
        }
      }
      kl += (double)kl_blk;
      // ---- final epilogue of the row block: W' = W (.) G, each group one half of the columns ----
      mbar_wait(g_full, rbc & 1u, p.err, 12);
      tc_fence_after();
      const int64_t row = (int64_t)rb * FBM + r;
#pragma unroll 1
      for (int cc = 0; cc < KP / 32; cc++) {
        const int col0 = grp * (KP / 2) + cc * 16;
        uint32_t v[16], wt[16];
        tmem_ld16_issue(g_tmem + lane_addr + col0, v);
        if (TSW) tmem_ld16_issue(w_tmem + lane_addr + col0, wt);    // W is still in its TMEM block (it left L2 long ago)
        tmem_ld16_wait(v);
        if (TSW) tmem_ld16_wait(wt);
        if (!p.only_kl && row < p.M && col0 < p.w_cols) {
          const float *wi = p.W + row * p.ldw + col0;
          float *wo = p.Wout + row * p.ldwo + col0;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float4 w;
            if (TSW) w = make_float4(__uint_as_float(wt[4 * i]), __uint_as_float(wt[4 * i + 1]), __uint_as_float(wt[4 * i + 2]),
                                     __uint_as_float(wt[4 * i + 3]));
            else w = __ldg(reinterpret_cast<const float4 *>(wi + 4 * i));
            if (p.Wlo) {
              const float4 l = __ldg(reinterpret_cast<const float4 *>(p.Wlo + row * p.ldw + col0 + 4 * i));
              w.x += l.x; w.y += l.y; w.z += l.z; w.w += l.w;
            }
            float4 g = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
            if (p.colbias) {
              // centered ratio: G = (Q-1).H^T + rowsum(H), clamped at zero like the unfused epilogue (dense_tc.cu)
              const float4 b = __ldg(reinterpret_cast<const float4 *>(p.colbias + col0 + 4 * i));
              g.x = fmaxf(g.x + b.x, 0.f); g.y = fmaxf(g.y + b.y, 0.f); g.z = fmaxf(g.z + b.z, 0.f); g.w = fmaxf(g.w + b.w, 0.f);
            }
            float4 o = make_float4(w.x * g.x, w.y * g.y, w.z * g.z, w.w * g.w);
            if (p.Wout_lo) {
              const float4 h = make_float4(tf32_round(o.x), tf32_round(o.y), tf32_round(o.z), tf32_round(o.w));
              *reinterpret_cast<float4 *>(p.Wout_lo + row * p.ldwo + col0 + 4 * i) =
                  make_float4(o.x - h.x, o.y - h.y, o.z - h.z, o.w - h.w);
              o = h;
            }
            *reinterpret_cast<float4 *>(wo + 4 * i) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(g_empty);
    }
    if (WRITE_Q && quarter == 0 && lane == 0) bulk_wait_all();
    if (p.kl != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
      if (lane == 0) atomicAdd(p.kl, kl);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, C::TMEM_COLS);
  }
}

struct FusedState {
  int *err_dev = nullptr;
};

template <int KP, bool TSW, int V = 0, bool WRITE_Q = false>
int launch_fused(klnmf_ctx *ctx, const FusedDesc &d, FusedParams p) {
  using C = FCfg<KP, TSW, V>;
  CUtensorMap tmW, tmHt, tmH, tmX, tmQ;
  // every operand is K-major with 128B swizzle; extents are the stored (zero padded) ones, everything beyond
  // them is zero-filled by TMA
  KL_TRY(make_map(&tmW, d.W, d.ldw, d.M, d.ldw, FBM, false));        // box 32 k x 128 rows
  KL_TRY(make_map(&tmHt, d.Ht, d.ldht, d.F, d.ldht, 32, false));     // box 32 k x 32 feature rows
  KL_TRY(make_map(&tmH, d.H, d.F, d.K, d.ldh, KP, false));           // box 32 features x KP rows
  KL_TRY(make_map(&tmX, d.X, d.F, d.M, d.ldx, FBM, false));          // box 32 features x 128 rows
  tmQ = tmX;
  if (WRITE_Q) KL_TRY(make_map(&tmQ, d.Q, d.F, d.M, d.ldq, FBM, false));   // box 32 features x 128 rows, clipped at the edges
  p.n_blocks = (int)ceil_div(d.M, FBM);
  p.n_steps = (int)ceil_div(d.F, FBN);
  if (p.n_blocks == 0 || p.n_steps == 0) return KLNMF_OK;
  const int grid = p.n_blocks < ctx->sm_count ? p.n_blocks : ctx->sm_count;
  auto kern = fused_coef_kernel<KP, WRITE_Q, TSW, V>;
  // per DEVICE: function attributes live in the device's context, and one process may drive several GPUs
  static bool attr_done[64] = {};
  if (!attr_done[ctx->device & 63]) {
    KL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done[ctx->device & 63] = true;
  }
  kern<<<grid, F_THREADS, C::SMEM_BYTES, ctx->stream>>>(tmW, tmHt, tmH, tmX, tmQ, p);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

}  // namespace

bool fused_supported(const klnmf_ctx *ctx, int fit) {
  const bool off = getenv("KLNMF_FUSED") && atoi(getenv("KLNMF_FUSED")) == 0;   // read per call: tests toggle it
  // k <= 128 (this file) and 128 < k <= 256 on CTA pairs (dense_fused256.cu): fit and transform
  const bool off256 = getenv("KLNMF_FUSED256") && atoi(getenv("KLNMF_FUSED256")) == 0;
  // the one-pass modes: TF32, and TF32R (rounded operands: the (hi, lo) state's high parts; centered, rounded ratio tile)
  return !off && (ctx->mode == KLNMF_MODE_TF32 || ctx->mode == KLNMF_MODE_TF32R) && !ctx->sparse && !ctx->debug_simt &&
         (ctx->k <= 128 || (!off256 && ctx->k <= 256));
}

int fused_coef_step(klnmf_ctx *ctx, const FusedDesc &d) {
  if (d.M <= 0) return KLNMF_OK;
  FusedState *st = (FusedState *)ctx->fused;
  if (!st) {
    st = new FusedState();
    if (cudaMalloc((void **)&st->err_dev, 4) != cudaSuccess) {
      delete st;
      set_error("fused_coef_step: cudaMalloc failed");
      return KLNMF_ENOMEM;
    }
    cudaMemsetAsync(st->err_dev, 0, 4, ctx->stream);
    ctx->fused = st;
  }
  if (d.K > 128) return fused_coef_step256(ctx, d, st->err_dev);
  KL_CHECK(d.K <= 128 && d.ldw % 32 == 0 && d.ldht % 32 == 0, KLNMF_EINVAL, "fused_coef_step: k=%lld not supported",
           (long long)d.K);
  FusedParams p{};
  p.M = d.M; p.F = d.F;
  p.W = (const float *)d.W; p.ldw = d.ldw;
  p.Wout = (float *)d.Wout; p.ldwo = d.ldwo;
  p.w_cols = d.ldw < d.ldwo ? d.ldw : d.ldwo;
  p.kl = d.kl; p.stop = d.stop; p.err = st->err_dev; p.only_kl = d.only_kl;
  p.qshift = d.qshift; p.colbias = d.colbias;
  p.Wlo = (const float *)d.Wlo; p.Wout_lo = (float *)d.Wout_lo; p.accurate = d.accurate;
  // The W block lives in TMEM ("TS" form of tcgen05.mma).  The round-1 variant that kept it in shared memory
  // (KLNMF_FUSED_TS=0) is gone: tools/fused_determinism.py showed it racy at k > 64 (a handful of rows off by 1e-2 in
  // one run out of two, both arithmetic modes -- within the old 3e-3 tolerance, hence unnoticed).
  const int v = getenv("KLNMF_FUSED_V") ? atoi(getenv("KLNMF_FUSED_V")) : 0;
  p.lookahead = d.K <= 64 ? 2 : 1;
  if (getenv("KLNMF_FUSED_LA")) p.lookahead = atoi(getenv("KLNMF_FUSED_LA")) == 2 ? 2 : 1;
  if (d.Q != nullptr) {   // fit: the ratio panel is written for the numerator contraction
    KL_CHECK(d.ldq >= round_up(d.F, 4), KLNMF_EINVAL, "fused_coef_step: ratio panel leading dimension too small");
    if (d.K <= 64) return launch_fused<64, true, 0, true>(ctx, d, p);
    return launch_fused<128, true, 0, true>(ctx, d, p);
  }
  if (d.K <= 64) return launch_fused<64, true>(ctx, d, p);
  if (v == 1) return launch_fused<128, true, 1>(ctx, d, p);
  return launch_fused<128, true>(ctx, d, p);
}

void fused_release(klnmf_ctx *ctx) {
  FusedState *st = (FusedState *)ctx->fused;
  if (!st) return;
  if (st->err_dev) cudaFree(st->err_dev);
  delete st;
  ctx->fused = nullptr;
}

}  // namespace klnmf
