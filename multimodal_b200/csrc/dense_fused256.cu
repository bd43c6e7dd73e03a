// Fused coefficient half-step for 128 < k <= 256 on a CLUSTER OF TWO CTAs that share one block of 128 samples.
//
// At k = 256 the coefficient accumulator G (128 lanes x 256 FP32 columns) and the resident W block (256 columns
// as the TMEM A operand) fill a whole TMEM by themselves, so the single-CTA kernel of dense_fused.cu has no
// column left for S.  Here the two SMs of a cluster split the COLUMNS of both contractions over their TMEMs:
//
//   CTA r:  S_r = W . H[:, 32 features of the step]      (first contraction, N = 32: its half of the 64-feature step)
//           Q_r = (X_r+eps)/(S_r+eps), objective          (TMEM epilogue; X_r = its 32 columns of X)
//           Q_r is written into the Q tile of BOTH CTAs  (local st.shared, then one 4 KB cp.async.bulk per warp
//                                                         from its own shared memory into the peer's)
//           G_r += Q . H[128 r : 128 r + 128, step]^T     (second contraction, N = 128: its half of the components)
//           W'[:, 128 r : 128 r + 128] = W (.) G_r        (final epilogue, W read back from TMEM)
//
// No MMA work is duplicated, Q crosses the cluster's distributed shared memory (16 KB per step and direction) and
// never HBM; each CTA holds the full W block in TMEM (A operand of the first contraction, TS form).  The protocol
// is the one of dense_fused.cu with these changes: q_full collects the four local epilogue warps plus the 16 KB the
// peer's bulk copies complete on it (complete_tx, like a TMA load); q_empty collects the tcgen05.commit of both
// issuers; the H^T and H tiles have their own rings and producer warps (the first contraction frees its tile three
// steps before the second one does); four S accumulators let the first contraction run three steps ahead.
//
// History (cfg3: n = 1e6, f = 4096, k = 256; the unfused form takes 9.69 ms), profiles/r1_fused256_v*_summary.csv:
//   19.9 ms  32-feature steps, Q stored into the peer with st.shared::cluster + mbarrier.arrive.release.cluster: the
//            all-space fence.proxy.async and the cluster-scope release (MEMBAR.ALL.GPU, ERRBAR) were 31 % of all stalls
//   11.9 ms  Q exchanged by cp.async.bulk.shared::cluster (half tiles in SWIZZLE_64B); the issuer was then the limit:
//            32 N=16 + 4 N=128 MMAs per step
//   10.1 ms  64-feature steps (N = 32: 20 MMAs per 32 features), separate H^T / H rings
//    9.7 ms  W re-read from TMEM in the final epilogue, next W block prefetched into L2, lookahead 3
// What paces it now (timing experiments, KLNMF_F256_DBG): without MMAs, ratio math and exchange a step still takes
// 1700 clk, and halving the dictionary bytes does not change that -- it is the round trip
// "tcgen05.commit -> producer -> TMA -> L2 -> full barrier" of a ring that shared memory limits to two stages per
// operand (2 x 32 KB H^T + 2 x 32 KB H + 2 x 32 KB Q + 2 x 16 KB X = 224 KB).  Four half-size stages were slower.
// Reference lines as in dense_fused.cu (nmf.py:325-343, metrics.py:18-20).  Fit (write_q): every warp also stores
// its 32 x 32 box of the ratio tile by TMA, once, for the numerator contraction N += W'^T.Q (nmf.py:349), which
// needs the finished W' of the whole panel.  TF32R (accurate): the TMEM block of W holds the round-to-nearest high
// parts, the ratio tile is centered (u = q - 1) and rounded to nearest, the objective takes its cancellation-free form,
// W' = (W_hi + W_lo) (.) max(G + rowsum(H), 0) leaves as a (hi, lo) pair.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace klnmf {

namespace {

constexpr int PBM = 128;                   // samples per row block
constexpr int PBN = 64;                    // features per step (32 per CTA)
constexpr int PHN = 32;                    // features per step and CTA
constexpr int PKP = 256;                   // padded components
constexpr int PKH = 128;                   // components per CTA (columns of G)
constexpr int P_THREADS = 384;
constexpr int P_EPI_WARPS = 8;
constexpr int QHALF_BYTES = PBM * PHN * 4;             // 16 KB: the half tile one CTA produces, 128 x 32 K-major, 128B swizzle
constexpr int QTILE_BYTES = 2 * QHALF_BYTES;           // 32 KB: the 128 x 64 ratio tile
constexpr int XHALF_BYTES = PBM * PHN * 4;             // 16 KB: 128 rows x 32 features (128B swizzle)
constexpr int PH1_BYTES = (PKP / 32) * PHN * 32 * 4;   // 32 KB: H^T tile, 8 K blocks of 32 feature rows x 32 k
constexpr int PH2_BYTES = PKH * PBN * 4;               // 32 KB: H tile, two boxes of 128 component rows x 32 features
constexpr int PS1 = 2;                                 // H^T stages (freed by the first contraction; 4 half-size stages were slower)
constexpr int PS2 = 2;                                 // H stages (freed by the second contraction)
constexpr int PXB = 2;                                 // X half-chunks in flight (one per epilogue group)
constexpr int PLA = 3;                                 // steps the first contraction runs ahead of the second
constexpr int PNS = PLA + 1;                           // S accumulators in TMEM
constexpr int P_SMEM_BYTES = PS1 * PH1_BYTES + PS2 * PH2_BYTES + 2 * QTILE_BYTES + PXB * XHALF_BYTES + 1024 + 512;
static_assert(P_SMEM_BYTES <= 232448, "shared memory budget exceeded");
// TMEM columns: G_r [0,128) | S buffers [128,160) [160,192) | W [192,448) | S buffers [448,480) [480,512)
constexpr int P_G_COL = 0, P_W_COL = 192, P_TMEM_COLS = 512;
__device__ __forceinline__ uint32_t s_col(uint32_t a) { return a < 2 ? 128u + 32u * a : 384u + 32u * a; }

struct Fused256Params {
  int64_t M, F;
  int n_blocks, n_steps, n_kb;      // n_kb: 32-wide blocks of components that exist (ldw / 32)
  const float *W;
  int64_t ldw;
  float *Wout;
  int64_t ldwo, w_cols;
  double *kl;
  const int *stop;
  int *err;
  float qshift;   // the ratio tile holds q - qshift (centered ratio, api.cu)
  const float *colbias;   // W' = W (.) (G + colbias[component])
  const float *Wlo;       // TF32R: low parts of W (W itself is the TF32-exact high part); nullptr otherwise
  float *Wout_lo;         // TF32R: low parts of W'
  int accurate;           // TF32R: cancellation-free objective, centered ratio stored as u, tile rounded to TF32
  int write_q;            // fit: every CTA also stores its half of the ratio tile (TMA) for the numerator N += W'^T.Q
  int dbg;        // timing experiments only (KLNMF_F256_DBG): 1 no exchange, 2 no S MMAs, 4 no G MMAs, 8 no ratio math
};

// bulk copy from our shared memory into the peer's; the bytes complete on a barrier of the PEER (both shared::cluster addresses)
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16i(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16w(uint32_t v[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld16w2(uint32_t v[16], uint32_t w[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(w[0]), "+r"(w[1]),
                 "+r"(w[2]), "+r"(w[3]), "+r"(w[4]), "+r"(w[5]), "+r"(w[6]), "+r"(w[7]), "+r"(w[8]), "+r"(w[9]), "+r"(w[10]),
                 "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st16p(uint32_t taddr, const uint32_t v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void umma_tf32_tsp(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
fused_coef256_kernel(const __grid_constant__ CUtensorMap tmHt, const __grid_constant__ CUtensorMap tmH,
                     const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ,
                     const Fused256Params p) {
  if (p.stop != nullptr && *p.stop != 0) return;     // uniform over the grid: both CTAs of a cluster leave together
  const uint32_t crank = cluster_ctarank();
  const uint32_t peer = crank ^ 1u;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t h1_s = smem_base;
  const uint32_t h2_s = h1_s + PS1 * PH1_BYTES;
  const uint32_t q_s = h2_s + PS2 * PH2_BYTES;
  const uint32_t x_s = q_s + 2 * QTILE_BYTES;
  const uint32_t bar_base = x_s + PXB * XHALF_BYTES;
  uint8_t *q_gen = smem_gen + (q_s - smem_base);
  uint8_t *x_gen = smem_gen + (x_s - smem_base);
  // barriers: w_full | w_empty | g_full | g_empty | h1_full[S1] | h1_empty[S1] | h2_full[S2] | h2_empty[S2] |
  //           s_full[NS] | s_empty[NS] | q_full[2] | q_empty[2] | x_full[XB] | x_empty[XB] | tmem_ptr
  const uint32_t w_full = bar_base, w_empty = bar_base + 8, g_full = bar_base + 16, g_empty = bar_base + 24;
  auto h1_full = [&](int s) { return bar_base + 32u + 8u * s; };
  auto h1_empty = [&](int s) { return bar_base + 32u + 8u * (PS1 + s); };
  const uint32_t b1 = bar_base + 32u + 16u * PS1;
  auto h2_full = [&](int s) { return b1 + 8u * s; };
  auto h2_empty = [&](int s) { return b1 + 8u * (PS2 + s); };
  const uint32_t b2 = b1 + 16u * PS2;
  auto s_full = [&](int a) { return b2 + 8u * a; };
  auto s_empty = [&](int a) { return b2 + 8u * (PNS + a); };
  const uint32_t b3 = b2 + 16u * PNS;
  auto q_full = [&](int a) { return b3 + 8u * a; };
  auto q_empty = [&](int a) { return b3 + 16u + 8u * a; };
  auto x_full = [&](int b) { return b3 + 32u + 8u * b; };
  auto x_empty = [&](int b) { return b3 + 32u + 8u * (PXB + b); };
  const uint32_t tmem_ptr_addr = b3 + 32u + 16u * PXB;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_ptr_addr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmHt); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmX);
    if (p.write_q) tma_prefetch_desc(&tmQ);
    mbar_init(w_full, P_EPI_WARPS); mbar_init(w_empty, 1); mbar_init(g_full, 1); mbar_init(g_empty, P_EPI_WARPS);
    for (int s = 0; s < PS1; s++) { mbar_init(h1_full(s), 1); mbar_init(h1_empty(s), 1); }
    for (int s = 0; s < PS2; s++) { mbar_init(h2_full(s), 1); mbar_init(h2_empty(s), 1); }
    for (int a = 0; a < PNS; a++) { mbar_init(s_full(a), 1); mbar_init(s_empty(a), P_EPI_WARPS / 2); }
    for (int a = 0; a < 2; a++) {
      mbar_init(q_full(a), P_EPI_WARPS / 2);  // the four local warps of the step's group (+ 16 KB of peer bulk copies)
      mbar_init(q_empty(a), 2);               // the tcgen05.commit of both issuers
    }
    for (int b = 0; b < PXB; b++) { mbar_init(x_full(b), 1); mbar_init(x_empty(b), P_EPI_WARPS / 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_addr, P_TMEM_COLS);
  tc_fence_before();
  cluster_sync();                              // barrier inits visible to the peer before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t g_tmem = tmem_base + P_G_COL, w_tmem = tmem_base + P_W_COL;

  const int nsteps = p.n_steps;
  const int nkb = p.n_kb;
  const int cl_first = blockIdx.x >> 1, cl_step = gridDim.x >> 1;    // row blocks are dealt to clusters

  if (warp == 0) {
    // =============================== TMA producer 1: the H^T tiles (B of S = W.H) of this CTA ===============================
    if (lane == 0) {
      uint32_t hc = 0;
      for (int rb = cl_first; rb < p.n_blocks; rb += cl_step)
        for (int j = 0; j < nsteps; j++, hc++) {
          const uint32_t s = hc % PS1, ph = (hc / PS1) & 1u;
          mbar_wait(h1_empty(s), ph ^ 1u, p.err, 2);
          mbar_expect_tx(h1_full(s), (uint32_t)nkb * (PHN * 128));
          // one 3D box: 32 components x 32 feature rows x all K blocks (one instruction instead of eight)
          tma_load_3d(h1_s + s * PH1_BYTES, &tmHt, h1_full(s), 0, j * PBN + (int)crank * PHN, 0);
        }
    }
  } else if (warp == 11) {
    // =============================== TMA producer 2: the H tiles (B of G += Q.H^T) of this CTA ===============================
    if (lane == 0) {
      uint32_t hc = 0;
      for (int rb = cl_first; rb < p.n_blocks; rb += cl_step)
        for (int j = 0; j < nsteps; j++, hc++) {
          const uint32_t s = hc % PS2, ph = (hc / PS2) & 1u;
          mbar_wait(h2_empty(s), ph ^ 1u, p.err, 14);
          mbar_expect_tx(h2_full(s), PH2_BYTES);
          const uint32_t dst = h2_s + s * PH2_BYTES;
          tma_load_2d(dst, &tmH, h2_full(s), j * PBN, (int)crank * PKH);
          tma_load_2d(dst + PKH * 128, &tmH, h2_full(s), j * PBN + 32, (int)crank * PKH);
        }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PHN >> 3) << 17) | ((uint32_t)(PBM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PKH >> 3) << 17) | ((uint32_t)(PBM >> 4) << 24);
    const uint32_t hi = desc_hi(1024u, 2u);
    const uint32_t h1_lo = desc_lo(h1_s, 16u), h2_lo = desc_lo(h2_s, 16u), q_lo = desc_lo(q_s, 16u);
    uint32_t c1 = 0, c2 = 0, rbc = 0;
    for (int rb = cl_first; rb < p.n_blocks; rb += cl_step, rbc++) {
      mbar_wait(w_full, rbc & 1u, p.err, 3);
      tc_fence_after();
      for (int j = 0; j < nsteps + PLA; j++) {
        if (j < nsteps) {
          // ---- S_r[a] = W . H^T-tile(j)[its 32 features] ----
          const uint32_t s = c1 % PS1, a = c1 % PNS;
          mbar_wait(h1_full(s), (c1 / PS1) & 1u, p.err, 4);
          mbar_wait(s_empty(a), ((c1 / PNS) & 1u) ^ 1u, p.err, 5);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bl = h1_lo + ((s * PH1_BYTES) >> 4);
#pragma unroll 2
            for (int kb = 0; kb < ((p.dbg & 2) ? 1 : nkb); kb++)
#pragma unroll
              for (int kk = 0; kk < 4; kk++)
                umma_tf32_tsp(tmem_base + s_col(a), w_tmem + kb * 32 + kk * 8,
                              desc_pack(bl + ((kb * (PHN * 128) + kk * 32) >> 4), hi), idesc1, (kb | kk) ? 1u : 0u);
            umma_commit<1>(s_full(a));
            umma_commit<1>(h1_empty(s));
          }
          __syncwarp();
          c1++;
        }
        if (j >= PLA) {
          // ---- G_r += Q(j-PLA) . H-tile(j-PLA)[its 128 components]^T ----
          const uint32_t s = c2 % PS2, b = c2 & 1u;
          mbar_wait(h2_full(s), (c2 / PS2) & 1u, p.err, 15);
          mbar_wait(q_full(b), (c2 >> 1) & 1u, p.err, 6);
          if (j == PLA) mbar_wait(g_empty, (rbc & 1u) ^ 1u, p.err, 7);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t al = q_lo + ((b * QTILE_BYTES) >> 4);
            const uint32_t bl = h2_lo + ((s * PH2_BYTES) >> 4);
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
              for (int kk = 0; kk < 4; kk++)
                if (!(p.dbg & 4) || (h | kk) == 0)
                  umma_tf32<1>(g_tmem, desc_pack(al + ((h * QHALF_BYTES + kk * 32) >> 4), hi),
                               desc_pack(bl + ((h * (PKH * 128) + kk * 32) >> 4), hi), idesc2, (j > PLA || (h | kk)) ? 1u : 0u);
            umma_commit<1>(h2_empty(s));
            umma_commit<1>(q_empty(b));                    // our Q tile may be rewritten ...
            umma_commit<1>(mapa(q_empty(b), peer));        // ... and the peer, who writes half of it, hears it too
            if (j == nsteps + PLA - 1) { umma_commit<1>(g_full); umma_commit<1>(w_empty); }
          }
          __syncwarp();
          c2++;
        }
      }
    }
  } else if (warp == 10) {
    // =============================== X loader: this CTA's 32 columns of every step ===============================
    // (pulling the chunks into L2 a few steps ahead with cp.async.bulk.prefetch.tensor changed nothing: 4.87 vs 4.92 ms)
    if (lane == 0) {
      uint32_t g = 0;
      for (int rb = cl_first; rb < p.n_blocks; rb += cl_step)
        for (int j = 0; j < nsteps; j++, g++) {
          const uint32_t b = g % PXB, ph = (g / PXB) & 1u;
          mbar_wait(x_empty(b), ph ^ 1u, p.err, 8);
          mbar_expect_tx(x_full(b), XHALF_BYTES);
          tma_load_2d(x_s + b * XHALF_BYTES, &tmX, x_full(b), j * PBN + (int)crank * PHN, rb * PBM);
        }
    }
  } else {
    // =============================== epilogue ===============================
    const int e = warp - 2;
    const int quarter = warp & 3;
    const int grp = e >> 2;                 // the two groups take alternate steps
    const int r = quarter * 32 + lane;
    const uint32_t sw = (uint32_t)(r & 7);  // SWIZZLE_128B: 16-byte chunk index ^ (row & 7)
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t q_half = q_s + grp * QTILE_BYTES + crank * QHALF_BYTES + quarter * (32 * 128);   // this warp's 32 rows of our half tile
    const uint32_t q_half_peer = mapa(q_half, peer);
    const uint32_t q_full_peer = mapa(q_full(grp), peer);
    // ---- the whole W block of a row block goes to TMEM (A operand of the first contraction): lane = row, column =
    //      component.  A warp fills the 64 columns it reads back itself in the final epilogue plus the matching 64 of
    //      the peer's half, so no other warp's pending TMEM reads are overwritten; 32 columns per round trip ----
    auto fill_col = [&](int cc) { return (int)((cc < 2 ? crank : peer) * PKH) + grp * (PKH / 2) + (cc & 1) * 32; };
    auto fill_w = [&](int rb_fill, uint32_t parity) {
      const int64_t frow = (int64_t)rb_fill * PBM + r;
      mbar_wait(w_empty, parity ^ 1u, p.err, 13);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < PKP / 64; cc++) {
        const int col0 = fill_col(cc);
        if (col0 >= nkb * 32) continue;                  // component blocks the contraction never reads
        uint32_t wv[32];
        if (frow < p.M) {
          const float *wi = p.W + frow * p.ldw + col0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(wi + 4 * i));
            wv[4 * i] = __float_as_uint(w.x); wv[4 * i + 1] = __float_as_uint(w.y);
            wv[4 * i + 2] = __float_as_uint(w.z); wv[4 * i + 3] = __float_as_uint(w.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i++) wv[i] = 0u;
        }
        tmem_st16p(w_tmem + lane_addr + col0, wv);
        tmem_st16p(w_tmem + lane_addr + col0 + 16, wv + 16);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_full);
    };
    // the next row block's coefficients are pulled into L2 a few steps before they are needed
    auto prefetch_w = [&](int rb_pf) {
      const int64_t prow = (int64_t)rb_pf * PBM + r;
      if (rb_pf < p.n_blocks && prow < p.M) {
#pragma unroll
        for (int cc = 0; cc < 4; cc++)
          if (fill_col(cc) < nkb * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + prow * p.ldw + fill_col(cc)));
      }
    };
    // TF32R: the low parts of this row block's coefficients are read in its final epilogue only -- pull them into L2
    auto prefetch_wlo = [&](int rb_cur) {
      const int64_t prow = (int64_t)rb_cur * PBM + r;
      if (p.Wlo != nullptr && prow < p.M) {
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
          const int col = (int)crank * PKH + grp * (PKH / 2) + cc * 32;
          if (col < p.w_cols) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.Wlo + prow * p.ldw + col));
        }
      }
    };
    const int j_pf = nsteps > 8 ? nsteps - 8 : 0;
    if (cl_first < p.n_blocks) fill_w(cl_first, 0u);
    uint32_t c = 0, rbc = 0;
    double kl = 0.0;
    for (int rb = cl_first; rb < p.n_blocks; rb += cl_step, rbc++) {
      float kl_blk = 0.f;
      const int64_t row = (int64_t)rb * PBM + r;
      const int rb_next = rb + cl_step;
#pragma unroll 1
      for (int j = 0; j < nsteps; j++, c++) {
        if (j == j_pf) { prefetch_w(rb_next); prefetch_wlo(rb); }
        if ((int)(c & 1u) != grp) continue;
        const uint32_t ph2 = (c >> 1) & 1u;         // PXB == 2: X buffer grp, same phase as the Q buffers
        const uint32_t sa = c % PNS;
        mbar_wait(s_full(sa), (c / PNS) & 1u, p.err, 9);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32_issue(tmem_base + s_col(sa) + lane_addr, v);
        mbar_wait(x_full(grp), ph2, p.err, 10);
        float x[32];
        const uint8_t *xrow = x_gen + grp * XHALF_BYTES + r * 128;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const float4 t = *reinterpret_cast<const float4 *>(xrow + ((((uint32_t)i) ^ sw) << 4));
          x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
        }
        tmem_ld32_wait(v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive_relaxed(s_empty(sa)); mbar_arrive(x_empty(grp)); }
        if (!(p.dbg & 8)) kl_blk += ratio_chunk32_dispatch(x, v, p.qshift, p.accurate);
        // both issuers are done with Q tile `grp` (ours and the peer's, of which we write half each)
        mbar_wait(q_empty(grp), ph2 ^ 1u, p.err, 11);
        if (p.write_q) {      // ... and so is the TMA store of this warp's box two steps ago
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
        }
        uint8_t *qrow = q_gen + grp * QTILE_BYTES + crank * QHALF_BYTES + r * 128;
#pragma unroll
        for (int i = 0; i < 8; i++)
          *reinterpret_cast<float4 *>(qrow + ((((uint32_t)i) ^ sw) << 4)) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        fence_proxy_async();                // generic-proxy stores before the async-proxy reads (our MMAs, the bulk copy)
        __syncwarp();
        if (lane == 0) {
          // fit: this warp's 32 rows x 32 features of the ratio tile also leave for HBM (the numerator N += W'^T.Q,
          // nmf.py:349, reads the panel once W' of all its rows exists); rows >= M and features >= F are clipped
          if (p.write_q) tma_store_2d(&tmQ, q_half, j * PBN + (int)crank * PHN, rb * PBM + quarter * 32);
          if (!(p.dbg & 1)) bulk_copy_to_peer(q_half_peer, q_half, 32 * 128, q_full_peer);
          if (quarter == 0 && !(p.dbg & 1)) mbar_expect_tx(q_full(grp), QHALF_BYTES);   // arrives, and expects the peer's four copies
          else mbar_arrive(q_full(grp));
        }
      }
      kl += (double)kl_blk;
      // ---- final epilogue: W'[:, 128 r + ...] = W (.) G_r, each group one half of this CTA's 128 columns; W is read
      //      back from its TMEM block (it left L2 long ago) ----
      mbar_wait(g_full, rbc & 1u, p.err, 12);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < PKH / 32; cc++) {
        const int lc = grp * (PKH / 2) + cc * 16;            // column inside G_r
        const int col0 = (int)crank * PKH + lc;              // component
        uint32_t v[16], w[16];
        tmem_ld16i(g_tmem + lane_addr + lc, v);
        tmem_ld16i(w_tmem + lane_addr + col0, w);
        tmem_ld16w2(v, w);
        if (row < p.M && col0 < p.w_cols) {
          float *wo = p.Wout + row * p.ldwo + col0;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            float4 wv = make_float4(__uint_as_float(w[4 * i]), __uint_as_float(w[4 * i + 1]), __uint_as_float(w[4 * i + 2]),
                                    __uint_as_float(w[4 * i + 3]));
            if (p.Wlo) {      // the TMEM block holds the TF32-exact high parts; the update multiplies the full FP32 state
              const float4 l = __ldg(reinterpret_cast<const float4 *>(p.Wlo + row * p.ldw + col0 + 4 * i));
              wv.x += l.x; wv.y += l.y; wv.z += l.z; wv.w += l.w;
            }
            float4 g = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
            if (p.colbias) {  // centered ratio: G = (Q-1).H^T + rowsum(H), clamped at zero like the unfused epilogue
              const float4 b = __ldg(reinterpret_cast<const float4 *>(p.colbias + col0 + 4 * i));
              g.x = fmaxf(g.x + b.x, 0.f); g.y = fmaxf(g.y + b.y, 0.f); g.z = fmaxf(g.z + b.z, 0.f); g.w = fmaxf(g.w + b.w, 0.f);
            }
            float4 o = make_float4(wv.x * g.x, wv.y * g.y, wv.z * g.z, wv.w * g.w);
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
      if (rb_next < p.n_blocks) fill_w(rb_next, (rbc + 1u) & 1u);
    }
    if (p.write_q && lane == 0) bulk_wait_all();
    if (p.kl != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
      if (lane == 0) atomicAdd(p.kl, kl);
    }
  }

  tc_fence_before();
  cluster_sync();                              // nobody leaves while the peer may still store into its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, P_TMEM_COLS);
  }
}

// 2D fp32 tensor map with a free inner box extent and swizzle mode
int make_map_ex(CUtensorMap *map, const void *base, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_rows,
                CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode();
  KL_CHECK(enc != nullptr, KLNMF_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  KL_CHECK(((uintptr_t)base % 16) == 0 && (ld * 4) % 16 == 0, KLNMF_EINVAL,
           "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%lld)", (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KL_CHECK(r == CUDA_SUCCESS, KLNMF_ECUDA, "cuTensorMapEncodeTiled failed with %d (inner=%lld outer=%lld ld=%lld)", (int)r,
           (long long)inner, (long long)outer, (long long)ld);
  return KLNMF_OK;
}

}  // namespace

int fused_coef_step256(klnmf_ctx *ctx, const FusedDesc &d, int *err_dev) {
  KL_CHECK(d.K <= PKP && d.ldw % 32 == 0 && d.ldht % 32 == 0, KLNMF_EINVAL,
           "fused_coef_step256: k=%lld not supported", (long long)d.K);
  CUtensorMap tmHt, tmH, tmX, tmQ;
  {
    // H^T (f x ldht) as a 3D tensor: 32 components (inner) x f feature rows x ldht/32 component blocks 128 B apart
    EncodeTiledFn enc = get_encode();
    KL_CHECK(enc != nullptr, KLNMF_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {32u, (cuuint64_t)d.F, (cuuint64_t)(d.ldw / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)d.ldht * 4, 128u};
    cuuint32_t box[3] = {32u, (cuuint32_t)PHN, (cuuint32_t)(d.ldw / 32)};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(&tmHt, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void *>(d.Ht), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KL_CHECK(r == CUDA_SUCCESS, KLNMF_ECUDA, "cuTensorMapEncodeTiled (3D H^T) failed with %d", (int)r);
  }
  KL_TRY(make_map_ex(&tmH, d.H, d.F, d.K, d.ldh, 32, PKH, CU_TENSOR_MAP_SWIZZLE_128B));          // 32 features x 128 rows
  KL_TRY(make_map_ex(&tmX, d.X, d.F, d.M, d.ldx, PHN, PBM, CU_TENSOR_MAP_SWIZZLE_128B));         // 32 features x 128 rows
  tmQ = tmX;
  if (d.Q != nullptr) {      // fit: one 32-feature x 32-row box per warp and step, clipped at the edges of the panel
    KL_CHECK(d.ldq >= round_up(d.F, 4), KLNMF_EINVAL, "fused_coef_step256: ratio panel leading dimension too small");
    KL_TRY(make_map_ex(&tmQ, d.Q, d.F, d.M, d.ldq, PHN, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  }
  Fused256Params p{};
  p.M = d.M; p.F = d.F;
  p.Wlo = (const float *)d.Wlo; p.Wout_lo = (float *)d.Wout_lo; p.accurate = d.accurate;
  p.write_q = d.Q != nullptr ? 1 : 0;
  p.n_blocks = (int)ceil_div(d.M, PBM);
  p.n_steps = (int)ceil_div(d.F, PBN);
  p.n_kb = (int)(d.ldw / 32);
  p.W = (const float *)d.W; p.ldw = d.ldw;
  p.Wout = (float *)d.Wout; p.ldwo = d.ldwo;
  p.w_cols = d.ldw < d.ldwo ? d.ldw : d.ldwo;
  p.kl = d.kl; p.stop = d.stop; p.err = err_dev;
  p.qshift = d.qshift; p.colbias = d.colbias;
  p.dbg = getenv("KLNMF_F256_DBG") ? atoi(getenv("KLNMF_F256_DBG")) : 0;
  if (p.n_blocks == 0 || p.n_steps == 0) return KLNMF_OK;
  const int clusters = p.n_blocks < ctx->sm_count / 2 ? p.n_blocks : ctx->sm_count / 2;
  // per DEVICE: function attributes live in the device's context, and one process may drive several GPUs
  static bool attr_done[64] = {};
  if (!attr_done[ctx->device & 63]) {
    KL_CUDA(cudaFuncSetAttribute(fused_coef256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    attr_done[ctx->device & 63] = true;
  }
  fused_coef256_kernel<<<clusters * 2, P_THREADS, P_SMEM_BYTES, ctx->stream>>>(tmHt, tmH, tmX, tmQ, p);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

}  // namespace klnmf
