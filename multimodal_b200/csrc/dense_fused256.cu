// Fused coefficient half-step for 128 < k <= 256 on a CLUSTER OF TWO CTAs that share one block of 128 samples.
//
// At k = 256 the coefficient accumulator G (128 lanes x 256 FP32 columns) and the resident W block (256 columns
// as the TMEM A operand) fill a whole TMEM by themselves, so the single-CTA kernel of dense_fused.cu has no
// column left for S.  Here the two SMs of a cluster split the COLUMNS of both contractions over their TMEMs:
//
//   CTA r:  S_r = W . H[:, 16 features of the step]      (first contraction, N = 16: its half of the 32-feature step)
//           Q_r = (X_r+eps)/(S_r+eps), objective          (TMEM epilogue; X_r = its 16 columns of X)
//           Q_r is written into the Q tile of BOTH CTAs  (local st.shared, then one 2 KB cp.async.bulk per warp
//                                                         from its own shared memory into the peer's)
//           G_r += Q . H[128 r : 128 r + 128, step]^T     (second contraction, N = 128: its half of the components)
//           W'[:, 128 r : 128 r + 128] = W (.) G_r        (final epilogue)
//
// No MMA work is duplicated, Q crosses the cluster's distributed shared memory (8 KB per step and direction) and
// never HBM; each CTA holds the full W block in TMEM (A operand of the first contraction, TS form).  The protocol
// is the one of dense_fused.cu with two changes: q_full collects the four local epilogue warps plus the 8 KB the
// peer's bulk copies complete on it (complete_tx, like a TMA load), and q_empty collects the tcgen05.commit of both
// issuers.  The Q tile is two K-major half tiles of 128 rows x 16 features (64-byte rows, SWIZZLE_64B), one per
// producing CTA, so that the 32 rows of a warp are one contiguous 2 KB range.  (The first version stored into the
// peer with st.shared::cluster and signalled with mbarrier.arrive.release.cluster: the all-space fence.proxy.async
// plus the cluster-scope release cost ~2300 clk per step and warp group -- MEMBAR.ALL.GPU, ERRBAR, FENCE.VIEW.ASYNC
// were 31 % of all stall samples, profiles/r1_fused256_cfg3_n262144_summary.csv -- and the kernel ran at half the
// speed of the unfused form.)
// Reference lines as in dense_fused.cu (nmf.py:325-343, metrics.py:18-20).  Transform only (fit keeps the
// three-contraction form at k > 128).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace klnmf {

namespace {

constexpr int PBM = 128;                   // samples per row block
constexpr int PBN = 32;                    // features per step (16 per CTA)
constexpr int PHN = 16;                    // features per step and CTA
constexpr int PKP = 256;                   // padded components
constexpr int PKH = 128;                   // components per CTA (columns of G)
constexpr int P_THREADS = 352;
constexpr int P_EPI_WARPS = 8;
constexpr int QTILE_BYTES = PBM * PBN * 4;             // 16 KB: the 128 x 32 ratio tile = two half tiles of 128 x 16 (K-major, 64B swizzle)
constexpr int QHALF_BYTES = PBM * PHN * 4;             // 8 KB: the half one CTA produces
constexpr int XHALF_BYTES = PBM * PHN * 4;             // 8 KB: 128 rows x 16 features, 64-byte rows (64B swizzle)
constexpr int PH1_BYTES = (PKP / 32) * PHN * 32 * 4;   // 16 KB: H^T tile, 8 K blocks of 16 feature rows x 32 k
constexpr int PH2_BYTES = PKH * PBN * 4;               // 16 KB: H tile, 128 component rows x 32 features
constexpr int PSTAGE_BYTES = PH1_BYTES + PH2_BYTES;
constexpr int PHS = 5;                                 // dictionary stages
constexpr int PXB = 4;                                 // X half-chunks in flight
constexpr int PLA = 2;                                 // steps the first contraction runs ahead of the second
constexpr int P_SMEM_BYTES = PHS * PSTAGE_BYTES + 2 * QTILE_BYTES + PXB * XHALF_BYTES + 1024 + 256;
static_assert(P_SMEM_BYTES <= 232448, "shared memory budget exceeded");
// TMEM columns: G_r [0,128) | S buffers [128,144) [160,176) | W [192,448)
constexpr int P_G_COL = 0, P_S_COL = 128, P_S_STRIDE = 32, P_W_COL = 192, P_TMEM_COLS = 512;

struct Fused256Params {
  int64_t M, F;
  int n_blocks, n_steps;
  const float *W;
  int64_t ldw;
  float *Wout;
  int64_t ldwo, w_cols;
  double *kl;
  const int *stop;
  int *err;
};

// bulk copy from our shared memory into the peer's; the bytes complete on a barrier of the PEER (both shared::cluster addresses)
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(bar_cluster)
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
                     const __grid_constant__ CUtensorMap tmX, const Fused256Params p) {
  if (p.stop != nullptr && *p.stop != 0) return;     // uniform over the grid: both CTAs of a cluster leave together
  const uint32_t crank = cluster_ctarank();
  const uint32_t peer = crank ^ 1u;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t h_s = smem_base;
  const uint32_t q_s = h_s + PHS * PSTAGE_BYTES;
  const uint32_t x_s = q_s + 2 * QTILE_BYTES;
  const uint32_t bar_base = x_s + PXB * XHALF_BYTES;
  uint8_t *q_gen = smem_gen + (q_s - smem_base);
  uint8_t *x_gen = smem_gen + (x_s - smem_base);
  // barriers: w_full | w_empty | g_full | g_empty | h_full[HS] | h_empty[HS] | s_full[2] | s_empty[2] |
  //           q_full[2] | q_empty[2] | x_full[XB] | x_empty[XB] | tmem_ptr
  const uint32_t w_full = bar_base, w_empty = bar_base + 8, g_full = bar_base + 16, g_empty = bar_base + 24;
  auto h_full = [&](int s) { return bar_base + 32u + 8u * s; };
  auto h_empty = [&](int s) { return bar_base + 32u + 8u * (PHS + s); };
  const uint32_t b2 = bar_base + 32u + 16u * PHS;
  auto s_full = [&](int a) { return b2 + 8u * a; };
  auto s_empty = [&](int a) { return b2 + 16u + 8u * a; };
  auto q_full = [&](int a) { return b2 + 32u + 8u * a; };
  auto q_empty = [&](int a) { return b2 + 48u + 8u * a; };
  auto x_full = [&](int b) { return b2 + 64u + 8u * b; };
  auto x_empty = [&](int b) { return b2 + 64u + 8u * (PXB + b); };
  const uint32_t tmem_ptr_addr = b2 + 64u + 16u * PXB;
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_ptr_addr - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmHt); tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmX);
    mbar_init(w_full, P_EPI_WARPS); mbar_init(w_empty, 1); mbar_init(g_full, 1); mbar_init(g_empty, P_EPI_WARPS);
    for (int s = 0; s < PHS; s++) { mbar_init(h_full(s), 1); mbar_init(h_empty(s), 1); }
    for (int a = 0; a < 2; a++) {
      mbar_init(s_full(a), 1); mbar_init(s_empty(a), P_EPI_WARPS / 2);
      mbar_init(q_full(a), P_EPI_WARPS / 2);  // the four local warps of the step's group (+ 8 KB of peer bulk copies)
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
  const uint32_t g_tmem = tmem_base + P_G_COL, s_tmem = tmem_base + P_S_COL, w_tmem = tmem_base + P_W_COL;

  const int nsteps = p.n_steps;
  const int cl_first = blockIdx.x >> 1, cl_step = gridDim.x >> 1;    // row blocks are dealt to clusters

  if (warp == 0) {
    // =============================== TMA producer: the two dictionary tiles of this CTA ===============================
    if (lane == 0) {
      uint32_t hc = 0;
      for (int rb = cl_first; rb < p.n_blocks; rb += cl_step)
        for (int j = 0; j < nsteps; j++, hc++) {
          const uint32_t s = hc % PHS, ph = (hc / PHS) & 1u;
          mbar_wait(h_empty(s), ph ^ 1u, p.err, 2);
          mbar_expect_tx(h_full(s), PSTAGE_BYTES);
          const uint32_t dst = h_s + s * PSTAGE_BYTES;
          for (int kb = 0; kb < PKP / 32; kb++)
            tma_load_2d(dst + kb * (PHN * 128), &tmHt, h_full(s), kb * 32, j * PBN + (int)crank * PHN);
          tma_load_2d(dst + PH1_BYTES, &tmH, h_full(s), j * PBN, (int)crank * PKH);
        }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PHN >> 3) << 17) | ((uint32_t)(PBM >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(PKH >> 3) << 17) | ((uint32_t)(PBM >> 4) << 24);
    const uint32_t hi = desc_hi(1024u, 2u);
    const uint32_t hi64 = desc_hi(512u, 4u);             // Q half tiles: SWIZZLE_64B, 8-row groups 512 B apart
    const uint32_t h_lo = desc_lo(h_s, 16u), q_lo = desc_lo(q_s, 16u);
    uint32_t c1 = 0, c2 = 0, rbc = 0;
    for (int rb = cl_first; rb < p.n_blocks; rb += cl_step, rbc++) {
      mbar_wait(w_full, rbc & 1u, p.err, 3);
      tc_fence_after();
      for (int j = 0; j < nsteps + PLA; j++) {
        if (j < nsteps) {
          // ---- S_r[a] = W . H^T-tile(j)[its 16 features] ----
          const uint32_t s = c1 % PHS, a = c1 & 1u;
          mbar_wait(h_full(s), (c1 / PHS) & 1u, p.err, 4);
          mbar_wait(s_empty(a), ((c1 >> 1) & 1u) ^ 1u, p.err, 5);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t bl = h_lo + ((s * PSTAGE_BYTES) >> 4);
#pragma unroll
            for (int kb = 0; kb < PKP / 32; kb++)
#pragma unroll
              for (int kk = 0; kk < 4; kk++)
                umma_tf32_tsp(s_tmem + a * P_S_STRIDE, w_tmem + kb * 32 + kk * 8,
                              desc_pack(bl + ((kb * (PHN * 128) + kk * 32) >> 4), hi), idesc1, (kb | kk) ? 1u : 0u);
            umma_commit<1>(s_full(a));
          }
          __syncwarp();
          c1++;
        }
        if (j >= PLA) {
          // ---- G_r += Q(j-PLA) . H-tile(j-PLA)[its 128 components]^T ----
          const uint32_t s = c2 % PHS, b = c2 & 1u;
          mbar_wait(q_full(b), (c2 >> 1) & 1u, p.err, 6);
          if (j == PLA) mbar_wait(g_empty, (rbc & 1u) ^ 1u, p.err, 7);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t al = q_lo + ((b * QTILE_BYTES) >> 4);
            const uint32_t bl = h_lo + ((s * PSTAGE_BYTES + PH1_BYTES) >> 4);
#pragma unroll
            for (int kk = 0; kk < 4; kk++)
              umma_tf32<1>(g_tmem, desc_pack(al + (((kk >> 1) * QHALF_BYTES + (kk & 1) * 32) >> 4), hi64),
                           desc_pack(bl + ((kk * 32) >> 4), hi), idesc2, (j > PLA || kk > 0) ? 1u : 0u);
            umma_commit<1>(h_empty(s));
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
    // =============================== X loader: this CTA's 16 columns of every step ===============================
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
    const uint32_t sw = (uint32_t)((r >> 1) & 3);        // SWIZZLE_64B: 16-byte chunk index ^ address bits [7,8]
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t q_half = q_s + crank * QHALF_BYTES + quarter * (32 * PHN * 4);   // this warp's 32 rows of our half tile
    const uint32_t q_half_peer = mapa(q_half, peer);
    uint32_t c = 0, rbc = 0;
    double kl = 0.0;
    for (int rb = cl_first; rb < p.n_blocks; rb += cl_step, rbc++) {
      float kl_blk = 0.f;
      const int64_t row = (int64_t)rb * PBM + r;
      {
        // ---- the whole W block of this row block goes to TMEM: lane = row, column = component ----
        mbar_wait(w_empty, (rbc & 1u) ^ 1u, p.err, 13);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < PKP / 32; cc++) {
          const int col0 = grp * (PKP / 2) + cc * 16;
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
          tmem_st16p(w_tmem + lane_addr + col0, wv);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(w_full);
      }
#pragma unroll 1
      for (int j = 0; j < nsteps; j++, c++) {
        if ((int)(c & 1u) != grp) continue;
        const uint32_t ph2 = (c >> 1) & 1u, xb = c % PXB, phx = (c / PXB) & 1u;
        mbar_wait(s_full(grp), ph2, p.err, 9);
        tc_fence_after();
        uint32_t v[16];
        tmem_ld16i(s_tmem + lane_addr + grp * P_S_STRIDE, v);
        mbar_wait(x_full(xb), phx, p.err, 10);
        float x[16];
        const uint8_t *xrow = x_gen + xb * XHALF_BYTES + r * 64;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float4 t = *reinterpret_cast<const float4 *>(xrow + ((((uint32_t)i) ^ sw) << 4));
          x[4 * i] = t.x; x[4 * i + 1] = t.y; x[4 * i + 2] = t.z; x[4 * i + 3] = t.w;
        }
        tmem_ld16w(v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive_relaxed(s_empty(grp)); mbar_arrive(x_empty(xb)); }
        float part0 = 0.f, part1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          float q0, q1;
          part0 += ratio_term<false>(x[i], __uint_as_float(v[i]), q0);
          part1 += ratio_term<false>(x[i + 1], __uint_as_float(v[i + 1]), q1);
          x[i] = q0; x[i + 1] = q1;
        }
        kl_blk += part0 + part1;
        // both issuers are done with Q tile `grp` (ours and the peer's, of which we write half each)
        mbar_wait(q_empty(grp), ph2 ^ 1u, p.err, 11);
        const uint32_t off = (uint32_t)(grp * QTILE_BYTES + crank * QHALF_BYTES + r * 64);
#pragma unroll
        for (int i = 0; i < 4; i++)
          *reinterpret_cast<float4 *>(q_gen + off + ((((uint32_t)i) ^ sw) << 4)) =
              make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        fence_proxy_async();                // generic-proxy stores before the async-proxy reads (our MMAs, the bulk copy)
        __syncwarp();
        if (lane == 0) {
          bulk_copy_to_peer(q_half_peer + grp * QTILE_BYTES, q_half + grp * QTILE_BYTES, 32 * PHN * 4, mapa(q_full(grp), peer));
          if (quarter == 0) mbar_expect_tx(q_full(grp), QHALF_BYTES);   // arrives, and expects the peer's four copies
          else mbar_arrive(q_full(grp));
        }
      }
      kl += (double)kl_blk;
      // ---- final epilogue: W'[:, 128 r + ...] = W (.) G_r, each group one half of this CTA's 128 columns ----
      mbar_wait(g_full, rbc & 1u, p.err, 12);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < PKH / 32; cc++) {
        const int lc = grp * (PKH / 2) + cc * 16;            // column inside G_r
        const int col0 = (int)crank * PKH + lc;              // component
        uint32_t v[16];
        tmem_ld16i(g_tmem + lane_addr + lc, v);
        tmem_ld16w(v);
        if (row < p.M && col0 < p.w_cols) {
          const float *wi = p.W + row * p.ldw + col0;
          float *wo = p.Wout + row * p.ldwo + col0;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(wi + 4 * i));
            *reinterpret_cast<float4 *>(wo + 4 * i) =
                make_float4(w.x * __uint_as_float(v[4 * i]), w.y * __uint_as_float(v[4 * i + 1]),
                            w.z * __uint_as_float(v[4 * i + 2]), w.w * __uint_as_float(v[4 * i + 3]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(g_empty);
    }
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
  KL_CHECK(d.K <= PKP && d.ldw % 32 == 0 && d.ldht % 32 == 0 && d.Q == nullptr, KLNMF_EINVAL,
           "fused_coef_step256: k=%lld not supported", (long long)d.K);
  CUtensorMap tmHt, tmH, tmX;
  KL_TRY(make_map_ex(&tmHt, d.Ht, d.ldht, d.F, d.ldht, 32, PHN, CU_TENSOR_MAP_SWIZZLE_128B));   // 32 k x 16 feature rows
  KL_TRY(make_map_ex(&tmH, d.H, d.F, d.K, d.ldh, 32, PKH, CU_TENSOR_MAP_SWIZZLE_128B));          // 32 features x 128 rows
  KL_TRY(make_map_ex(&tmX, d.X, d.F, d.M, d.ldx, PHN, PBM, CU_TENSOR_MAP_SWIZZLE_64B));          // 16 features x 128 rows
  Fused256Params p{};
  p.M = d.M; p.F = d.F;
  p.n_blocks = (int)ceil_div(d.M, PBM);
  p.n_steps = (int)ceil_div(d.F, PBN);
  p.W = (const float *)d.W; p.ldw = d.ldw;
  p.Wout = (float *)d.Wout; p.ldwo = d.ldwo;
  p.w_cols = d.ldw < d.ldwo ? d.ldw : d.ldwo;
  p.kl = d.kl; p.stop = d.stop; p.err = err_dev;
  if (p.n_blocks == 0 || p.n_steps == 0) return KLNMF_OK;
  const int clusters = p.n_blocks < ctx->sm_count / 2 ? p.n_blocks : ctx->sm_count / 2;
  static bool attr_done = false;
  if (!attr_done) {
    KL_CUDA(cudaFuncSetAttribute(fused_coef256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES));
    attr_done = true;
  }
  fused_coef256_kernel<<<clusters * 2, P_THREADS, P_SMEM_BYTES, ctx->stream>>>(tmHt, tmH, tmX, p);
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

}  // namespace klnmf
