// tcgen05 / TMEM / TMA contraction engine for sm_100a with the KL-NMF epilogues fused in.
//
// One persistent, warp-specialised kernel serves the three contractions of a reference
// iteration (nmf.py:325-351) plus W0 = X.H0^T (nmf.py:156) and internal.dot(dico) (learner.py:81):
//
//   warp 0      TMA producer   : cp.async.bulk.tensor 2D boxes (128B swizzle) -> smem ring
//   warp 1      MMA issuer     : tcgen05.mma.cta_group::1.kind::tf32, accumulators in TMEM
//   warps 2..9  epilogue       : tcgen05.ld 32x32b -> registers -> fused epilogue -> global
//
// D tile = 128 x BN fp32 in TMEM, double buffered (2*BN <= 512 columns), so the epilogue of
// tile i overlaps the mainloop of tile i+1.  Operands may be K-major or MN-major (both are legal
// for kind::tf32), which lets W, H, Q be consumed exactly as they lie in HBM:
//   ratio      S = W.H        A = W  (K-major)   B = H  (MN-major)   epilogue: Q=(X+eps)/(S+eps), KL
//   coefficient G = Q.H^T     A = Q  (K-major)   B = H  (K-major)    epilogue: W' = W (.) G
//   numerator  N += W'^T.Q    A = W' (MN-major)  B = Q  (MN-major)   epilogue: red.add (split over rows)
// Split-TF32 ("tf32x3"): every operand is a (hi, lo) pair with hi exactly representable in TF32;
// three MMAs hi*hi + hi*lo + lo*hi per K step give FP32-grade products.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace klnmf {

namespace {

constexpr int BM = 128;          // tile rows  (UMMA M, cta_group::1)
constexpr int BK = 32;           // K per stage in floats = one 128-byte swizzle span
constexpr int NUM_THREADS = 320; // producer warp + MMA warp + 8 epilogue warps
constexpr int NUM_THREADS_XT = 352; // + one warp that streams X chunks by TMA (ratio epilogue)
constexpr int EPI_WARPS = 8;
constexpr int XCHUNK_BYTES = BM * 32 * 4;   // one 128-row x 32-column fp32 chunk of X or Q (128B-swizzled)
constexpr int QWARP_BYTES = 32 * 32 * 4;    // per epilogue warp: one 32-row x 32-column Q staging box

struct TcParams {
  int64_t M, N, K;
  int m_tiles, n_tiles, splits, m_fastest;
  int64_t kb_total, kb_per_split;        // K blocks (of BK)
  int epi, only_kl, accurate;
  float *out, *out_lo;
  int64_t ldo, n_store;                  // columns < n_store are written (multiple of 32, <= ldo)
  const float *aux, *aux_lo;
  int64_t ldaux;
  double *kl;
  const int *stop;
  int *err;
  uint32_t mn_lt, mn_lbo, mn_sbo, mn_kadv;   // MN-major descriptor parameters (bring-up overridable)
  int no_prefetch;                           // default 1; KLNMF_TC_PF=1 re-enables the L2 prefetch of the next X tile
  int relaxed;                               // accumulator hand-back with relaxed arrives (KLNMF_TC_RELAXED=0: release)
  float qshift;                              // EPI_RATIO: the stored ratio is q - qshift (centered ratio, api.cu)
  int round_out;                             // EPI_RATIO (TMA-staged form): store the ratio rounded to nearest TF32
  const float *colbias;                      // EPI_MULW: out = aux * (acc + colbias[col])
  int dbg;                                   // timing experiments (KLNMF_TC_DBG): 1 no ratio math, 2 no Q store, 4 one MMA per K block only
  uint32_t k_lt;                             // K-major layout type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (experiment)
};

// XT = the ratio epilogue with TMA-staged X and Q: three operand stages instead of four make
// room for XBUFS X chunks plus one Q staging chunk per epilogue half.
// XB = X chunks in flight per CTA (XT only; their HBM latency is hidden by the L2 prefetch of the next tile,
// so the ring covers L2 latency only).  Everything the X ring and the Q staging do not take goes to operand
// stages: with K = 512 the ratio contraction is bound by the TMA latency of its operand ring, not by smem.
// QIP = the ratio is written back IN PLACE into the X chunk it was computed from and leaves by TMA store from
// there: the 32 KB of Q staging boxes become two more X chunks in flight.
// Wide tiles (the plain contractions of a CTA pair, one-pass modes): MS = 2 stacks two 256-row blocks on one B tile
// (512 x 256: the numerator, M = k), BN = 512 puts two 256-column MMAs on one A block (256 x 512: the coefficient
// contraction, N = k).  A 256 x 256 pair tile pulls 64 B/clk per SM out of L2 at the full TF32 rate -- more than the
// fabric delivers; the wide tiles need 48 B/clk and read the ratio panel from HBM once instead of
// once per 256 components.  Their accumulator fills the whole TMEM (512 columns), so it is single-buffered: the
// epilogue of a tile (1-2 % of its mainloop at these contraction lengths) is not overlapped.
template <int BN, bool SPLIT, bool XT = false, int CG = 1, int XB = 2, bool QIP = false, int MS = 1>
struct Cfg {
  static constexpr int NI = BN > 256 ? 256 : BN;         // N of one tcgen05.mma
  static constexpr int NS = BN / NI;                     // MMAs side by side on one A block
  static constexpr int A_BYTES = MS * BM * BK * 4;
  static constexpr int B_BYTES = (BN / CG) * BK * 4;     // a CTA pair splits the B tile
  static constexpr int STAGE_BYTES = (A_BYTES + B_BYTES) * (SPLIT ? 2 : 1);
  static constexpr int ACC_COLS = MS * BN;               // TMEM columns of one accumulator
  static constexpr int NACC = 2 * ACC_COLS <= 512 ? 2 : 1;
  static constexpr int XBUFS = XB;
  // split-TF32 contractions (no XT form): one 32 x 32 transpose box per epilogue warp, so that X arrives and the
  // (hi, lo) ratio leaves as full 128-byte lines (epilogue_ratio_staged)
  static constexpr int XQ_BYTES = XT ? XB * XCHUNK_BYTES + (QIP ? 0 : EPI_WARPS * QWARP_BYTES) : (SPLIT ? EPI_WARPS * QWARP_BYTES : 0);
  static constexpr int STAGES = (XT ? (224 * 1024 - XQ_BYTES) : 192 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_COLS = NACC * ACC_COLS;   // power of two: 256 or 512
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + XQ_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------------------------------------
// fused epilogue on one 32-column chunk held by its row-owning thread
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_row32(float *p, const float v[32]) {
#pragma unroll
  for (int j = 0; j < 8; j++)
    *reinterpret_cast<float4 *>(p + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void ld_row32(const float *p, float v[32]) {
#pragma unroll
  for (int j = 0; j < 8; j++) {
    float4 t = __ldg(reinterpret_cast<const float4 *>(p + 4 * j));
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

template <bool SPLIT>
__device__ __forceinline__ void epilogue_chunk(const TcParams &p, int64_t row, int64_t col0, const uint32_t acc_u[32],
                                               double &kl) {
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; j++) acc[j] = __uint_as_float(acc_u[j]);
  if (row >= p.M || col0 >= p.n_store) return;
  const int epi = p.epi;
  if (epi == EPI_STORE) {
    if (p.only_kl) return;     // diagnostic (KLNMF_BENCH_NOSTORE): contraction without the output traffic
    if (p.out_lo) {
      float hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; j++) { hi[j] = tf32_round(acc[j]); lo[j] = acc[j] - hi[j]; }
      st_row32(p.out + row * p.ldo + col0, hi);
      st_row32(p.out_lo + row * p.ldo + col0, lo);
    } else {
      st_row32(p.out + row * p.ldo + col0, acc);
    }
  } else if (epi == EPI_RATIO) {
    float x[32], q[32];
    ld_row32(p.aux + row * p.ldaux + col0, x);
    float part = 0.f;
    if (p.accurate) {
      // TF32R on a tile shape the TMA-staged form does not serve (f <= 128, single CTAs): the same arithmetic as there --
      // cancellation-free objective, centered ratio rounded to nearest before the consumers multiply it
#pragma unroll
      for (int j = 0; j < 32; j++) part += ratio_term_cf(x[j], acc[j], q[j]);
    } else {
      // (split-TF32 contractions take epilogue_ratio_staged; SPLIT here only in bring-up configurations)
#pragma unroll
      for (int j = 0; j < 32; j++) part += ratio_term<SPLIT>(x[j], acc[j], q[j]);
    }
    // columns >= N hold x = 0, s = 0: q = 1, the term is exactly 0
    kl += (double)part;
    if (!p.only_kl) {
#pragma unroll
      for (int j = 0; j < 32; j++) q[j] -= p.qshift;
      if (p.round_out && !p.out_lo) {
#pragma unroll
        for (int j = 0; j < 32; j++) q[j] = tf32_round(q[j]);
      }
      if (p.out_lo) {
        float lo[32];
#pragma unroll
        for (int j = 0; j < 32; j++) { float h = tf32_round(q[j]); lo[j] = q[j] - h; q[j] = h; }
        st_row32(p.out + row * p.ldo + col0, q);
        st_row32(p.out_lo + row * p.ldo + col0, lo);
      } else {
        st_row32(p.out + row * p.ldo + col0, q);
      }
    }
  } else if (epi == EPI_MULW) {
    float w[32];
    ld_row32(p.aux + row * p.ldaux + col0, w);
    if (p.aux_lo) {
      float wl[32];
      ld_row32(p.aux_lo + row * p.ldaux + col0, wl);
#pragma unroll
      for (int j = 0; j < 32; j++) w[j] += wl[j];
    }
    if (p.colbias) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(p.colbias + col0 + 4 * j));
        // Q.H^T is a sum of non-negative terms; its centered form (Q-1).H^T + rowsum(H) can come out a rounding error
        // below zero where the ratio vanishes (all-zero samples), and a negative coefficient would poison log(W.H)
        acc[4 * j] = fmaxf(acc[4 * j] + b.x, 0.f); acc[4 * j + 1] = fmaxf(acc[4 * j + 1] + b.y, 0.f);
        acc[4 * j + 2] = fmaxf(acc[4 * j + 2] + b.z, 0.f); acc[4 * j + 3] = fmaxf(acc[4 * j + 3] + b.w, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; j++) w[j] *= acc[j];
    if (p.out_lo) {
      float lo[32];
#pragma unroll
      for (int j = 0; j < 32; j++) { float h = tf32_round(w[j]); lo[j] = w[j] - h; w[j] = h; }
      st_row32(p.out + row * p.ldo + col0, w);
      st_row32(p.out_lo + row * p.ldo + col0, lo);
    } else {
      st_row32(p.out + row * p.ldo + col0, w);
    }
  } else {  // EPI_ACC
    float *o = p.out + row * p.ldo + col0;
#pragma unroll
    for (int j = 0; j < 8; j++)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(acc[4 * j]),
                   "f"(acc[4 * j + 1]), "f"(acc[4 * j + 2]), "f"(acc[4 * j + 3])
                   : "memory");
  }
}

// (Fetching the X block one piece ahead into registers was measured: 168 registers with spills, ratio contraction
// 11.6 -> 14.9 ms at the cfg5 shape -- dropped.)
// The ratio epilogue of the split-TF32 mode on one 32-row x 32-column block of a warp.  The row-owning thread
// layout of tcgen05.ld makes every direct global access a 16-byte piece per row; here the block goes through a
// warp-private 4 KB box (16-byte chunks XOR-swizzled by the row): X is read and Q_hi / Q_lo are written as
// four full 128-byte lines per instruction, and the box is read / written by the row owners without bank conflicts.
template <bool SPLIT>
__device__ __forceinline__ void epilogue_ratio_staged(const TcParams &p, int64_t row0, int lane, int64_t col0,
                                                      const uint32_t acc_u[32], double &kl, uint8_t *box) {
  if (col0 >= p.n_store) return;                 // warp-uniform
  const int lr = lane >> 3, lc = lane & 7;       // coalesced phase: 4 rows x 8 chunks of 16 bytes per instruction
  const uint32_t sw = (uint32_t)(lane & 7);
  uint8_t *mine = box + lane * 128;
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const int r = 4 * it + lr;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < p.M) t = __ldg(reinterpret_cast<const float4 *>(p.aux + (row0 + r) * p.ldaux + col0 + 4 * lc));
    *reinterpret_cast<float4 *>(box + r * 128 + ((lc ^ (r & 7)) << 4)) = t;
  }
  __syncwarp();
  float q[32];
  float part = 0.f;
  if (p.accurate) {
    f32x2_t part2 = splat2(0.f);
    const bool store_u = p.qshift == 1.f;          // centered: u = q - 1 without the cancellation
    const f32x2_t nshift2 = splat2(-p.qshift);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float4 t = *reinterpret_cast<const float4 *>(mine + ((j ^ sw) << 4));
      f32x2_t qa, ua, qb, ub;
      part2 = add2(part2, ratio_pair_cf<true>(pack2(t.x, t.y), pack2(__uint_as_float(acc_u[4 * j]), __uint_as_float(acc_u[4 * j + 1])), qa, ua));
      part2 = add2(part2, ratio_pair_cf<true>(pack2(t.z, t.w), pack2(__uint_as_float(acc_u[4 * j + 2]), __uint_as_float(acc_u[4 * j + 3])), qb, ub));
      qa = store_u ? ua : add2(qa, nshift2);
      qb = store_u ? ub : add2(qb, nshift2);
      unpack2(qa, q[4 * j], q[4 * j + 1]);
      unpack2(qb, q[4 * j + 2], q[4 * j + 3]);
    }
    float pa, pb;
    unpack2(part2, pa, pb);
    part = pa + pb;
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const float4 t = *reinterpret_cast<const float4 *>(mine + ((j ^ sw) << 4));
      part += ratio_term<false>(t.x, __uint_as_float(acc_u[4 * j]), q[4 * j]);
      part += ratio_term<false>(t.y, __uint_as_float(acc_u[4 * j + 1]), q[4 * j + 1]);
      part += ratio_term<false>(t.z, __uint_as_float(acc_u[4 * j + 2]), q[4 * j + 2]);
      part += ratio_term<false>(t.w, __uint_as_float(acc_u[4 * j + 3]), q[4 * j + 3]);
    }
#pragma unroll
    for (int j = 0; j < 32; j++) q[j] -= p.qshift;
  }
  // rows >= M hold x = 0 against an accumulator of 0: q = 1, the term is exactly 0
  kl += (double)part;
  if (p.only_kl) return;                         // warp-uniform
  const bool two = SPLIT && p.out_lo != nullptr;
#pragma unroll 1
  for (int part_i = 0; part_i < (two ? 2 : 1); part_i++) {
    // pass 0: the TF32-exact high part (or the plain value), pass 1: the residual
    __syncwarp();                                // the coalesced reads of the previous pass are done
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float4 t;
      if (two && part_i == 0) {
        t = make_float4(tf32_round(q[4 * j]), tf32_round(q[4 * j + 1]), tf32_round(q[4 * j + 2]), tf32_round(q[4 * j + 3]));
        q[4 * j] -= t.x; q[4 * j + 1] -= t.y; q[4 * j + 2] -= t.z; q[4 * j + 3] -= t.w;
      } else {
        t = make_float4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
      }
      *reinterpret_cast<float4 *>(mine + ((j ^ sw) << 4)) = t;
    }
    __syncwarp();
    float *dst = part_i == 0 ? p.out : p.out_lo;
#pragma unroll
    for (int it = 0; it < 8; it++) {
      const int r = 4 * it + lr;
      const float4 t = *reinterpret_cast<const float4 *>(box + r * 128 + ((lc ^ (r & 7)) << 4));
      if (row0 + r < p.M) *reinterpret_cast<float4 *>(dst + (row0 + r) * p.ldo + col0 + 4 * lc) = t;
    }
  }
  __syncwarp();                                  // the box is free for the next block
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// XT = true: the ratio contraction (EPI_RATIO, A K-major, B MN-major) with X streamed into a
// 128B-swizzled smem ring by TMA (warp 10) and Q leaving through smem + TMA store, so that both
// cross HBM as full 128-byte lines instead of one 16-byte piece per thread and row.
template <int BN, bool A_MN, bool B_MN, bool SPLIT, bool XT, int CG, int XB, bool QIP, int MS>
__global__ void __launch_bounds__(XT ? NUM_THREADS_XT : NUM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ, const TcParams p) {
  using C = Cfg<BN, SPLIT, XT, CG, XB, QIP, MS>;
  constexpr int STAGES = C::STAGES;
  constexpr int XBUFS = C::XBUFS;
  constexpr int NI = C::NI, NS = C::NS, NACC = C::NACC, ACC_COLS = C::ACC_COLS;
  static_assert(!XT || (MS == 1 && NS == 1), "the TMA-staged ratio epilogue works on 256-column tiles");
  static_assert((MS == 1 && NS == 1) || (CG == 2 && !SPLIT), "wide tiles exist for CTA pairs in the one-pass modes");
  if (p.stop != nullptr && *p.stop != 0) return;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;   // position in the CTA pair; 0 issues the MMAs

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t xq_base = smem_base + STAGES * C::STAGE_BYTES;          // X ring, then Q staging (XT only)
  const uint32_t bar_base = xq_base + C::XQ_BYTES;
  // barrier block: full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem_ptr | xfull[XBUFS] | xempty[XBUFS]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * STAGES + 4);
  auto xfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 5 + b); };
  auto xempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 5 + XBUFS + b); };
  volatile uint32_t *tmem_ptr_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + STAGES * C::STAGE_BYTES +
                                                                           C::XQ_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (SPLIT) { tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmBlo); }
    if (XT) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmQ); }
    for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS * CG); }
    if (XT)
      for (int b = 0; b < XBUFS; b++) { mbar_init(xfull_bar(b), 1); mbar_init(xempty_bar(b), EPI_WARPS / 2); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_ptr_addr, C::TMEM_COLS);
  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();     // barrier inits visible to the peer before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;

  // work units are per CTA pair when CG == 2 (p.m_tiles counts 256-row tiles then)
  const int64_t total_units = (int64_t)p.m_tiles * p.n_tiles * p.splits;
  const int64_t u_first = blockIdx.x / CG, u_step = gridDim.x / CG;
  auto decode = [&](int64_t u, int &mi, int &ni, int &si) {
    if (p.m_fastest) { mi = (int)(u % p.m_tiles); u /= p.m_tiles; ni = (int)(u % p.n_tiles); si = (int)(u / p.n_tiles); }
    else { ni = (int)(u % p.n_tiles); u /= p.n_tiles; mi = (int)(u % p.m_tiles); si = (int)(u / p.m_tiles); }
  };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // MN-major operands need one 32 x 32 box per 32-wide M/N group (up to 4 + 8 boxes per stage, twice
    // that for split-TF32): every lane issues at most one box, so a stage costs one issue slot of the
    // warp instead of a dozen serial UTMALDGs on one thread.
    {
      // boxes per stage: MN-major operands one 32 x 32 box per 32-wide M/N group, K-major operands one box per
      // 128-row block (MS blocks of A, NS blocks of B per CTA)
      constexpr int NA = A_MN ? MS * BM / 32 : MS, NB = B_MN ? BN / 32 / CG : NS;
      constexpr int GB = NI / 32 / CG;         // 32-wide groups of one MMA's B block held by this CTA
      static_assert(NA + NB <= 16, "lo operands use lanes 16..31");
      const int l = lane & 15;
      const bool is_lo = lane >= 16;
      const bool is_a = l < NA;
      const int grp = is_a ? l : l - NA;
      const bool active = l < NA + NB && (SPLIT || !is_lo);
      const CUtensorMap *map = is_a ? (is_lo ? &tmAlo : &tmA) : (is_lo ? &tmBlo : &tmB);
      const bool mn = is_a ? A_MN : B_MN;
      // block (sub-tile) and 32-wide group inside it
      const int sub = is_a ? (A_MN ? grp / (BM / 32) : grp) : (B_MN ? grp / GB : grp);
      const int g32 = is_a ? (A_MN ? grp % (BM / 32) : 0) : (B_MN ? grp % GB : 0);
      const uint32_t dst_off = (is_a ? 0u : (uint32_t)C::A_BYTES * (SPLIT ? 2 : 1)) +
                               (is_lo ? (uint32_t)(is_a ? C::A_BYTES : C::B_BYTES) : 0u) +
                               (uint32_t)sub * (uint32_t)(BM * BK * 4) + (uint32_t)g32 * 4096u;
      static_assert((NI / CG) * BK * 4 == BM * BK * 4 || NS == 1, "a B block of a wide tile is 128 rows per CTA");
      int stage = 0; uint32_t phase = 0;
      for (int64_t u = u_first; u < total_units; u += u_step) {
        int mi, ni, si;
        decode(u, mi, ni, si);
        const int64_t kb0 = (int64_t)si * p.kb_per_split;
        const int64_t kb1 = kb0 + p.kb_per_split < p.kb_total ? kb0 + p.kb_per_split : p.kb_total;
        const int32_t mn0 = (is_a ? ((mi * MS + sub) * CG + (int)crank) * BM : ni * BN + sub * NI + (int)crank * (NI / CG)) + 32 * g32;
        for (int64_t kb = kb0; kb < kb1; kb++) {
          mbar_wait(empty_bar(stage), phase ^ 1, p.err, 1);
          // the leader's barrier collects the bytes of both CTAs of a pair
          if (lane == 0 && crank == 0) mbar_expect_tx(full_bar(stage), C::STAGE_BYTES * CG);
          __syncwarp();
          if (active) {
            const int32_t k0 = (int32_t)(kb * BK);
            const uint32_t dst = smem_base + stage * C::STAGE_BYTES + dst_off;
            if (CG == 1) {
              if (mn) tma_load_2d(dst, map, full_bar(stage), mn0, k0);
              else tma_load_2d(dst, map, full_bar(stage), k0, mn0);
            } else {
              const uint32_t bar = mapa(full_bar(stage), 0);
              if (mn) tma_load_2d_pair(dst, map, bar, mn0, k0);
              else tma_load_2d_pair(dst, map, bar, k0, mn0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp walks the loop (warp-uniform control flow and descriptors, so they live in uniform
    // registers); one elected lane issues the tcgen05 instructions.  Per K block: one barrier wait, four
    // (twelve) MMAs whose descriptors differ by 32-bit adds, one commit.
    if (crank == 0) {
      // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4 (M = 256 across a CTA pair)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(NI >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
      // K-major: +32 B inside the 128 B swizzle row; MN-major: next 8-row group (+1024 B)
      const uint32_t a_kadv = (A_MN ? p.mn_kadv : 32u) >> 4, b_kadv = (B_MN ? p.mn_kadv : 32u) >> 4;
      const uint32_t a_hi = desc_hi(A_MN ? p.mn_sbo : 1024u, A_MN ? p.mn_lt : p.k_lt);
      const uint32_t b_hi = desc_hi(B_MN ? p.mn_sbo : 1024u, B_MN ? p.mn_lt : p.k_lt);
      const uint32_t a_lo0 = desc_lo(smem_base, A_MN ? p.mn_lbo : 16u);
      const uint32_t b_lo0 = desc_lo(smem_base + C::A_BYTES * (SPLIT ? 2 : 1), B_MN ? p.mn_lbo : 16u);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t u = u_first; u < total_units; u += u_step) {
        int mi, ni, si;
        decode(u, mi, ni, si);
        const int64_t kb0 = (int64_t)si * p.kb_per_split;
        const int64_t kb1 = kb0 + p.kb_per_split < p.kb_total ? kb0 + p.kb_per_split : p.kb_total;
        const int nkb = (int)(kb1 - kb0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1, p.err, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(full_bar(stage), phase, p.err, 3);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t soff = (uint32_t)(stage * C::STAGE_BYTES) >> 4;
            const uint32_t al = a_lo0 + soff, bl = b_lo0 + soff;
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; kk++) {
              if ((p.dbg & 4) && kk > 0) break;
              const uint64_t da = desc_pack(al + kk * a_kadv, a_hi);
              const uint64_t db = desc_pack(bl + kk * b_kadv, b_hi);
              const uint32_t first = (kb > 0 || kk > 0) ? 1u : 0u;
              if (SPLIT) {
                const uint64_t dal = desc_pack(al + (C::A_BYTES >> 4) + kk * a_kadv, a_hi);
                const uint64_t dbl = desc_pack(bl + (C::B_BYTES >> 4) + kk * b_kadv, b_hi);
                umma_tf32<CG>(d_tmem, dal, db, idesc, first);   // lo*hi
                umma_tf32<CG>(d_tmem, da, dbl, idesc, 1u);      // hi*lo
                umma_tf32<CG>(d_tmem, da, db, idesc, 1u);       // hi*hi
              } else if (MS == 1 && NS == 1) {
                umma_tf32<CG>(d_tmem, da, db, idesc, first);
              } else {
                // wide tile: MS blocks of A x NS blocks of B, each block 128 rows x 32 floats (16 KB) per CTA
#pragma unroll
                for (int ms = 0; ms < MS; ms++)
#pragma unroll
                  for (int ns = 0; ns < NS; ns++)
                    umma_tf32<CG>(d_tmem + (uint32_t)((ms * NS + ns) * NI),
                                  desc_pack(al + ms * ((BM * BK * 4) >> 4) + kk * a_kadv, a_hi),
                                  desc_pack(bl + ns * ((BM * BK * 4) >> 4) + kk * b_kadv, b_hi), idesc, first);
              }
            }
            umma_commit<CG>(empty_bar(stage));            // smem slot free once these MMAs retire
            if (kb + 1 == nkb) umma_commit<CG>(tfull_bar(acc));   // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (XT && warp == 10) {
    // =============================== X loader (ratio epilogue) ===============================
    // The ring of XBUFS chunks (16 KB each) covers the HBM latency by itself: 13 B/clk of X per SM x ~2000 clk.
    // (A rolling chunk-wise L2 prefetch 4 / 8 / 16 chunks ahead was measured too: 4.44 -> 4.64 / 4.53 / 4.92 ms and
    // 10.7 -> 14.5 GB of DRAM reads per launch at n = 262144, cfg5 shape -- profiles/r1_s4_run41_ratio_x_prefetch.log.)
    if (lane == 0) {
      auto x_tile = [&](int64_t u, int32_t &m0, int32_t &n0, int &nch) {
        int mi, ni, si;
        decode(u, mi, ni, si);
        m0 = (mi * CG + (int)crank) * BM; n0 = ni * BN;
        const int64_t left = (p.n_store - (int64_t)n0) / 32;
        nch = (int)(left < BN / 32 ? left : BN / 32);
      };
      uint32_t g = 0;
      if (u_first < total_units && !p.no_prefetch) {
        int32_t m0, n0; int nch;
        x_tile(u_first, m0, n0, nch);
        for (int c = 0; c < nch; c++) tma_prefetch_l2_2d(&tmX, n0 + 32 * c, m0);
      }
      for (int64_t u = u_first; u < total_units; u += u_step) {
        int32_t m0, n0; int nch;
        if (u + u_step < total_units && !p.no_prefetch) {
          x_tile(u + u_step, m0, n0, nch);
          for (int c = 0; c < nch; c++) tma_prefetch_l2_2d(&tmX, n0 + 32 * c, m0);
        }
        x_tile(u, m0, n0, nch);
        for (int c = 0; c < nch; c++, g++) {
          const uint32_t b = g % XBUFS, ph = (g / XBUFS) & 1u;
          mbar_wait(xempty_bar(b), ph ^ 1u, p.err, 5);
          mbar_expect_tx(xfull_bar(b), XCHUNK_BYTES);
          tma_load_2d(xq_base + b * XCHUNK_BYTES, &tmX, xfull_bar(b), n0 + 32 * c, m0);
        }
      }
    }
  } else if (XT) {
    // =============================== ratio epilogue, X and Q through shared memory ===============================
    const int e = warp - 2;                 // 0..7
    const int quarter = warp & 3;           // TMEM lanes this warp may touch: [32*quarter, +32)
    const int half = e >> 2;                // chunks c with (c & 1) == half belong to this half
    const int r = quarter * 32 + lane;      // row inside the tile = TMEM lane
    const uint32_t sw = (uint32_t)(lane & 7);
    uint8_t *xq_gen = smem_gen + STAGES * C::STAGE_BYTES;
    // Q leaves through one private 32 x 32 staging box per warp (128B-swizzled) and one TMA store per chunk:
    // no barrier between warps; the store of chunk i has long finished reading the box when the arithmetic of
    // chunk i+1 is done
    uint8_t *qst_gen = xq_gen + XBUFS * XCHUNK_BYTES + e * QWARP_BYTES + lane * 128;
    const uint32_t qst = xq_base + XBUFS * XCHUNK_BYTES + e * QWARP_BYTES;
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t gbase = 0;
    double kl = 0.0;
    // QIP: the chunk whose Q store may still be reading its box; handed back to the X loader one chunk later
    // (or at the end of the tile), when that read has long finished
    bool pending = false;
    uint32_t pend_b = 0;
    auto release_pending = [&]() {
      if (pending) {
        if (lane == 0) { bulk_wait_read0(); mbar_arrive(xempty_bar(pend_b)); }
        pending = false;
      }
    };
    for (int64_t u = u_first; u < total_units; u += u_step) {
      int mi, ni, si;
      decode(u, mi, ni, si);
      const int32_t m0 = (mi * CG + (int)crank) * BM, n0 = ni * BN;
      const int64_t left = (p.n_store - (int64_t)n0) / 32;
      const int nch = (int)(left < BN / 32 ? left : BN / 32);
      mbar_wait(tfull_bar(acc), acc_phase, p.err, 4);
      tc_fence_after();
      float kl_tile = 0.f;      // <= 128 terms per thread in FP32, then FP64 across tiles
#pragma unroll 1
      for (int c = half; c < nch; c += 2) {
        const uint32_t g = gbase + c, b = g % XBUFS, ph = (g / XBUFS) & 1u;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + c * 32);
        uint32_t v[32];
        tmem_ld32_issue(taddr, v);
        if (QIP) release_pending();
        mbar_wait(xfull_bar(b), ph, p.err, 6);
        float x[32];
        const uint8_t *xrow = xq_gen + b * XCHUNK_BYTES + r * 128;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const float4 t = *reinterpret_cast<const float4 *>(xrow + ((j ^ sw) << 4));
          x[4 * j] = t.x; x[4 * j + 1] = t.y; x[4 * j + 2] = t.z; x[4 * j + 3] = t.w;
        }
        tmem_ld32_wait(v);
        __syncwarp();
        if ((!QIP || p.only_kl) && lane == 0) mbar_arrive(xempty_bar(b));
        // rows >= M and columns >= N hold x = 0, s = 0 (TMA zero fill): q = 1, the term is exactly 0
        // two elements per instruction (FFMA2 / FADD2 / FMUL2) and the arithmetic mode as a template parameter
        // (ratio_chunk32, tc_ptx.cuh): the eight epilogue warps are bound by their issue rate once the objective takes
        // its cancellation-free form.  TF32R (accurate): the consumers multiply exactly what is stored (rounded to
        // nearest; the tensor core would truncate)
        if (!(p.dbg & 1)) kl_tile += ratio_chunk32_dispatch(x, v, p.qshift, p.accurate);
        if (!p.only_kl) {
          if (QIP) {
            // in place: every thread overwrites the row of X it has just read; the warp's 32 rows are one 4 KB box
            uint8_t *qrow = xq_gen + b * XCHUNK_BYTES + r * 128;
#pragma unroll
            for (int j = 0; j < 8; j++)
              *reinterpret_cast<float4 *>(qrow + ((j ^ sw) << 4)) =
                  make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(p.dbg & 2)) tma_store_2d(&tmQ, xq_base + b * XCHUNK_BYTES + quarter * QWARP_BYTES, n0 + 32 * c, m0 + quarter * 32);
            pending = true;
            pend_b = b;
          } else {
            if (lane == 0) bulk_wait_read0();         // the previous store has finished reading the box
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; j++)
              *reinterpret_cast<float4 *>(qst_gen + ((j ^ sw) << 4)) =
                  make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0 && !(p.dbg & 2)) tma_store_2d(&tmQ, qst, n0 + 32 * c, m0 + quarter * 32);
          }
        }
      }
      if (QIP) release_pending();
      gbase += (uint32_t)nch;
      kl += (double)kl_tile;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // the accumulator of BOTH CTAs is free once all their epilogue warps are done
        if (p.relaxed) {
          if (CG == 1 || crank == 0) mbar_arrive_relaxed(tempty_bar(acc));
          else mbar_arrive_cluster_relaxed(mapa(tempty_bar(acc), 0));
        } else {
          if (CG == 1 || crank == 0) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(mapa(tempty_bar(acc), 0));
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all();
    if (p.kl != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
      if (lane == 0) atomicAdd(p.kl, kl);
    }
  } else if (warp < 10) {
    // =============================== epilogue ===============================
    const int e = warp - 2;                 // 0..7
    const int quarter = warp & 3;           // TMEM lanes this warp may touch: [32*quarter, +32)
    const int half = e >> 2;                // column half of the tile
    int acc = 0; uint32_t acc_phase = 0;
    double kl = 0.0;
    for (int64_t u = u_first; u < total_units; u += u_step) {
      int mi, ni, si;
      decode(u, mi, ni, si);
      mbar_wait(tfull_bar(acc), acc_phase, p.err, 4);
      tc_fence_after();
#pragma unroll 1
      for (int ms = 0; ms < MS; ms++) {
        const int64_t row = (((int64_t)mi * MS + ms) * CG + crank) * BM + quarter * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 64; c++) {
          const int col_in_tile = half * (BN / 2) + c * 32;
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + ms * BN + col_in_tile);
          uint32_t v[32];
          tmem_ld32(taddr, v);
          if (SPLIT && p.epi == EPI_RATIO)
            epilogue_ratio_staged<SPLIT>(p, row - lane, lane, (int64_t)ni * BN + col_in_tile, v, kl,
                                         smem_gen + STAGES * C::STAGE_BYTES + e * QWARP_BYTES);
          else
            epilogue_chunk<SPLIT>(p, row, (int64_t)ni * BN + col_in_tile, v, kl);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // the accumulator of BOTH CTAs is free once all their epilogue warps are done
        if (p.relaxed) {
          if (CG == 1 || crank == 0) mbar_arrive_relaxed(tempty_bar(acc));
          else mbar_arrive_cluster_relaxed(mapa(tempty_bar(acc), 0));
        } else {
          if (CG == 1 || crank == 0) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_cluster(mapa(tempty_bar(acc), 0));
        }
      }
      if (++acc == NACC) { acc = 0; acc_phase ^= 1; }
    }
    if (p.epi == EPI_RATIO && p.kl != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
      if (lane == 0) atomicAdd(p.kl, kl);
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync(); else __syncthreads();     // nobody leaves while the peer may still signal or read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct TcState {
  int *err_dev = nullptr;
};

template <int BN, bool A_MN, bool B_MN, bool SPLIT, bool XT = false, int CG = 1, int XB = 2, bool QIP = false, int MS = 1>
int launch_cfg(klnmf_ctx *ctx, const GemmDesc &d, TcParams p) {
  using C = Cfg<BN, SPLIT, XT, CG, XB, QIP, MS>;
  static_assert(C::STAGES >= 2, "pipeline too shallow");
  static_assert(C::SMEM_BYTES <= 232448, "shared memory budget exceeded");
  CUtensorMap tmA, tmAlo, tmB, tmBlo, tmX, tmQ;
  // A: K-major = memory M x K (inner K);  MN-major = memory K x M (inner M)
  // KLNMF_TC_KSWZ32=1 (experiment): K-major operands staged with the 32-byte-atom swizzle as well, i.e. the
  // SAME smem image a MN-major descriptor reads -- what a fused back-to-back kernel needs to use one
  // dictionary tile as the MN-major B of S = W.H and as the K-major B of G = Q.H^T
  static const bool kswz32 = getenv("KLNMF_TC_KSWZ32") && atoi(getenv("KLNMF_TC_KSWZ32")) == 1;
  p.k_lt = kswz32 ? 1u : 2u;
  if (!A_MN) KL_TRY(make_map(&tmA, d.A, d.K, d.M, d.a_sm, BM, kswz32));
  else KL_TRY(make_map(&tmA, d.A, d.M, d.K, d.a_sk, 32, true));
  if (!B_MN) KL_TRY(make_map(&tmB, d.B, d.K, d.N, d.b_sn, C::NI / CG, kswz32));
  else KL_TRY(make_map(&tmB, d.B, d.N, d.K, d.b_sk, 32, true));
  tmAlo = tmA;
  tmBlo = tmB;
  if (SPLIT) {
    if (!A_MN) KL_TRY(make_map(&tmAlo, d.A_lo, d.K, d.M, d.a_sm, BM, kswz32));
    else KL_TRY(make_map(&tmAlo, d.A_lo, d.M, d.K, d.a_sk, 32, true));
    if (!B_MN) KL_TRY(make_map(&tmBlo, d.B_lo, d.K, d.N, d.b_sn, C::NI / CG, kswz32));
    else KL_TRY(make_map(&tmBlo, d.B_lo, d.N, d.K, d.b_sk, 32, true));
  }
  tmX = tmA;
  tmQ = tmA;
  if (XT) {
    // X (aux) and Q (out): row-major M x N, 32-column x 128-row boxes, 128B swizzle; columns >= N and
    // rows >= M are zero-filled on load and clipped on store
    KL_TRY(make_map(&tmX, d.aux, d.N, d.M, d.ldaux, BM, false));
    if (!p.only_kl) KL_TRY(make_map(&tmQ, d.out, d.N, d.M, d.ldo, 32, false));
  }
  p.m_tiles = (int)ceil_div(d.M, BM * CG * MS);      // tiles of a CTA pair span 256 (MS = 2: 512) rows
  p.n_tiles = (int)ceil_div(d.N, BN);
  const int64_t slots = ctx->sm_count / CG;     // CTAs or CTA pairs resident at once
  p.kb_total = ceil_div(d.K, BK);
  p.splits = 1;
  p.kb_per_split = p.kb_total;
  if (p.epi == EPI_ACC) {
    // split the contraction so that the persistent grid is filled evenly (>= 95 % wave efficiency)
    const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
    const int64_t max_s = p.kb_total / 16 > 0 ? p.kb_total / 16 : 1;   // >= 512 of K per unit
    int64_t best = 1; double best_eff = 0.0;
    for (int64_t s = 1; s <= max_s && s <= 64; s++) {
      const int64_t units = tiles * s;
      const double eff = (double)units / (double)(ceil_div(units, slots) * slots);
      if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
      if (eff >= 0.95) { best = s; break; }
    }
    p.kb_per_split = ceil_div(p.kb_total, best);
    p.splits = (int)ceil_div(p.kb_total, p.kb_per_split);
  }
  const int64_t units = (int64_t)p.m_tiles * p.n_tiles * p.splits;
  if (units == 0 || p.kb_total == 0) return KLNMF_OK;
  const int grid = (int)(units < slots ? units : slots) * CG;
  auto kern = tc_gemm_kernel<BN, A_MN, B_MN, SPLIT, XT, CG, XB, QIP, MS>;
  // per DEVICE: function attributes live in the device's context, and one process may drive several GPUs
  static bool attr_done[64] = {};
  if (!attr_done[ctx->device & 63]) {
    KL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_done[ctx->device & 63] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(XT ? NUM_THREADS_XT : NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KL_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmAlo, tmB, tmBlo, tmX, tmQ, p));
  ctx->n_launch++;
  KL_CUDA(cudaGetLastError());
  return KLNMF_OK;
}

// Wide tiles (N > 128) run on CTA pairs (cta_group::2); KLNMF_TC_CG=1 keeps single CTAs for comparison.
template <bool A_MN, bool B_MN>
int launch_major(klnmf_ctx *ctx, const GemmDesc &d, const TcParams &p, bool split, bool narrow) {
  static const bool pair = !(getenv("KLNMF_TC_CG") && atoi(getenv("KLNMF_TC_CG")) == 1);
  if (narrow) return split ? launch_cfg<128, A_MN, B_MN, true>(ctx, d, p) : launch_cfg<128, A_MN, B_MN, false>(ctx, d, p);
  if (pair)
    return split ? launch_cfg<256, A_MN, B_MN, true, false, 2>(ctx, d, p) : launch_cfg<256, A_MN, B_MN, false, false, 2>(ctx, d, p);
  return split ? launch_cfg<256, A_MN, B_MN, true>(ctx, d, p) : launch_cfg<256, A_MN, B_MN, false>(ctx, d, p);
}

}  // namespace

int tc_gemm(klnmf_ctx *ctx, int epi, const GemmDesc &d) {
  if (d.M <= 0 || d.N <= 0) return KLNMF_OK;
  const bool a_mn = d.a_sm == 1 && d.a_sk != 1;
  const bool b_mn = d.b_sn == 1 && d.b_sk != 1;
  KL_CHECK(a_mn || d.a_sk == 1, KLNMF_EINVAL, "tc_gemm: A must have a unit stride");
  KL_CHECK(b_mn || d.b_sk == 1, KLNMF_EINVAL, "tc_gemm: B must have a unit stride");
  TcState *st = (TcState *)ctx->tc;
  if (!st) {
    st = new TcState();
    if (cudaMalloc((void **)&st->err_dev, 4) != cudaSuccess) {
      delete st;
      set_error("tc_gemm: cudaMalloc failed");
      return KLNMF_ENOMEM;
    }
    cudaMemsetAsync(st->err_dev, 0, 4, ctx->stream);
    ctx->tc = st;
  }
  const bool split = ctx->split && !d.single_pass && d.A_lo != nullptr && d.B_lo != nullptr;
  KL_CHECK(!ctx->split || split || d.single_pass, KLNMF_EINVAL, "tc_gemm: split-TF32 mode needs (hi, lo) operands");
  TcParams p{};
  p.M = d.M; p.N = d.N; p.K = d.K;
  p.epi = epi; p.only_kl = d.only_kl;
  p.qshift = d.qshift; p.colbias = d.colbias; p.round_out = d.round_out;
  if (epi == EPI_STORE && getenv("KLNMF_BENCH_NOSTORE")) p.only_kl = 1;
  // The split-TF32 ratio epilogue evaluates the objective in its cancellation-free form (ratio_term_cf, tc_ptx.cuh): no
  // IEEE division, no logf, and more accurate than both once a fit has converged (cfg3 shape: 21.1 -> 17.8 ms; the
  // epilogue, not the three MMAs per step, paces this contraction).  The plain MUFU forms (KLNMF_TC_FASTMATH=1, 15.8 ms)
  // give the same W and H to three digits, but lg2.approx is off by ~2e-7 of sum(X) in the objective: 1.5e-4 on the
  // 200-iteration golden case against the stated 2e-5 (profiles/r1_s4_run50_*.log, r1_s4_run52_*.log, r1_s4_run56_*.log).
  p.accurate = (ctx->mode == KLNMF_MODE_TF32X3 || ctx->mode == KLNMF_MODE_TF32R) ? 1 : 0;
  if (getenv("KLNMF_TC_FASTMATH") && atoi(getenv("KLNMF_TC_FASTMATH")) == 1) p.accurate = 0;
  p.out = (float *)d.out; p.out_lo = (float *)d.out_lo; p.ldo = d.ldo;
  p.n_store = round_up(d.N, 32);
  KL_CHECK(epi == EPI_RATIO && d.only_kl ? true : p.n_store <= d.ldo, KLNMF_EINVAL,
           "tc_gemm: output leading dimension %lld too small for N=%lld rounded to 32", (long long)d.ldo, (long long)d.N);
  p.aux = (const float *)d.aux; p.aux_lo = (const float *)d.aux_lo; p.ldaux = d.ldaux;
  KL_CHECK(!(epi == EPI_RATIO || epi == EPI_MULW) || (d.aux && d.ldaux >= p.n_store), KLNMF_EINVAL,
           "tc_gemm: aux operand missing or its leading dimension is below N rounded to 32");
  p.kl = d.kl; p.stop = d.stop; p.err = st->err_dev;
  p.m_fastest = (epi == EPI_ACC) ? 1 : 0;
  p.mn_lt = 1; p.mn_lbo = 4096; p.mn_sbo = 512; p.mn_kadv = 1024;
  // measured (profiles/r1_dram_traffic_variants.log): cp.async.bulk.prefetch.tensor of the next X tile made the
  // kernel read X from HBM twice (8.6 GB instead of 4.6 GB at n = 131072) and 14 % slower -- off by default
  p.dbg = getenv("KLNMF_TC_DBG") ? atoi(getenv("KLNMF_TC_DBG")) : 0;
  p.no_prefetch = !(getenv("KLNMF_TC_PF") && atoi(getenv("KLNMF_TC_PF")) == 1);
  p.relaxed = !(getenv("KLNMF_TC_RELAXED") && atoi(getenv("KLNMF_TC_RELAXED")) == 0);
  if (const char *o = getenv("KLNMF_TC_MN")) {   // "layout_type,lbo,sbo,kadv" (bring-up only)
    unsigned a, b, c, e;
    if (sscanf(o, "%u,%u,%u,%u", &a, &b, &c, &e) == 4) { p.mn_lt = a; p.mn_lbo = b; p.mn_sbo = c; p.mn_kadv = e; }
  }
  const char *force = getenv("KLNMF_TC_BN");
  bool narrow = d.N <= 128;
  if (force) narrow = atoi(force) == 128;
  // wide tiles of the one-pass modes (Cfg): 256 x 512 when the contraction has more than 256 output columns and both
  // operands are K-major (coefficient contraction, W0 = X.H0^T: N = k), 512 x 256 for the numerator (M = k > 256, both
  // operands MN-major).  KLNMF_TC_WIDE=0 keeps the 256 x 256 pair tiles for comparison.
  static const bool wide = !(getenv("KLNMF_TC_WIDE") && atoi(getenv("KLNMF_TC_WIDE")) == 0) &&
                           !(getenv("KLNMF_TC_CG") && atoi(getenv("KLNMF_TC_CG")) == 1);
  // 256 x 512 for the coefficient contraction is OFF by default: its single-buffered accumulator exposes the epilogue
  // (W hi + lo in, W' hi + lo out), and at the cfg5 shape that costs more than the narrower operand stream saves
  // (n = 262144: 3.00 ms with 256 x 256 tiles, 3.60 ms wide; profiles/r2_run5.log).  KLNMF_TC_WIDE_N=1 selects it.
  static const bool wide_n = getenv("KLNMF_TC_WIDE_N") && atoi(getenv("KLNMF_TC_WIDE_N")) == 1;
  static const bool wide_m = !(getenv("KLNMF_TC_WIDE_M") && atoi(getenv("KLNMF_TC_WIDE_M")) == 0);
  if (wide && !split && !force) {
    if (wide_n && !a_mn && !b_mn && d.N > 256 && (epi == EPI_MULW || epi == EPI_STORE))
      return launch_cfg<512, false, false, false, false, 2>(ctx, d, p);
    if (wide_m && a_mn && b_mn && d.M > 256 && d.N > 128 && epi == EPI_ACC)
      return launch_cfg<256, true, true, false, false, 2, 2, false, 2>(ctx, d, p);
  }
  if (!a_mn && !b_mn) return launch_major<false, false>(ctx, d, p, split, narrow);
  // the ratio contraction of the loop: X and Q go through shared memory by TMA
  if (!a_mn && b_mn && epi == EPI_RATIO && !split && !narrow && !getenv("KLNMF_TC_NO_XT")) {
    if (getenv("KLNMF_TC_CG") && atoi(getenv("KLNMF_TC_CG")) == 1) return launch_cfg<256, false, true, false, true, 1, 4>(ctx, d, p);
    // Ring balance (profiles/r1_ratio_ring_balance.log, r1_s4_run42/43_*.log): k >= 512 wants 5 operand stages; the
    // ratio is then written back IN PLACE into the X chunk it came from (QIP), which turns the 32 KB of Q staging boxes
    // into two more X chunks in flight (4.51 -> 4.31 ms at n = 262144, cfg5 shape).  Shorter contractions consume X
    // faster than operands: 4 stages + 4 chunks + staging boxes.  KLNMF_TC_XB / KLNMF_TC_QIP force a variant.
    int xb = d.K >= 512 ? 2 : 4;
    bool qip = d.K >= 512;
    if (getenv("KLNMF_TC_XB")) xb = atoi(getenv("KLNMF_TC_XB"));
    if (getenv("KLNMF_TC_QIP")) qip = atoi(getenv("KLNMF_TC_QIP")) == 1;
    if (qip) {
      if (xb == 6) return launch_cfg<256, false, true, false, true, 2, 6, true>(ctx, d, p);
      return launch_cfg<256, false, true, false, true, 2, 4, true>(ctx, d, p);
    }
    if (xb == 2) return launch_cfg<256, false, true, false, true, 2, 2>(ctx, d, p);
    return launch_cfg<256, false, true, false, true, 2, 4>(ctx, d, p);
  }
  if (!a_mn && b_mn) return launch_major<false, true>(ctx, d, p, split, narrow);
  if (a_mn && b_mn) return launch_major<true, true>(ctx, d, p, split, narrow);
  return launch_major<true, false>(ctx, d, p, split, narrow);
}

int tc_selftest(int *n_fail, char *, int) { *n_fail = 0; return KLNMF_OK; }

void tc_release(klnmf_ctx *ctx) {
  TcState *st = (TcState *)ctx->tc;
  if (!st) return;
  if (st->err_dev) cudaFree(st->err_dev);
  delete st;
  ctx->tc = nullptr;
}

}  // namespace klnmf
