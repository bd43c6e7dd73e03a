// PTX wrappers shared by the tcgen05 kernels of libklnmf (dense_tc.cu, dense_fused.cu): mbarriers, TMA,
// tcgen05 alloc / mma / commit / ld, shared-memory matrix descriptors, MUFU helpers.  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace klnmf {
namespace {

constexpr int UMMA_K = 8;        // kind::tf32: 32 bytes of K per instruction

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as an error, never as a hung GPU.  The bound only has to be far above any
// legitimate wait (microseconds); build with -DKLNMF_WATCHDOG_NS=... to change it (e.g. on a time-sliced GPU, where a
// preempted CTA can wait for seconds), -DKLNMF_WATCHDOG_NS=0 to wait forever.
#ifndef KLNMF_WATCHDOG_NS
#define KLNMF_WATCHDOG_NS 10000000000ull   // 10 s
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int *err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    for (int i = 0; i < 2048; i++)
      if (mbar_try_wait(bar, parity)) return;
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (KLNMF_WATCHDOG_NS != 0 && t1 - t0 > KLNMF_WATCHDOG_NS) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      asm volatile("trap;");
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap *map, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// CG = 1: one CTA per tile.  CG = 2: a CTA pair (cluster of 2 on one TPC) works on a 256-row tile with
// tcgen05 cta_group::2 -- each CTA stages its own 128 rows of A and HALF of the B tile, the leader issues
// the MMAs for both, accumulators land in each CTA's own TMEM.
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// completion of all MMAs issued so far by this thread -> mbarrier (CG = 2: the barrier at the same smem
// offset in BOTH CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {   // shared::cta -> shared::cluster of CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// hand-back of a TMEM accumulator: nothing in memory is published (the tcgen05.ld results are already in
// registers, ordered by tcgen05.wait::ld + fence::before_thread_sync), so a relaxed arrive is enough --
// the default release form costs a MEMBAR + ERRBAR per tile and warp
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by one CTA of a pair into its OWN smem, completing bytes on a barrier of either CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *map, uint32_t bar_cluster, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load split in two, so that the X chunk can be fetched from shared memory while TMEM is read:
// the wait names the 32 registers as in/out operands, which keeps every use of them behind it
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t v[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
// MUFU approximations (flush-to-zero forms: no denormal range fix-up code, so 32 independent
// chains per thread interleave freely).  Arguments here are >= eps = 1e-8, far from denormals.
__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float lg2_approx(float v) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// one element of the ratio epilogue (nmf.py:325-336 + metrics.py:18-20): q = (x+eps)/(s+eps),
// returns x*log(q) + s - x
template <bool ACCURATE>
__device__ __forceinline__ float ratio_term(float x, float s, float &q) {
  if (ACCURATE) {
    q = (x + (float)KL_EPS) / (s + (float)KL_EPS);
    return fmaf(x, logf(q), s - x);
  }
  q = (x + (float)KL_EPS) * rcp_approx(s + (float)KL_EPS);
  return fmaf(x * 0.69314718055994531f, lg2_approx(q), s - x);
}
// The same element in a form that needs no accurate logarithm (split-TF32 epilogue).  With d = s + eps,
// u = (x - s)/d = q - 1 and P(u) = log1p(u) - u:
//     x log(q) - x + s  =  x P(u) + (x - s)(x - s - eps)/d
// -- every product on the right is small when the fit has converged, where the left side cancels x log(q) against
// s - x.  Used for |u| < 1/4 with P(u) a short polynomial (truncation 3.5e-7 u^2); beyond, the left side is evaluated
// as it stands with lg2.approx (|log q| >= 0.22 there, and for large u the right side would cancel x u against
// (x - s) u instead: measured 1.7e-4 on the objective of tests' rise_case, whose ratios reach 1e3).
// q = (x + eps) * rcp(d) with one Newton step on the reciprocal.
__device__ __forceinline__ float ratio_term_cf(float x, float s, float &q) {
  const float d = s + (float)KL_EPS;
  float r = rcp_approx(d);
  r = fmaf(fmaf(-d, r, 1.f), r, r);
  q = (x + (float)KL_EPS) * r;
  const float xs = x - s;
  const float u = xs * r;
  float R = -0.1f;
  R = fmaf(R, u, 1.f / 9.f);
  R = fmaf(R, u, -0.125f);
  R = fmaf(R, u, 1.f / 7.f);
  R = fmaf(R, u, -1.f / 6.f);
  R = fmaf(R, u, 0.2f);
  R = fmaf(R, u, -0.25f);
  R = fmaf(R, u, 1.f / 3.f);
  R = fmaf(R, u, -0.5f);
  const float p_small = u * u * R;
  // |u| >= 1/4: the plain form -- there |log q| >= 0.22 makes the MUFU's 1e-7 harmless, and the products of the
  // cancellation-free form (x u against (x - s) u) would themselves cancel once u is large
  const float t_large = fmaf(x * 0.69314718055994531f, lg2_approx(fmaxf(q, 1e-37f)), -xs);
  const float t_small = fmaf(x, p_small, xs * (xs - (float)KL_EPS) * r);
  return fabsf(u) < 0.25f ? t_small : t_large;
}
// ---- the same two forms on PAIRS of elements with the packed FP32 instructions of sm_100 (fma/add/mul.f32x2 ->
// FFMA2 / FADD2 / FMUL2: two FP32 operations per issue slot).  The ratio epilogue is bound by the issue rate of its
// eight warps, not by the tensor pipe, once the objective takes its cancellation-free form (~31 instructions per
// element in scalar code, ~16 here); the MUFU operations stay scalar. ----
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t splat2(float a) { return pack2(a, a); }
// two elements of ratio_term<false>: q = (x+eps) * rcp(s+eps); returns the packed terms x ln2 lg2(q) + (s - x)
__device__ __forceinline__ f32x2_t ratio_pair_fast(f32x2_t x, f32x2_t s, f32x2_t &q) {
  const f32x2_t eps2 = splat2((float)KL_EPS);
  const f32x2_t d = add2(s, eps2);
  float d0, d1;
  unpack2(d, d0, d1);
  const f32x2_t r = pack2(rcp_approx(d0), rcp_approx(d1));
  q = mul2(add2(x, eps2), r);
  float q0, q1;
  unpack2(q, q0, q1);
  const f32x2_t l = pack2(lg2_approx(q0), lg2_approx(q1));
  const f32x2_t smx = fma2(x, splat2(-1.f), s);
  return fma2(mul2(x, splat2(0.69314718055994531f)), l, smx);
}
// two elements of ratio_term_cf; u = q - 1 is returned as well (the centered ratio, free of the cancellation of q - 1).
// NEWTON: one Newton step on the reciprocal (FP32-grade split mode); without it the MUFU's 1 ulp is kept (TF32R, whose
// stored ratio is rounded to TF32 anyway).
template <bool NEWTON>
__device__ __forceinline__ f32x2_t ratio_pair_cf(f32x2_t x, f32x2_t s, f32x2_t &q, f32x2_t &u) {
  const f32x2_t eps2 = splat2((float)KL_EPS);
  const f32x2_t d = add2(s, eps2);
  float d0, d1;
  unpack2(d, d0, d1);
  f32x2_t r = pack2(rcp_approx(d0), rcp_approx(d1));
  if (NEWTON) r = fma2(fma2(mul2(d, splat2(-1.f)), r, splat2(1.f)), r, r);
  q = mul2(add2(x, eps2), r);
  const f32x2_t xs = fma2(s, splat2(-1.f), x);
  u = mul2(xs, r);
  f32x2_t R = splat2(-0.1f);
  R = fma2(R, u, splat2(1.f / 9.f));
  R = fma2(R, u, splat2(-0.125f));
  R = fma2(R, u, splat2(1.f / 7.f));
  R = fma2(R, u, splat2(-1.f / 6.f));
  R = fma2(R, u, splat2(0.2f));
  R = fma2(R, u, splat2(-0.25f));
  R = fma2(R, u, splat2(1.f / 3.f));
  R = fma2(R, u, splat2(-0.5f));
  const f32x2_t p_small = mul2(mul2(u, u), R);
  const f32x2_t tail = mul2(mul2(xs, add2(xs, splat2(-(float)KL_EPS))), r);
  const f32x2_t t_small = fma2(x, p_small, tail);
  float q0, q1, u0, u1, ts0, ts1, tl0, tl1;
  unpack2(q, q0, q1);
  unpack2(u, u0, u1);
  // q > 0 in exact arithmetic; the clamp keeps lg2 finite where x = 0 and (x+eps) * r underflows
  const f32x2_t l = pack2(lg2_approx(fmaxf(q0, 1e-37f)), lg2_approx(fmaxf(q1, 1e-37f)));
  // |u| >= 1/4: the plain form x ln q + (s - x) (see ratio_term_cf)
  const f32x2_t t_large = fma2(mul2(x, splat2(0.69314718055994531f)), l, fma2(x, splat2(-1.f), s));
  unpack2(t_small, ts0, ts1);
  unpack2(t_large, tl0, tl1);
  return pack2(fabsf(u0) < 0.25f ? ts0 : tl0, fabsf(u1) < 0.25f ? ts1 : tl1);
}
__device__ __forceinline__ float tf32_round(float v);
// One 32-element chunk of the ratio epilogue held by its row-owning thread: x[] holds the data on entry and the stored
// ratio values (q - qshift; ACC && STORE_U: u = q - 1 itself; ACC: rounded to nearest TF32) on exit; returns the chunk's
// objective partial.  The arithmetic mode is a TEMPLATE parameter on purpose: a run-time branch inside the unrolled loop
// keeps the compiler from interleaving the 16 independent pairs, and their ~130-cycle dependent chains then run one
// after the other (measured: the fused k = 256 kernel 9.2 -> 12.2 ms).
template <bool ACC, bool STORE_U>
__device__ __forceinline__ float ratio_chunk32(float x[32], const uint32_t v[32], float qshift) {
  f32x2_t part2 = splat2(0.f);
  const f32x2_t nshift2 = splat2(-qshift);
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    const f32x2_t x2 = pack2(x[j], x[j + 1]);
    const f32x2_t s2 = pack2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
    f32x2_t q2, out2;
    if (ACC) {
      f32x2_t u2;
      part2 = add2(part2, ratio_pair_cf<false>(x2, s2, q2, u2));
      out2 = STORE_U ? u2 : add2(q2, nshift2);
    } else {
      part2 = add2(part2, ratio_pair_fast(x2, s2, q2));
      out2 = add2(q2, nshift2);
    }
    unpack2(out2, x[j], x[j + 1]);
  }
  if (ACC) {
#pragma unroll
    for (int j = 0; j < 32; j++) x[j] = tf32_round(x[j]);
  }
  float p0, p1;
  unpack2(part2, p0, p1);
  return p0 + p1;
}
// run-time mode -> template instance (one warp-uniform branch per chunk, outside the loop)
__device__ __forceinline__ float ratio_chunk32_dispatch(float x[32], const uint32_t v[32], float qshift, int accurate) {
  if (accurate) {
    if (qshift == 1.f) return ratio_chunk32<true, true>(x, v, qshift);
    return ratio_chunk32<true, false>(x, v, qshift);
  }
  return ratio_chunk32<false, false>(x, v, qshift);
}
// Round to nearest TF32 (ties away from zero, the tensor core's own cvt.rna.tf32.f32): add half a TF32 ulp to the
// magnitude and drop the 13 low bits.  ptxas expands cvt.rna.tf32.f32 into the same two integer instructions PLUS an
// infinity test (FSETP + predicated IADD + LOP3); Inf and NaN survive this form as well (Inf + 0x1000 masks back to
// Inf), so the test is dropped: 2 instructions per element in epilogues that round 32 elements per chunk.
__device__ __forceinline__ float tf32_round(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
}

// Shared-memory matrix descriptor (sm_100 UMMA).
//   K-major : SWIZZLE_128B (type 2, 16-byte swizzle atoms): rows of 128 B (32 floats of K),
//             8-row groups SBO = 1024 B apart.
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (type 1, 32-byte swizzle atoms,
//             TMA mode 128B_ATOM_32B): rows of 128 B (32 floats of M/N), one row per K index;
//             4-row K groups SBO = 512 B apart, 32-wide M/N groups LBO bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
// The same descriptor as two 32-bit halves: the high word is a loop constant, the low word advances with
// the stage and the K step by plain 32-bit adds (addresses stay below 2^18, so the 14-bit field never carries).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint64_t desc_pack(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }


// ---------------------------------------------------------------------------------------------
// tensor maps (host)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

// 2D fp32 tensor map: inner (contiguous) extent `inner`, `outer` rows `ld` elements apart; box = 32 x box_rows.
inline int make_map(CUtensorMap *map, const void *base, int64_t inner, int64_t outer, int64_t ld, int box_rows, bool mn_major) {
  EncodeTiledFn enc = get_encode();
  KL_CHECK(enc != nullptr, KLNMF_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
  KL_CHECK(((uintptr_t)base % 16) == 0 && (ld * 4) % 16 == 0, KLNMF_EINVAL,
           "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ld=%lld)", (long long)ld);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMapSwizzle swz = mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (mn_major) {
    const char *o = getenv("KLNMF_TC_MN_TMA");
    if (o) swz = (CUtensorMapSwizzle)atoi(o);
  }
  static const int promo = getenv("KLNMF_TC_PROMO") ? atoi(getenv("KLNMF_TC_PROMO")) : (int)CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz,
                   (CUtensorMapL2promotion)promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KL_CHECK(r == CUDA_SUCCESS, KLNMF_ECUDA, "cuTensorMapEncodeTiled failed with %d (inner=%lld outer=%lld ld=%lld)", (int)r,
           (long long)inner, (long long)outer, (long long)ld);
  return KLNMF_OK;
}


}  // namespace
}  // namespace klnmf
