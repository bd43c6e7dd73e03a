// Microbenchmark for DESIGN.md section 9, item 1: does TMA multicast relieve the L2 -> SM fabric on B200?
//
// Every CTA of a cluster needs the SAME 16 KB chunk stream (the situation of two CTA pairs that share an H tile in
// the ratio contraction).  Two ways to get it into all their shared memories:
//   mode 0 (unicast)   each CTA loads the whole chunk itself                     -> CS loads of the chunk cross the fabric
//   mode 1 (multicast) each CTA loads 1/CS of the chunk and multicasts its slice -> one load of the chunk crosses it
// The kernel does nothing else: a ring of NS stages, a producer thread, a consumer thread that frees the stage at once.
// Reported: bytes DELIVERED to shared memory per second (chunk bytes x iterations x CTAs / time) -- if the fabric is
// the limit, multicast delivers up to CS times more; if the notes are right that small clusters are already
// deduplicated on the way, it delivers the same.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mc_bench tools/mc_bench.cu && gpurun_out/mc_bench
//
// Standalone (not part of libklnmf, not built by __graft_entry__.build()).
// First run (profiles/r1_s4_mc_bench_first_run.log, last GPU seconds of round 1): the benchmark itself is latency-bound
// -- 16 B/clk/SM unicast where the real kernels pull ~48 -- and the multicast variant is slower (9.9 / 3.9 / 2.4 B/clk/SM
// at cluster size 2 / 4 / 8): its stage hand-back uses mbarrier.arrive.release.cluster to every peer, the same
// MEMBAR.GPU + ERRBAR cost that crippled the first k=256 cluster kernel.  Before it can answer the fabric question it
// needs relaxed remote arrives (nothing is published by the consumer), larger chunks / more stages per SM and several
// producer lanes; the protocol and the multicast bulk copies themselves work.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int CHUNK = 16384;     // bytes per stage
constexpr int NS = 8;            // stages

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_load_multicast(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "h"(mask)
      : "memory");
}

__global__ void __launch_bounds__(64, 1) mc_kernel(const char *buf, size_t buf_bytes, int iters, int cs, int multicast) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t ring = smem_u32(smem);
  const uint32_t full0 = ring + NS * CHUNK, empty0 = full0 + 8 * NS;
  const uint32_t rank = cs > 1 ? cluster_rank() : 0u;
  const uint32_t cluster_id = blockIdx.x / cs;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; s++) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, multicast ? cs : 1);     // multicast: every CTA's consumer frees the stage in every CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (cs > 1) cluster_sync_all(); else __syncthreads();

  const size_t n_chunks = buf_bytes / CHUNK;
  if (threadIdx.x == 0) {                                  // producer
    const uint32_t slice = CHUNK / cs;
    for (int it = 0; it < iters; it++) {
      const int s = it % NS;
      const uint32_t ph = (it / NS) & 1u;
      mbar_wait(empty0 + 8 * s, ph ^ 1u);
      mbar_expect_tx(full0 + 8 * s, CHUNK);
      const char *src = buf + (((size_t)cluster_id * iters + it) % n_chunks) * CHUNK;
      if (multicast)
        bulk_load_multicast(ring + s * CHUNK + rank * slice, src + rank * slice, slice, full0 + 8 * s, (uint16_t)((1u << cs) - 1u));
      else
        bulk_load(ring + s * CHUNK, src, CHUNK, full0 + 8 * s);
    }
  } else if (threadIdx.x == 32) {                          // consumer: frees the stage as soon as it is full
    for (int it = 0; it < iters; it++) {
      const int s = it % NS;
      const uint32_t ph = (it / NS) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      if (multicast) {
        for (int r = 0; r < cs; r++) mbar_arrive_remote(mapa(empty0 + 8 * s, (uint32_t)r));
      } else {
        mbar_arrive_remote(mapa(empty0 + 8 * s, rank));
      }
    }
  }
  if (cs > 1) cluster_sync_all(); else __syncthreads();   // nobody leaves while a peer may still write its shared memory
}

static float run(const char *buf, size_t buf_bytes, int ctas, int iters, int cs, int multicast) {
  const size_t smem = NS * CHUNK + 16 * NS + 1024;
  cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(mc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ctas / cs * cs), 1, 1);
  cfg.blockDim = dim3(64, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {                      // the first repetition also warms the L2
    cudaEventRecord(a);
    cudaError_t e = cudaLaunchKernelEx(&cfg, mc_kernel, buf, buf_bytes, iters, cs, multicast);
    cudaEventRecord(b);
    if (e != cudaSuccess || cudaEventSynchronize(b) != cudaSuccess) {
      printf("launch failed (cs=%d multicast=%d): %s\n", cs, multicast, cudaGetErrorString(cudaGetLastError()));
      return -1.f;
    }
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  const size_t buf_mb = argc > 2 ? (size_t)atoi(argv[2]) : 48;      // below the 126 MB L2: the stream is served from L2
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  char *buf = nullptr;
  cudaMalloc(&buf, buf_mb << 20);
  cudaMemset(buf, 1, buf_mb << 20);
  printf("%s, %d SMs, chunk %d B, %d stages, %d iterations per CTA, buffer %zu MB\n", prop.name, sms, CHUNK, NS, iters, buf_mb);
  printf("%8s %10s %10s %14s %16s\n", "cluster", "mode", "ms", "delivered TB/s", "B/clk/SM@1.965GHz");
  for (int cs : {1, 2, 4, 8}) {
    for (int mc = 0; mc < 2; mc++) {
      if (cs == 1 && mc == 1) continue;
      const int ctas = sms / cs * cs;
      const float ms = run(buf, buf_mb << 20, ctas, iters, cs, mc);
      if (ms <= 0.f) continue;
      const double bytes = (double)CHUNK * iters * ctas;
      printf("%8d %10s %10.3f %14.2f %16.1f\n", cs, mc ? "multicast" : "unicast", ms, bytes / ms / 1e9,
             bytes / ms / 1e-3 / ctas / 1.965e9);
    }
  }
  cudaFree(buf);
  return 0;
}
