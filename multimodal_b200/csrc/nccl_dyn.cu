// NCCL bound at run time (dlopen) so that libklnmf.so has no link-time dependency: the
// library also has to load on machines without NCCL / without a GPU (symbol-export test).
// Used for the ONE exchange step of the path: the all-reduce of the k x f dictionary
// numerator plus the objective partials.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace klnmf {

namespace {

typedef struct { char internal[128]; } nccl_uid_t;
typedef void *nccl_comm_t;
typedef int (*fn_get_uid)(nccl_uid_t *);
typedef int (*fn_comm_init)(nccl_comm_t *, int, nccl_uid_t, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef const char *(*fn_errstr)(int);
typedef int (*fn_group)(void);

struct NcclApi {
  void *handle = nullptr;
  fn_get_uid get_uid = nullptr;
  fn_comm_init comm_init = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_errstr errstr = nullptr;
  fn_group group_start = nullptr, group_end = nullptr;
} g_nccl;

constexpr int kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0;

int check(int r, const char *what) {
  if (r == 0) return KLNMF_OK;
  set_error("NCCL %s failed: %s", what, g_nccl.errstr ? g_nccl.errstr(r) : "?");
  return KLNMF_ENCCL;
}

}  // namespace

int nccl_load(const char *path) {
  if (g_nccl.handle) return KLNMF_OK;
  const char *cands[] = {path, "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *c : cands) {
    if (!c) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  KL_CHECK(h != nullptr, KLNMF_ENCCL, "cannot dlopen NCCL (%s): %s", path ? path : "libnccl.so.2", dlerror());
  g_nccl.get_uid = (fn_get_uid)dlsym(h, "ncclGetUniqueId");
  g_nccl.comm_init = (fn_comm_init)dlsym(h, "ncclCommInitRank");
  g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
  g_nccl.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
  g_nccl.group_start = (fn_group)dlsym(h, "ncclGroupStart");
  g_nccl.group_end = (fn_group)dlsym(h, "ncclGroupEnd");
  KL_CHECK(g_nccl.get_uid && g_nccl.comm_init && g_nccl.allreduce && g_nccl.comm_destroy, KLNMF_ENCCL,
           "NCCL library lacks required symbols");
  g_nccl.handle = h;
  return KLNMF_OK;
}

int nccl_unique_id(void *id128) {
  KL_TRY(nccl_load(nullptr));
  nccl_uid_t id;
  KL_TRY(check(g_nccl.get_uid(&id), "ncclGetUniqueId"));
  memcpy(id128, &id, sizeof(id));
  return KLNMF_OK;
}

// A communicator that outlives the contexts it serves: creating one costs 0.2-1 s (ncclCommInitRank), a context is
// created per fit / transform call -- callers keep the communicator and attach it to each new context.
int nccl_comm_create(void **out, int device, const void *id128, int rank, int world) {
  KL_TRY(nccl_load(nullptr));
  KL_CHECK(out && world >= 1 && rank >= 0 && rank < world, KLNMF_EINVAL, "bad rank/world %d/%d", rank, world);
  nccl_uid_t id;
  memcpy(&id, id128, sizeof(id));
  KL_CUDA(cudaSetDevice(device));
  nccl_comm_t comm = nullptr;
  KL_TRY(check(g_nccl.comm_init(&comm, world, id, rank), "ncclCommInitRank"));
  *out = comm;
  return KLNMF_OK;
}

int nccl_comm_free(void *comm) {
  if (comm && g_nccl.comm_destroy) return check(g_nccl.comm_destroy((nccl_comm_t)comm), "ncclCommDestroy");
  return KLNMF_OK;
}

int nccl_comm_init(klnmf_ctx *ctx, const void *id128, int rank, int world) {
  void *comm = nullptr;
  KL_TRY(nccl_comm_create(&comm, ctx->device, id128, rank, world));
  nccl_comm_destroy(ctx);
  ctx->comm = comm;
  ctx->comm_owned = true;
  ctx->rank = rank;
  ctx->world = world;
  return KLNMF_OK;
}

int nccl_group_start() { return g_nccl.group_start ? check(g_nccl.group_start(), "ncclGroupStart") : KLNMF_OK; }
int nccl_group_end() { return g_nccl.group_end ? check(g_nccl.group_end(), "ncclGroupEnd") : KLNMF_OK; }

int nccl_allreduce_sum(klnmf_ctx *ctx, void *buf, int64_t count, int es) {
  if (ctx->world <= 1 || count <= 0) return KLNMF_OK;
  ctx->n_nccl++;
  return check(g_nccl.allreduce(buf, buf, (size_t)count, es == 8 ? kNcclFloat64 : kNcclFloat32, kNcclSum,
                                (nccl_comm_t)ctx->comm, ctx->stream),
               "ncclAllReduce");
}

int nccl_allreduce_sum_on(klnmf_ctx *ctx, void *buf, int64_t count, int es, cudaStream_t stream) {
  if (ctx->world <= 1 || count <= 0) return KLNMF_OK;
  ctx->n_nccl++;
  return check(g_nccl.allreduce(buf, buf, (size_t)count, es == 8 ? kNcclFloat64 : kNcclFloat32, kNcclSum,
                                (nccl_comm_t)ctx->comm, stream),
               "ncclAllReduce");
}

int nccl_allreduce_sum_f64(klnmf_ctx *ctx, double *buf, int64_t count) {
  return nccl_allreduce_sum(ctx, buf, count, 8);
}

void nccl_comm_destroy(klnmf_ctx *ctx) {
  if (ctx->comm && ctx->comm_owned && g_nccl.comm_destroy) g_nccl.comm_destroy((nccl_comm_t)ctx->comm);
  ctx->comm = nullptr;
  ctx->comm_owned = false;
}

}  // namespace klnmf
