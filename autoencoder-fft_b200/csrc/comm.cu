// Collectives inside the engine: one NCCL communicator per ctx (= per process / GPU), all operations queued on the ctx
// stream so they order with the kernels that produce / consume the buffers.  The reference has no multi-GPU path
// (SURVEY 2.2); the engine adds exactly two exchanges:
//   * data-parallel frames: ONE all-reduce(sum) of the fused raw gradient block per training step (aefft_net_step), and
//     one all-reduce(avg) of the kernel-space gradient block per backprop_fft iteration -- both before the non-linear clip;
//   * frequency-bin sharding: an all-to-all of row-transformed frame slabs (ncclSend/ncclRecv group) and the
//     all-reduce(sum) of the partial kernel-space gradient block.
// NCCL is bound at run time (dlopen): a process that already holds a libnccl (e.g. torch's bundled one) shares it, a plain
// C++ caller gets the system library, and libaefft.so itself loads on machines without NCCL (single-GPU use).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace aefft {

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi* api() {
  static NcclApi a;
  static bool tried = false;
  if (tried) return a.ok ? &a : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already uses, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return nullptr;
#define AEFFT_NCCL_SYM(field, name) \
  a.field = (decltype(a.field))dlsym(h, name); \
  if (!a.field) return nullptr;
  AEFFT_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  AEFFT_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  AEFFT_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  AEFFT_NCCL_SYM(AllReduce, "ncclAllReduce")
  AEFFT_NCCL_SYM(Send, "ncclSend")
  AEFFT_NCCL_SYM(Recv, "ncclRecv")
  AEFFT_NCCL_SYM(GroupStart, "ncclGroupStart")
  AEFFT_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  AEFFT_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef AEFFT_NCCL_SYM
  a.ok = true;
  return &a;
}

#define AE_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) {                                                                      \
      aefft::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, api()->GetErrorString(r__));   \
      return AEFFT_ERR_CUDA;                                                                       \
    }                                                                                              \
  } while (0)

}  // namespace

int comm_allreduce(aefft_ctx* ctx, float* dev, int64_t n, int op) {
  if (ctx->comm_world <= 1 || n <= 0) return AEFFT_OK;
  NcclApi* a = api();
  AE_ARG(a && ctx->comm);
  AE_NCCL(a->AllReduce(dev, dev, (size_t)n, ncclFloat, op == 1 ? ncclAvg : ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
  return AEFFT_OK;
}

int comm_alltoall(aefft_ctx* ctx, const float* send, float* recv, int64_t chunk) {
  NcclApi* a = api();
  AE_ARG(a && ctx->comm && ctx->comm_world > 1 && chunk > 0);
  AE_NCCL(a->GroupStart());
  for (int r = 0; r < ctx->comm_world; r++) {
    AE_NCCL(a->Send(send + (size_t)r * chunk, (size_t)chunk, ncclFloat, r, (ncclComm_t)ctx->comm, ctx->stream));
    AE_NCCL(a->Recv(recv + (size_t)r * chunk, (size_t)chunk, ncclFloat, r, (ncclComm_t)ctx->comm, ctx->stream));
  }
  AE_NCCL(a->GroupEnd());
  return AEFFT_OK;
}

// all-to-all with per-peer counts (floats) and offsets: the frame-sharded -> bin-sharded exchange of spectrum slabs
int comm_alltoallv(aefft_ctx* ctx, const float* send, const int64_t* scount, const int64_t* soff, float* recv,
                   const int64_t* rcount, const int64_t* roff) {
  NcclApi* a = api();
  AE_ARG(a && ctx->comm && ctx->comm_world > 1);
  AE_NCCL(a->GroupStart());
  for (int r = 0; r < ctx->comm_world; r++) {
    if (scount[r] > 0) AE_NCCL(a->Send(send + soff[r], (size_t)scount[r], ncclFloat, r, (ncclComm_t)ctx->comm, ctx->stream));
    if (rcount[r] > 0) AE_NCCL(a->Recv(recv + roff[r], (size_t)rcount[r], ncclFloat, r, (ncclComm_t)ctx->comm, ctx->stream));
  }
  AE_NCCL(a->GroupEnd());
  return AEFFT_OK;
}

}  // namespace aefft

using namespace aefft;

extern "C" {

int aefft_comm_unique_id(void* id128) {
  AE_ARG(id128);
  NcclApi* a = api();
  if (!a) { set_error("aefft_comm_unique_id: libnccl.so.2 not found"); return AEFFT_ERR_UNSUPPORTED; }
  static_assert(sizeof(ncclUniqueId) == AEFFT_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  AE_NCCL(a->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return AEFFT_OK;
}

int aefft_comm_init(aefft_ctx* ctx, const void* id128, int rank, int world) {
  AE_ARG(ctx && id128 && world >= 1 && rank >= 0 && rank < world && !ctx->comm);
  if (world == 1) { ctx->comm_rank = 0; ctx->comm_world = 1; return AEFFT_OK; }
  NcclApi* a = api();
  if (!a) { set_error("aefft_comm_init: libnccl.so.2 not found"); return AEFFT_ERR_UNSUPPORTED; }
  AE_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  AE_NCCL(a->CommInitRank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return AEFFT_OK;
}

int aefft_comm_destroy(aefft_ctx* ctx) {
  AE_ARG(ctx);
  if (ctx->comm) {
    AE_CUDA(cudaSetDevice(ctx->device));
    AE_CUDA(cudaStreamSynchronize(ctx->stream));
    AE_NCCL(api()->CommDestroy((ncclComm_t)ctx->comm));
  }
  ctx->comm = nullptr;
  ctx->comm_rank = 0;
  ctx->comm_world = 1;
  return AEFFT_OK;
}

int aefft_comm_rank(const aefft_ctx* ctx) { return ctx ? ctx->comm_rank : -1; }
int aefft_comm_world(const aefft_ctx* ctx) { return ctx ? ctx->comm_world : -1; }

int aefft_comm_allreduce(aefft_ctx* ctx, float* dev, int64_t n_floats, int op) {
  AE_ARG(ctx && dev && n_floats >= 0 && (op == 0 || op == 1));
  AE_CUDA(cudaSetDevice(ctx->device));
  return comm_allreduce(ctx, dev, n_floats, op);
}

}  // extern "C"
