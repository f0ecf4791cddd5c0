// comm.cu — communicator of the multi-GPU proving path (see comm.h for the two transports).
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only: the library itself is dlopen'ed in nccl_api()
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "context.h"

namespace zkb {

// ---------------------------------------------------------------- NCCL, loaded on demand
struct NcclApi {
  void* so = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
  std::string err;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    // a libnccl already mapped into the process (PyTorch ships its own) must be the one we use:
    // two NCCL copies in one process cannot share communicators, and its SONAME is the same
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) {
      const char* e = dlerror();
      api.err = std::string("cannot load libnccl.so.2: ") + (e ? e : "?");
      return;
    }
    api.so = so;
#define ZKB_NCCL_SYM(field, name)                                        \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(so, name));    \
  if (!api.field) api.err = std::string("libnccl lacks ") + name;
    ZKB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    ZKB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    ZKB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    ZKB_NCCL_SYM(AllReduce, "ncclAllReduce")
    ZKB_NCCL_SYM(AllGather, "ncclAllGather")
    ZKB_NCCL_SYM(Send, "ncclSend")
    ZKB_NCCL_SYM(Recv, "ncclRecv")
    ZKB_NCCL_SYM(GroupStart, "ncclGroupStart")
    ZKB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    ZKB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    ZKB_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef ZKB_NCCL_SYM
  });
  return api.err.empty() ? &api : nullptr;
}

static int nccl_fail(zkb_ctx* ctx, ncclResult_t r, const char* what) {
  NcclApi* a = nccl_api();
  return set_error(ctx, ZKB_ECOMM, std::string("NCCL error: ") + (a ? a->GetErrorString(r) : "?") + " in " + what);
}
#define ZKB_NCCL(ctx, expr)                                          \
  do {                                                               \
    ncclResult_t _r = (expr);                                        \
    if (_r != ncclSuccess) return nccl_fail((ctx), _r, #expr);       \
  } while (0)

// ---------------------------------------------------------------- LOCAL transport
constexpr int kEvRing = 4;

struct LocalGroup {
  int world = 0;
  zkb_ctx* ctx[kMaxRanks] = {};
  cudaEvent_t ev[kMaxRanks][kEvRing] = {};
  void* pub[kMaxRanks][2] = {};       // pointers published for the collective in flight
  // reusable host barrier
  std::mutex mu;
  std::condition_variable cv;
  int waiting = 0;
  uint64_t generation = 0;
  bool failed = false;
  int alive = 0;                      // communicators still attached
};

// all ranks' host threads meet here; false = a peer failed or did not arrive within the timeout
static bool host_barrier(LocalGroup* lg) {
  std::unique_lock<std::mutex> lk(lg->mu);
  if (lg->failed) return false;
  const uint64_t gen = lg->generation;
  if (++lg->waiting == lg->world) {
    lg->waiting = 0;
    lg->generation++;
    lg->cv.notify_all();
    return true;
  }
  const bool ok = lg->cv.wait_for(lk, std::chrono::seconds(120), [&] { return lg->generation != gen || lg->failed; });
  if (!ok || lg->failed) {
    lg->failed = true;
    lg->cv.notify_all();
    return false;
  }
  return true;
}

// ---------------------------------------------------------------- collectives
int comm_barrier(zkb_ctx* ctx) {
  Comm* c = ctx->comm;
  if (!c || c->world == 1) return ZKB_OK;
  c->n_barriers++;
  if (c->mode == 1) {
    NcclApi* a = nccl_api();
    ZKB_NCCL(ctx, a->AllReduce(c->small, (char*)c->small + 256, 1, ncclInt32, ncclSum, (ncclComm_t)c->nccl, ctx->stream));
    return ZKB_OK;
  }
  LocalGroup* lg = c->lg;
  const int slot = int(c->seq++ % kEvRing);
  ZKB_CUDA(ctx, cudaEventRecord(lg->ev[c->rank][slot], ctx->stream));
  if (!host_barrier(lg)) return set_error(ctx, ZKB_ECOMM, "a peer rank failed or did not reach the barrier");
  for (int r = 0; r < c->world; r++)
    if (r != c->rank) ZKB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, lg->ev[r][slot], 0));
  return ZKB_OK;
}

int comm_allgather(zkb_ctx* ctx, const void* d_send, void* d_recv, size_t bytes) {
  Comm* c = ctx->comm;
  if (!c || c->world == 1) {
    if (d_recv != d_send) ZKB_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return ZKB_OK;
  }
  if (c->mode == 1) {
    NcclApi* a = nccl_api();
    ZKB_NCCL(ctx, a->AllGather(d_send, d_recv, bytes, ncclUint8, (ncclComm_t)c->nccl, ctx->stream));
    return ZKB_OK;
  }
  // LOCAL: publish the send pointer, meet, pull every peer's block over NVLink, meet again so that
  // nobody overwrites a send buffer a peer is still reading
  LocalGroup* lg = c->lg;
  lg->pub[c->rank][0] = const_cast<void*>(d_send);
  int rc = comm_barrier(ctx);
  if (rc) return rc;
  for (int r = 0; r < c->world; r++) {
    char* dst = (char*)d_recv + size_t(r) * bytes;
    if (r == c->rank) {
      if (dst != d_send) ZKB_CUDA(ctx, cudaMemcpyAsync(dst, d_send, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
      ZKB_CUDA(ctx, cudaMemcpyPeerAsync(dst, ctx->device, lg->pub[r][0], lg->ctx[r]->device, bytes, ctx->stream));
    }
  }
  return comm_barrier(ctx);
}

static void close_peer_windows(Comm* c) {
  if (c->mode != 1 || !c->p2p) return;
  for (int k = 0; k < 2; k++)
    for (int r = 0; r < c->world; r++)
      if (r != c->rank && c->peer_win[k][r]) {
        cudaIpcCloseMemHandle(c->peer_win[k][r]);
        c->peer_win[k][r] = nullptr;
      }
}

struct WinInfo {                 // what every rank tells the others about its two windows
  cudaIpcMemHandle_t h[2];
  uint64_t pid;
  uint64_t ptr[2];
  uint64_t ok;                   // cudaIpcGetMemHandle worked
};

int comm_ensure_windows(zkb_ctx* ctx, size_t bytes) {
  Comm* c = ctx->comm;
  if (!c) return set_error(ctx, ZKB_EINVAL, "no communicator on this context (zkb_comm_init first)");
  if (bytes <= c->win_bytes) return ZKB_OK;
  // the old windows may still be in use by a peer's in-flight kernel: drain everybody first
  int rc = comm_barrier(ctx);
  if (rc) return rc;
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  close_peer_windows(c);
  if (c->world > 1 && c->mode == 0 && !host_barrier(c->lg)) return set_error(ctx, ZKB_ECOMM, "peer failed");
  for (int k = 0; k < 2; k++) {
    if (c->win[k]) cudaFree(c->win[k]);
    c->win[k] = nullptr;
  }
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr;
  c->win_bytes = 0;
  for (int k = 0; k < 2; k++) ZKB_CUDA(ctx, cudaMalloc(&c->win[k], bytes));
  c->parity = 0;
  for (int k = 0; k < 2; k++) c->peer_win[k][c->rank] = c->win[k];
  if (c->world == 1) {
    c->p2p = true;
  } else if (c->mode == 0) {
    LocalGroup* lg = c->lg;
    lg->pub[c->rank][0] = c->win[0];
    lg->pub[c->rank][1] = c->win[1];
    if (!host_barrier(lg)) return set_error(ctx, ZKB_ECOMM, "peer failed");
    for (int r = 0; r < c->world; r++)
      for (int k = 0; k < 2; k++) c->peer_win[k][r] = lg->pub[r][k];
    if (!host_barrier(lg)) return set_error(ctx, ZKB_ECOMM, "peer failed");   // pub[] is free again
    c->p2p = true;
  } else {
    // NCCL transport: all-gather the IPC handles through the communicator, then map the peers
    NcclApi* a = nccl_api();
    WinInfo mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.pid = uint64_t(getpid());
    mine.ok = 1;
    for (int k = 0; k < 2; k++) {
      mine.ptr[k] = uint64_t(reinterpret_cast<uintptr_t>(c->win[k]));
      if (cudaIpcGetMemHandle(&mine.h[k], c->win[k]) != cudaSuccess) {
        cudaGetLastError();
        mine.ok = 0;
      }
    }
    static_assert(sizeof(WinInfo) * kMaxRanks + 1024 <= 65536, "scratch too small");
    char* d_mine = (char*)c->small + 1024;
    char* d_all = d_mine + sizeof(WinInfo);
    WinInfo all[kMaxRanks];
    ZKB_CUDA(ctx, cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    ZKB_NCCL(ctx, a->AllGather(d_mine, d_all, sizeof(WinInfo), ncclUint8, (ncclComm_t)c->nccl, ctx->stream));
    ZKB_CUDA(ctx, cudaMemcpyAsync(all, d_all, sizeof(WinInfo) * c->world, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bool ok = true;
    for (int r = 0; r < c->world; r++) ok = ok && all[r].ok;
    for (int r = 0; ok && r < c->world; r++) {
      if (r == c->rank) continue;
      for (int k = 0; k < 2; k++) {
        if (all[r].pid == mine.pid) {            // ranks as threads of one process: plain peer pointers
          c->peer_win[k][r] = reinterpret_cast<void*>(uintptr_t(all[r].ptr[k]));
        } else if (cudaIpcOpenMemHandle(&c->peer_win[k][r], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          c->peer_win[k][r] = nullptr;
          ok = false;
        }
      }
    }
    // every rank must take the same decision: p2p only if it worked everywhere
    int32_t flag = ok ? 0 : 1, total = 0;
    ZKB_CUDA(ctx, cudaMemcpyAsync(c->small, &flag, 4, cudaMemcpyHostToDevice, ctx->stream));
    ZKB_NCCL(ctx, a->AllReduce(c->small, (char*)c->small + 256, 1, ncclInt32, ncclSum, (ncclComm_t)c->nccl, ctx->stream));
    ZKB_CUDA(ctx, cudaMemcpyAsync(&total, (char*)c->small + 256, 4, cudaMemcpyDeviceToHost, ctx->stream));
    ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ZKB_CUDA(ctx, cudaMemsetAsync(c->small, 0, 512, ctx->stream));
    c->p2p = total == 0 && !ctx->comm_force_staging;
    if (!c->p2p) {
      c->p2p = true;           // close_peer_windows() only acts on p2p communicators
      close_peer_windows(c);
      c->p2p = false;
      ZKB_CUDA(ctx, cudaMalloc(&c->stage, bytes));
    }
  }
  c->win_bytes = bytes;
  return ZKB_OK;
}

void* comm_send_slot(zkb_ctx* ctx, int dst, size_t block_bytes) {
  Comm* c = ctx->comm;
  if (c->p2p || dst == c->rank) return (char*)c->peer_win[c->parity][dst] + size_t(c->rank) * block_bytes;
  return (char*)c->stage + size_t(dst) * block_bytes;
}

int comm_exchange(zkb_ctx* ctx, size_t block_bytes, void** my_window) {
  Comm* c = ctx->comm;
  c->n_exchanges++;
  void* w = c->win[c->parity];
  if (c->world > 1) {
    if (c->p2p) {
      // every block already sits in its destination: the producing kernels stored through NVLink
      int rc = comm_barrier(ctx);
      if (rc) return rc;
    } else {
      NcclApi* a = nccl_api();
      ZKB_NCCL(ctx, a->GroupStart());
      for (int r = 0; r < c->world; r++) {
        if (r == c->rank) continue;
        ZKB_NCCL(ctx, a->Send((char*)c->stage + size_t(r) * block_bytes, block_bytes, ncclUint8, r, (ncclComm_t)c->nccl, ctx->stream));
        ZKB_NCCL(ctx, a->Recv((char*)w + size_t(r) * block_bytes, block_bytes, ncclUint8, r, (ncclComm_t)c->nccl, ctx->stream));
      }
      ZKB_NCCL(ctx, a->GroupEnd());
    }
  }
  c->parity ^= 1;
  *my_window = w;
  return ZKB_OK;
}

void comm_destroy(zkb_ctx* ctx) {
  Comm* c = ctx->comm;
  if (!c) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  close_peer_windows(c);
  for (int k = 0; k < 2; k++)
    if (c->win[k]) cudaFree(c->win[k]);
  if (c->stage) cudaFree(c->stage);
  if (c->small) cudaFree(c->small);
  if (c->mode == 1 && c->nccl) {
    NcclApi* a = nccl_api();
    if (a) a->CommDestroy((ncclComm_t)c->nccl);
  }
  if (c->mode == 0 && c->lg) {
    LocalGroup* lg = c->lg;
    bool last;
    {
      std::lock_guard<std::mutex> lk(lg->mu);
      for (int k = 0; k < kEvRing; k++)
        if (lg->ev[c->rank][k]) cudaEventDestroy(lg->ev[c->rank][k]);
      last = --lg->alive == 0;
    }
    if (last) delete lg;
  }
  delete c;
  ctx->comm = nullptr;
}

static int comm_common_init(zkb_ctx* ctx, Comm* c) {
  ZKB_CUDA(ctx, cudaSetDevice(ctx->device));
  ZKB_CUDA(ctx, cudaMalloc(&c->small, 65536));
  ZKB_CUDA(ctx, cudaMemsetAsync(c->small, 0, 65536, ctx->stream));
  ZKB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKB_OK;
}

}  // namespace zkb

using namespace zkb;

extern "C" {

int zkb_comm_unique_id(uint8_t id_out[ZKB_COMM_ID_BYTES]) {
  if (!id_out) return ZKB_EINVAL;
  NcclApi* a = nccl_api();
  if (!a) return ZKB_ECOMM;
  static_assert(sizeof(ncclUniqueId) == ZKB_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != ncclSuccess) return ZKB_ECOMM;
  std::memcpy(id_out, &id, sizeof(id));
  return ZKB_OK;
}

int zkb_comm_init(zkb_ctx* ctx, const uint8_t id[ZKB_COMM_ID_BYTES], int rank, int world) {
  if (!ctx) return ZKB_EINVAL;
  BusyGuard guard(ctx);
  if (!guard.ok) return ZKB_EBUSY;
  if (!id || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || (world & (world - 1)))
    return set_error(ctx, ZKB_EINVAL, "world must be a power of two <= 16 and 0 <= rank < world");
  if (ctx->comm) return set_error(ctx, ZKB_EINVAL, "context already has a communicator");
  NcclApi* a = nccl_api();
  if (!a) return set_error(ctx, ZKB_ECOMM, "NCCL is not available (libnccl.so.2 could not be loaded)");
  Comm* c = new (std::nothrow) Comm();
  if (!c) return set_error(ctx, ZKB_ENOMEM, "out of host memory");
  c->rank = rank;
  c->world = world;
  c->mode = 1;
  ctx->comm = c;
  int rc = comm_common_init(ctx, c);
  if (!rc) {
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    ncclResult_t r = a->CommInitRank(&comm, world, uid, rank);
    if (r != ncclSuccess) rc = nccl_fail(ctx, r, "ncclCommInitRank");
    c->nccl = comm;
  }
  if (rc) {
    comm_destroy(ctx);
    return rc;
  }
  return ZKB_OK;
}

int zkb_comm_init_local(zkb_ctx* const* ctxs, int world) {
  if (!ctxs || world < 1 || world > kMaxRanks || (world & (world - 1))) return ZKB_EINVAL;
  for (int r = 0; r < world; r++)
    if (!ctxs[r] || ctxs[r]->comm) return ZKB_EINVAL;
  LocalGroup* lg = new (std::nothrow) LocalGroup();
  if (!lg) return ZKB_ENOMEM;
  lg->world = world;
  lg->alive = world;
  int rc = ZKB_OK;
  for (int r = 0; r < world && !rc; r++) {
    zkb_ctx* ctx = ctxs[r];
    lg->ctx[r] = ctx;
    Comm* c = new (std::nothrow) Comm();
    if (!c) {
      rc = ZKB_ENOMEM;
      break;
    }
    c->rank = r;
    c->world = world;
    c->mode = 0;
    c->lg = lg;
    ctx->comm = c;
    rc = comm_common_init(ctx, c);
    for (int k = 0; k < kEvRing && !rc; k++)
      if (cudaEventCreateWithFlags(&lg->ev[r][k], cudaEventDisableTiming) != cudaSuccess) rc = ZKB_ECUDA;
    // NVLink peer access to every other GPU of the group (already-enabled is fine)
    for (int q = 0; q < world && !rc; q++) {
      if (ctxs[q]->device == ctx->device) continue;
      int can = 0;
      cudaDeviceCanAccessPeer(&can, ctx->device, ctxs[q]->device);
      if (!can) {
        rc = set_error(ctx, ZKB_ECOMM, "GPUs of the group cannot access each other's memory");
        break;
      }
      cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
      cudaGetLastError();
    }
  }
  if (rc) {
    for (int r = 0; r < world; r++)
      if (ctxs[r]->comm) comm_destroy(ctxs[r]);
    return rc;
  }
  return ZKB_OK;
}

int zkb_comm_rank(const zkb_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->rank : 0; }
int zkb_comm_world(const zkb_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->world : 1; }
/* 0 none, 1 peer-memory windows (NVLink stores from the NTT epilogue), 2 NCCL send/recv staging */
int zkb_comm_transport(const zkb_ctx* ctx) {
  if (!ctx || !ctx->comm) return 0;
  return ctx->comm->win_bytes == 0 ? 0 : (ctx->comm->p2p ? 1 : 2);
}
void zkb_comm_destroy(zkb_ctx* ctx) {
  if (ctx) comm_destroy(ctx);
}

}  // extern "C"
