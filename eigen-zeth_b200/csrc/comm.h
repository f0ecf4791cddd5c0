// comm.h — the multi-GPU communicator behind zkb_comm_* / zkb_*_dist (include/zkb.h).
//
// The reference has no distributed backend: one operator talks to one prover
// address (reference src/config/env.rs:21, scripts/start-l2.sh:3).  Everything here is new work
// behind that single address: one zkb_ctx per GPU, joined into a communicator.
//
// Two transports behind the same collective calls:
//   LOCAL  one process, one host thread per GPU (the prover service): ranks share a LocalGroup,
//          receive windows are plain peer pointers (cudaDeviceEnablePeerAccess), stream barriers are
//          CUDA events exchanged behind a host-side barrier.  No NCCL.
//   NCCL   one process per GPU (torchrun / bench.py): rendezvous and small collectives through NCCL
//          (loaded with dlopen at init, so single-GPU users never need it); receive windows are
//          mapped into every peer with CUDA IPC, so the NTT epilogue stores straight into remote HBM
//          over NVLink and the "all-to-all" degenerates to a barrier.  If IPC mapping is refused,
//          the windows become local staging and the exchange is a grouped ncclSend / ncclRecv.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

struct zkb_ctx;

namespace zkb {

constexpr int kMaxRanks = 16;

struct LocalGroup;

struct Comm {
  int rank = 0, world = 1;
  int mode = 0;                       // 0 LOCAL, 1 NCCL
  LocalGroup* lg = nullptr;
  void* nccl = nullptr;               // ncclComm_t
  // receive windows of the fused NTT exchange, double buffered (parity flips per exchange)
  void* win[2] = {nullptr, nullptr};
  size_t win_bytes = 0;
  void* peer_win[2][kMaxRanks] = {};  // every rank's windows as seen from this rank (self included)
  bool p2p = false;                   // peer_win are directly addressable (NVLink peer memory)
  void* stage = nullptr;              // !p2p: local send staging of win_bytes
  uint32_t parity = 0;
  uint64_t seq = 0;                   // barrier sequence (LOCAL event ring)
  void* small = nullptr;              // 64 KiB device scratch (barrier token, handle exchange)
  // statistics
  uint64_t n_barriers = 0, n_exchanges = 0;
};

// all collective: every rank of the communicator calls them in the same order
int comm_barrier(zkb_ctx* ctx);                                            // stream-ordered
int comm_allgather(zkb_ctx* ctx, const void* d_send, void* d_recv, size_t bytes_per_rank);
// make sure both receive windows hold at least `bytes`; (re)maps them into the peers
int comm_ensure_windows(zkb_ctx* ctx, size_t bytes);
// Fused exchange, sender side: where rank `dst` wants this rank's block of `block_bytes` bytes for
// the exchange about to start (p2p: inside dst's window; otherwise inside the local staging buffer).
void* comm_send_slot(zkb_ctx* ctx, int dst, size_t block_bytes);
// Fused exchange, after the producing kernel: make every rank's blocks visible in my window and
// return it (p2p: a barrier; otherwise the grouped send/recv).  Flips the window parity.
int comm_exchange(zkb_ctx* ctx, size_t block_bytes, void** my_window);
void comm_destroy(zkb_ctx* ctx);

}  // namespace zkb
