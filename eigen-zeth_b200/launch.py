"""Rendezvous helpers for zkb_comm_init (one process per GPU).  The communicator needs one 128-byte
NCCL unique id made by rank 0 and known to every rank; how it travels is the host's business."""
from __future__ import annotations

import os
import time


def unique_id_via_file(path: str, rank: int, timeout_s: float = 120.0) -> bytes:
    """rank 0 writes the id to `path` (atomically), the others wait for it — enough for ranks that
    share a file system (a node, or a cluster scratch directory); no torch involved"""
    from . import lib
    if rank == 0:
        uid = lib.comm_unique_id()
        tmp = "%s.%d.tmp" % (path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid
    t0 = time.time()
    while time.time() - t0 < timeout_s:
        try:
            with open(path, "rb") as f:
                uid = f.read()
            if len(uid) == lib.COMM_ID_BYTES:
                return uid
        except FileNotFoundError:
            pass
        time.sleep(0.01)
    raise TimeoutError("no unique id at %s after %.0f s" % (path, timeout_s))


def unique_id_via_torch(device=None) -> bytes:
    """under torchrun, with torch.distributed already initialised: broadcast from rank 0"""
    import torch
    import torch.distributed as dist
    from . import lib
    uid = lib.comm_unique_id() if dist.get_rank() == 0 else bytes(lib.COMM_ID_BYTES)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.frombuffer(bytearray(uid), dtype=torch.uint8).to(dev)
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy())
