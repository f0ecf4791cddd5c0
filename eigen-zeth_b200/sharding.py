"""Host-side bookkeeping of the multi-GPU proving path: point ranges, the slot layout of the partial
sums, the index map of the four-step NTT layouts.

The multi-GPU execution itself lives in libzkb.so (csrc/comm.cu, csrc/dist.cu, the distributed
prover in csrc/groth16.cu) behind zkb_comm_* / zkb_*_dist of include/zkb.h.  What is kept here is
the arithmetic-free logic those calls implement, in a form that runs with gloo on a machine without
GPUs (tests/test_sharding_gloo.py): the MSM is a sum, so rank g owns the point range
[g*N/G, (g+1)*N/G), every rank reduces its range to ONE affine point and the only
exchange is an all-gather of 64 B (G1) / 128 B (G2) per rank followed by G-1 point additions.
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, first+count) of rank `rank`; ranges tile [0, n) and differ by at most one point."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_shards(n: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(n, world, r) for r in range(world)]


def gather_partials(partial, world: int):
    """all-gather one fixed-size uint8 tensor per rank -> tensor of shape [world, len]."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return partial.reshape(1, -1)
    out = torch.empty((world, partial.numel()), dtype=torch.uint8, device=partial.device)
    dist.all_gather_into_tensor(out.reshape(-1), partial.reshape(-1))
    return out


def sharded_msm(local_msm: Callable[[int, int], "object"], point_sum: Callable[["object"], "object"], n: int,
                world: int, rank: int):
    """local_msm(first, count) -> uint8 tensor (affine partial); point_sum(tensor[world, sz]) -> result."""
    first, count = shard_range(n, world, rank)
    partial = local_msm(first, count)
    return point_sum(gather_partials(partial, world))


# slot layout of the 384-byte MSM-sum block (include/zkb.h): (offset, size, is_g2)
SLOTS = ((0, 64, False), (64, 64, False), (128, 128, True), (256, 64, False), (320, 64, False))


def combine_partial_sums(partials, point_sum: Callable[[bytes, bool], bytes]) -> bytes:
    """partials: list of 384-byte blocks (one per rank) -> slot-wise sums (384 bytes).
    point_sum(points_bytes, is_g2) adds affine points (zkb_g1_sum / zkb_g2_sum on the GPU)."""
    out = bytearray(384)
    for off, size, g2 in SLOTS:
        pts = b"".join(bytes(p[off:off + size]) for p in partials)
        out[off:off + size] = point_sum(pts, g2)
    return bytes(out)


def local_index_map(log_n: int, d: int, world: int, rank: int):
    """"Layout d" of the four-step NTT (csrc/dist.cu): a length-2^log_n vector is indexed
    idx = low + 2^d * high; rank `rank` owns the lows [rank * 2^d / world, (rank + 1) * 2^d / world)
    and all highs.  Returns the global index of every local position as rows [low_loc][high]."""
    lows = (1 << d) // world
    highs = 1 << (log_n - d)
    return [[(rank * lows + l) + (h << d) for h in range(highs)] for l in range(lows)]
