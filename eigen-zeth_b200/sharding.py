"""Multi-GPU composition of the proving path: one process per GPU, torch.distributed for plumbing.

The MSM is a sum, so rank g owns the point range [g*N/G, (g+1)*N/G); every rank
reduces its range to ONE affine point with the full single-GPU pipeline and the only exchange is
an all-gather of 64 B (G1) / 128 B (G2) per rank followed by G-1 point additions.  No data-path
collective is needed, which is why bench.py reports weak scaling.
The point arithmetic itself always runs in libzkb.so on the GPU; `combine` below takes the callable
that sums points so that the host-side logic (ranges, gather, ordering) can be exercised with gloo
on a machine without GPUs (tests/test_sharding_gloo.py).
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


def groth16_prove_sharded(ctx, pk, witness, r: bytes, s: bytes, world: int):
    """One proof with every MSM sharded by point range over `world` ranks (each rank's `pk` holds
    its ranges, zkb_pk_synthetic_shard).  Every rank returns the same proof bytes."""
    import torch
    partial, timings = ctx.groth16_prove_partial(pk, witness)
    if world == 1:
        parts = [partial]
    else:
        dev = torch.device("cuda", ctx.device)
        t = torch.frombuffer(bytearray(partial), dtype=torch.uint8).to(dev)
        parts = [bytes(row.cpu().numpy()) for row in gather_partials(t, world)]
    sums = combine_partial_sums(parts, lambda pts, g2: ctx.point_sum(pts, g2=g2))
    return ctx.groth16_assemble(pk, sums, r, s), timings


class FourStepProver:
    """Groth16 prove over `world` ranks with the witness map distributed too: the seven NTTs run as
    four-step transforms (one NCCL all-to-all each, dist_ntt.py), every MSM is sharded by points,
    and h never leaves the layout the last transform produced it in (the key's h_query shard is
    stored in that layout).  One instance per rank; `pk` must come from
    ctx.pk_synthetic_shard(log_n, seed, rank, world, h_low_bits=prover.h_low_bits)."""

    def __init__(self, ctx, log_n: int, world: int, rank: int):
        import torch
        from .dist_ntt import FourStepNtt, COSET_GEN, R
        self.ctx, self.log_n, self.world, self.rank = ctx, log_n, world, rank
        self.ntt = FourStepNtt(ctx, log_n, world, rank)
        self.d0 = log_n // 2                 # layout of the evaluation vectors
        self.d1 = log_n - self.d0            # layout after one transform = layout of h
        n = 1 << log_n
        zinv = pow((pow(COSET_GEN, n, R) - 1) % R, R - 2, R)
        self.t_coset = self.ntt.index_power_table(self.d1, COSET_GEN)
        self.t_uncoset = self.ntt.index_power_table(self.d1, pow(COSET_GEN, R - 2, R), zinv)
        self.bufs = [torch.empty(self.ntt.n_loc * 32, dtype=torch.uint8, device=self.ntt.device) for _ in range(3)]

    @property
    def h_low_bits(self) -> int:
        return self.d1

    def _stage_witness(self, witness):
        """With several ranks and a (pinned) CPU tensor as witness, every rank uploads only its 1/N
        slice over PCIe and the slices are all-gathered over NVLink; returns a pointer the library
        accepts (host or device)."""
        import torch
        if isinstance(witness, torch.Tensor):
            if self.world == 1 or witness.is_cuda:
                self._keep = witness
                return witness.data_ptr()
            import torch.distributed as dist
            nbytes = witness.numel() * witness.element_size()
            flat = witness.view(torch.uint8).reshape(-1)
            per = nbytes // self.world
            if not hasattr(self, "_wfull") or self._wfull.numel() != nbytes:
                self._wfull = torch.empty(nbytes, dtype=torch.uint8, device=self.ntt.device)
                self._wown = torch.empty(per, dtype=torch.uint8, device=self.ntt.device)
                torch.cuda.current_stream(self.ntt.device).synchronize()
            def stage():
                self._wown.copy_(flat[self.rank * per:(self.rank + 1) * per], non_blocking=True)
                dist.all_gather_into_tensor(self._wfull, self._wown)
            self.ntt.run_on_torch_stream(stage)
            return self._wfull.data_ptr()
        return witness

    def witness_map(self, pk, witness):
        """-> torch tensor with this rank's h coefficients (canonical) in layout d1."""
        ctx, nt, L = self.ctx, self.ntt, self.log_n
        ctx.groth16_witness_layout_dev(pk, self._stage_witness(witness), self.d0, self.world, self.rank,
                                       [b.data_ptr() for b in self.bufs])
        for buf in self.bufs:
            nt.transform(buf, self.d0, inverse=True, post=self.t_coset)      # coefficients * g^k
        for buf in self.bufs:
            nt.transform(buf, self.d1, inverse=False)                        # evaluations on the coset
        a, b, c = self.bufs
        ctx.fr_mul_sub_dev(a.data_ptr(), b.data_ptr(), c.data_ptr(), nt.n_loc)
        nt.transform(a, self.d0, inverse=True, post=self.t_uncoset)          # h * 1  (g^-k / Z folded)
        ctx.fr_convert_dev(a.data_ptr(), nt.n_loc, False)
        return a

    def prove(self, pk, witness, r: bytes, s: bytes, timings: dict | None = None) -> bytes:
        """`timings` (optional) receives wall-clock seconds per stage; asking for it inserts stream
        synchronisations between the stages."""
        import time
        import torch

        def lap(name, t0):
            if timings is not None:
                self.ctx.sync()
                timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
            return time.perf_counter()

        t0 = time.perf_counter()
        h = self.witness_map(pk, witness)
        t0 = lap("witness_map", t0)
        partial = self.ctx.groth16_msm_dev(pk, h.data_ptr())
        t0 = lap("msm", t0)
        if self.world == 1:
            parts = [partial]
        else:
            t = torch.frombuffer(bytearray(partial), dtype=torch.uint8).to(self.ntt.device)
            parts = [bytes(row.cpu().numpy()) for row in gather_partials(t, self.world)]
        sums = combine_partial_sums(parts, lambda pts, g2: self.ctx.point_sum(pts, g2=g2))
        t0 = lap("gather_combine", t0)
        proof = self.ctx.groth16_assemble(pk, sums, r, s)
        lap("assemble", t0)
        return proof
