"""Multi-GPU four-step Fr NTT: one process per GPU, the transpose between the two steps is an NCCL
all-to-all over NVLink (torch.distributed.all_to_all_single); everything else runs in libzkb.so.

Data layout ("layout d").  A length-n vector (n = 2^L) is indexed  idx = low + 2^d * high  with
low in [0, 2^d), high in [0, 2^(L-d)).  Rank g of G owns the lows of its block,
low = g * 2^d / G + low_loc, and ALL highs: the local array is [low_loc][high], `high` contiguous.
A transform reads layout d and writes layout L - d (the classic four-step index swap
i = i1*M + i2  ->  k = k1 + N1*k2 with M = 2^d, N1 = 2^(L-d)), so transforms chain without any
re-distribution; with L even and d = L/2 the layout is a fixed point.

Steps on rank g (M = 2^d, N1 = 2^(L-d), Mloc = M/G, N1loc = N1/G):
  (a) Mloc contiguous size-N1 transforms over `high`, output multiplied by the inter-step twiddle
      w_n^{i2*k1} (table, i2 = g*Mloc + low_loc)                 [zkb_fr_ntt_batch_dev]
  (b) pack   [Mloc][G][N1loc] -> [G][Mloc][N1loc]                [zkb_fr_permute_dev]
  (c) all-to-all, equal splits: rank g' receives every rank's k1-block g'     [NCCL]
  (d) unpack [M][N1loc] -> [N1loc][M]                            [zkb_fr_permute_dev]
  (e) N1loc contiguous size-M transforms                         [zkb_fr_ntt_batch_dev]
Coset shifts and 1/n are folded into the tables of (a) / (e).

`exchange` is injectable so that G ranks can be emulated inside one process on one GPU (tests):
it maps a list of per-rank send buffers to per-rank receive buffers.
"""
from __future__ import annotations

from typing import Optional

import torch

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
COSET_GEN = 5


def _fr(x: int) -> bytes:
    return (x % R).to_bytes(32, "little")


def local_index_map(log_n: int, d: int, world: int, rank: int) -> torch.Tensor:
    """global index of every local position of layout d, shape [2^d / world, 2^(L-d)] (int64, CPU)."""
    lows = (1 << d) // world
    low = torch.arange(rank * lows, (rank + 1) * lows, dtype=torch.int64).unsqueeze(1)
    high = torch.arange(1 << (log_n - d), dtype=torch.int64).unsqueeze(0)
    return low + (high << d)


class FourStepNtt:
    """per-rank engine; buffers are torch uint8 CUDA tensors of n/world * 32 bytes (Montgomery form)."""

    def __init__(self, ctx, log_n: int, world: int, rank: int, device: Optional[torch.device] = None):
        if world & (world - 1):
            raise ValueError("world size must be a power of two")
        if (1 << (log_n // 2)) < world:
            raise ValueError("transform too small for this many ranks")
        self.ctx, self.log_n, self.world, self.rank = ctx, log_n, world, rank
        self.device = device or torch.device("cuda", ctx.device)
        self.n_loc = (1 << log_n) // world
        self._tables = {}
        self._tmp = [torch.empty(self.n_loc * 32, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=self.device)

    # ---- tables (built once per (layout, direction, variant), resident)
    def _twiddle(self, d: int, inverse: bool, scale: int) -> torch.Tensor:
        key = ("tw", d, inverse, scale)
        if key not in self._tables:
            from . import lib
            L, G, g = self.log_n, self.world, self.rank
            mloc, n1 = (1 << d) // G, 1 << (L - d)
            t = torch.empty(self.n_loc * 32, dtype=torch.uint8, device=self.device)
            self.ctx.fr_table_dev(t.data_ptr(), lib.fr_root_of_unity(L, inverse), _fr(scale), mloc, n1, g * mloc, 0, 0)
            self._tables[key] = t
        return self._tables[key]

    def index_power_table(self, d: int, base: int, scale: int = 1) -> torch.Tensor:
        """table[pos] = scale * base^{global index of pos} for layout d (coset scaling, 1/Z folding)."""
        key = ("pw", d, base % R, scale % R)
        if key not in self._tables:
            L, G, g = self.log_n, self.world, self.rank
            lows = (1 << d) // G
            t = torch.empty(self.n_loc * 32, dtype=torch.uint8, device=self.device)
            self.ctx.fr_table_dev(t.data_ptr(), _fr(base), _fr(scale), lows, 1 << (L - d), g * lows, 1 << d, 1)
            self._tables[key] = t
        return self._tables[key]

    def run_on_torch_stream(self, fn):
        """Run `fn` (a collective / torch copy) on torch's own current stream, ordered after the work
        already queued on the library stream and before whatever the library queues next.  The
        library stream is never made torch's *current* stream: the caching allocator and NCCL's
        record_stream bookkeeping then only ever see streams torch owns."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.stream.record_event())
        out = fn()
        self.stream.wait_event(cur.record_event())
        return out

    # ---- one transform: layout d -> layout L - d
    def phase1(self, x: torch.Tensor, d: int, inverse: bool = False, pre: Optional[torch.Tensor] = None) -> torch.Tensor:
        """steps (a)-(b): returns the packed send buffer [G][Mloc][N1loc]."""
        L, G = self.log_n, self.world
        mloc, n1 = (1 << d) // G, 1 << (L - d)
        ninv = pow(1 << L, R - 2, R) if inverse else 1
        tw = self._twiddle(d, inverse, ninv)
        a, b = self._tmp
        self.ctx.fr_ntt_batch_dev(x.data_ptr(), a.data_ptr(), L - d, mloc, inverse=inverse, skip_ninv=True,
                                  pre=pre.data_ptr() if pre is not None else 0, post=tw.data_ptr())
        if G == 1:
            return a
        self.ctx.fr_permute_dev(a.data_ptr(), b.data_ptr(), mloc, G, n1 // G)
        return b

    def phase2(self, recv: torch.Tensor, x: torch.Tensor, d: int, inverse: bool = False,
               post: Optional[torch.Tensor] = None) -> torch.Tensor:
        """steps (d)-(e): recv = [G][Mloc][N1loc] as delivered by the all-to-all; result into x."""
        L, G = self.log_n, self.world
        n1loc = (1 << (L - d)) // G
        spare = self._tmp[1] if recv.data_ptr() == self._tmp[0].data_ptr() else self._tmp[0]
        self.ctx.fr_permute_dev(recv.data_ptr(), spare.data_ptr(), 1 << d, n1loc, 1)
        self.ctx.fr_ntt_batch_dev(spare.data_ptr(), x.data_ptr(), d, n1loc, inverse=inverse, skip_ninv=True,
                                  post=post.data_ptr() if post is not None else 0)
        return x

    def transform(self, x: torch.Tensor, d: int, inverse: bool = False, pre: Optional[torch.Tensor] = None,
                  post: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: local data in layout d (overwritten with the result in layout L - d).  `pre` / `post`
        are optional element-wise tables in the input / output layout.  An inverse transform carries
        1/n (folded into the inter-step twiddle table)."""
        send = self.phase1(x, d, inverse, pre)
        if self.world == 1:
            return self.phase2(send, x, d, inverse, post)
        import torch.distributed as dist
        recv = self._tmp[0]                      # phase1 left the packed data in _tmp[1]
        self.run_on_torch_stream(lambda: dist.all_to_all_single(recv, send))
        return self.phase2(recv, x, d, inverse, post)

    # ---- helpers for tests / callers that start from a replicated natural-order vector
    def scatter_from_full(self, full: torch.Tensor, d: int) -> torch.Tensor:
        """full: uint8 CUDA tensor of n*32 bytes in natural order -> this rank's layout-d block."""
        L, G, g = self.log_n, self.world, self.rank
        lows = (1 << d) // G
        v = full.view(1 << (L - d), G, lows, 32)[:, g]            # [high][low_loc]
        return v.permute(1, 0, 2).contiguous().view(-1)            # [low_loc][high]

    def gather_to_full(self, locals_by_rank, d: int) -> torch.Tensor:
        """inverse of scatter_from_full given every rank's block (tests)."""
        L, G = self.log_n, self.world
        lows = (1 << d) // G
        full = torch.empty((1 << (L - d), G, lows, 32), dtype=torch.uint8, device=self.device)
        for g, loc in enumerate(locals_by_rank):
            full[:, g] = loc.view(lows, 1 << (L - d), 32).permute(1, 0, 2)
        return full.view(-1)
