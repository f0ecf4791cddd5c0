"""Four-step multi-GPU NTT (eigen-zeth_b200/dist_ntt.py): G ranks emulated inside one process on
one GPU (the all-to-all is replaced by the equivalent host-side block exchange), checked against the
single-GPU NTT, which is itself checked against the oracles.  The real NCCL path is exercised by
tests/test_multigpu.py when more than one GPU is visible."""
import pytest
import torch

import oracle_a as o
import oracle_b as ob

pytestmark = pytest.mark.gpu


def _emulated_exchange(sends, world):
    """what all_to_all_single does: rank r receives chunk r of every rank's send buffer, in rank order."""
    chunks = [s.view(world, -1) for s in sends]
    return [torch.cat([chunks[src][dst] for src in range(world)]).contiguous() for dst in range(world)]


def _to_dev(ctx, data: bytes):
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    ctx.fr_convert_dev(t.data_ptr(), len(data) // 32, True)
    return t


def _run(ctx, engines, blocks, d, inverse, pre=None, post=None):
    world = len(engines)
    sends = []
    for g, e in enumerate(engines):
        send = e.phase1(blocks[g], d, inverse, pre[g] if pre else None)
        ctx.sync()                       # the library stream does not synchronise with torch's
        sends.append(send.clone())
    torch.cuda.synchronize()
    recvs = _emulated_exchange(sends, world) if world > 1 else sends
    torch.cuda.synchronize()
    for g, e in enumerate(engines):
        e.phase2(recvs[g], blocks[g], d, inverse, post[g] if post else None)
        ctx.sync()
    return blocks


@pytest.mark.parametrize("log_n,world,d", [(6, 1, 3), (8, 2, 4), (10, 4, 5), (11, 2, 5), (11, 4, 6), (14, 8, 7), (17, 4, 8)])
def test_four_step_matches_single_gpu(ctx, zkb, log_n, world, d):
    from eigen_zeth_b200.dist_ntt import FourStepNtt, local_index_map
    n = 1 << log_n
    vec = ob.random_scalars(n, 0xD157 + log_n)
    full = _to_dev(ctx, vec)
    ctx.sync()
    engines = [FourStepNtt(ctx, log_n, world, g) for g in range(world)]
    for inverse in (False, True):
        blocks = [e.scatter_from_full(full, d) for e in engines]
        # the index map documents the layout: block[pos] == full[map[pos]]
        m = local_index_map(log_n, d, world, world - 1).reshape(-1)
        got = blocks[-1].view(-1, 32).cpu()
        assert torch.equal(got, full.view(-1, 32).cpu()[m])
        blocks = _run(ctx, engines, blocks, d, inverse)
        out = engines[0].gather_to_full(blocks, log_n - d)
        ctx.fr_convert_dev(out.data_ptr(), n, False)
        ctx.sync()
        assert bytes(out.cpu().numpy()) == ctx.fr_ntt(vec, inverse=inverse)
    if log_n <= 11:
        assert ctx.fr_ntt(vec) == ob.fr_ntt(vec)


def test_four_step_coset_chain_like_groth16(ctx, zkb):
    """iNTT -> coset NTT -> coset iNTT with the shifts folded into index-power tables, chained
    without re-distribution (layout d -> L-d -> d -> L-d)."""
    from eigen_zeth_b200.dist_ntt import FourStepNtt, COSET_GEN
    log_n, world, d = 12, 4, 6
    n = 1 << log_n
    vec = ob.random_scalars(n, 77)
    full = _to_dev(ctx, vec)
    ctx.sync()
    engines = [FourStepNtt(ctx, log_n, world, g) for g in range(world)]
    blocks = [e.scatter_from_full(full, d) for e in engines]
    d1 = log_n - d
    gi = pow(COSET_GEN, o.R - 2, o.R)
    post1 = [e.index_power_table(d1, COSET_GEN) for e in engines]          # coefficients * g^k
    post3 = [e.index_power_table(d1, gi) for e in engines]                 # ... * g^-k
    blocks = _run(ctx, engines, blocks, d, True, post=post1)               # -> layout d1
    blocks = _run(ctx, engines, blocks, d1, False)                         # -> layout d
    mid = engines[0].gather_to_full(blocks, d).clone()
    ctx.fr_convert_dev(mid.data_ptr(), n, False)
    ctx.sync()
    want_mid = ctx.fr_ntt(ctx.fr_ntt(vec, inverse=True), coset=True)
    assert bytes(mid.cpu().numpy()) == want_mid
    blocks = _run(ctx, engines, blocks, d, True, post=post3)               # -> layout d1
    out = engines[0].gather_to_full(blocks, d1)
    ctx.fr_convert_dev(out.data_ptr(), n, False)
    ctx.sync()
    assert bytes(out.cpu().numpy()) == ctx.fr_ntt(want_mid, inverse=True, coset=True)


def test_batched_ntt_and_permute(ctx):
    log_size, batch = 9, 5
    n = 1 << log_size
    vec = ob.random_scalars(n * batch, 5)
    t = _to_dev(ctx, vec)
    out = torch.empty_like(t)
    ctx.fr_ntt_batch_dev(t.data_ptr(), out.data_ptr(), log_size, batch)
    ctx.fr_convert_dev(out.data_ptr(), n * batch, False)
    ctx.sync()
    got = bytes(out.cpu().numpy())
    for b in range(batch):
        assert got[b * n * 32:(b + 1) * n * 32] == ob.fr_ntt(vec[b * n * 32:(b + 1) * n * 32])
    A, B, C = 3, 5, 4
    src = torch.arange(A * B * C * 32, dtype=torch.int32).to(torch.uint8).cuda()
    dst = torch.empty_like(src)
    ctx.fr_permute_dev(src.data_ptr(), dst.data_ptr(), A, B, C)
    ctx.sync()
    assert torch.equal(dst.view(B, A, C, 32), src.view(A, B, C, 32).permute(1, 0, 2, 3))
