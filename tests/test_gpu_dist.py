"""The multi-GPU path inside libzkb.so (csrc/comm.cu, csrc/dist.cu, zkb_groth16_prove_dist) with the
ranks of a communicator emulated on ONE GPU: zkb_comm_init_local puts `world` contexts (one host
thread each) on the same device, so the fused exchange — the NTT epilogue storing into the
destination rank's receive window, the next prologue reading it transposed — the event barriers and
the on-device combination of the MSM partial sums all run exactly as they do across NVLink.
Everything is compared with the single-GPU results, which are themselves pinned to the oracles.
The NCCL / CUDA-IPC transport between processes is exercised by tests/test_multigpu.py."""
import pytest
import torch

import oracle_a as o
import oracle_b as ob

pytestmark = pytest.mark.gpu


def _to_dev(ctx, data: bytes):
    t = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    ctx.fr_convert_dev(t.data_ptr(), len(data) // 32, True)
    ctx.sync()
    return t


def _blocks(group, full, log_n, d):
    """every rank's block of layout d, cut out of a natural-order Montgomery vector"""
    out = []
    n_loc = (1 << log_n) // group.world
    for r, ctx in enumerate(group.ctxs):
        blk = torch.empty(n_loc * 32, dtype=torch.uint8, device="cuda")
        ctx.fr_gather_layout_dev(full.data_ptr(), blk.data_ptr(), log_n, d, group.world, r)
        ctx.sync()
        out.append(blk)
    return out


def _assemble(group, blocks, log_n, d):
    """inverse of _blocks (layout d -> natural order), canonical bytes"""
    world, lows, highs = group.world, (1 << d) // group.world, 1 << (log_n - d)
    full = torch.empty((highs, world, lows, 32), dtype=torch.uint8, device="cuda")
    for g, blk in enumerate(blocks):
        full[:, g] = blk.view(lows, highs, 32).permute(1, 0, 2)
    full = full.view(-1).contiguous()
    torch.cuda.synchronize()          # the library streams are non-blocking: they do not wait for torch's copies
    group.ctxs[0].fr_convert_dev(full.data_ptr(), 1 << log_n, False)
    group.ctxs[0].sync()
    return bytes(full.cpu().numpy())


@pytest.fixture(scope="module")
def groups(zkb):
    made = {}

    def get(world):
        if world not in made:
            made[world] = zkb.lib.DeviceGroup([0] * world)
        return made[world]

    yield get
    for g in made.values():
        g.close()


@pytest.mark.parametrize("log_n,world,d", [(6, 1, 3), (8, 2, 4), (10, 4, 5), (11, 2, 5), (11, 4, 6), (14, 8, 7),
                                           (17, 4, 8), (18, 2, 9), (20, 8, 10), (21, 16, 10)])
def test_four_step_ntt_fused_exchange(ctx, groups, log_n, world, d):
    group = groups(world)
    n = 1 << log_n
    vec = ob.random_scalars(n, 0xD157 + log_n)
    full = _to_dev(ctx, vec)
    for inverse in (False, True):
        blocks = _blocks(group, full, log_n, d)
        torch.cuda.synchronize()

        def run(r, c):
            c.fr_ntt_dist_dev(blocks[r].data_ptr(), log_n, d, inverse=inverse)
            c.sync()

        group.run(run)
        assert _assemble(group, blocks, log_n, log_n - d) == ctx.fr_ntt(vec, inverse=inverse), (log_n, world, d, inverse)
    assert group.ctxs[0].comm_transport.startswith("peer-memory")


def test_four_step_coset_chain_like_groth16(ctx, groups):
    """iNTT -> coset NTT -> coset iNTT with the shifts folded into index-power tables, chained
    without re-distribution (layout d -> L-d -> d -> L-d), as the distributed prover runs them."""
    log_n, world, d = 12, 4, 6
    group = groups(world)
    n, d1 = 1 << log_n, log_n - d
    vec = ob.random_scalars(n, 77)
    full = _to_dev(ctx, vec)
    blocks = _blocks(group, full, log_n, d)
    gi = pow(5, o.R - 2, o.R)
    torch.cuda.synchronize()
    mids = [None] * world

    def run(r, c):
        t_coset = c.fr_index_power_table_dist(log_n, d1, 5)
        t_un = c.fr_index_power_table_dist(log_n, d1, gi)
        p = blocks[r].data_ptr()
        c.fr_ntt_dist_dev(p, log_n, d, inverse=True, post=t_coset)      # coefficients * g^k, layout d1
        c.fr_ntt_dist_dev(p, log_n, d1)                                  # evaluations on the coset, layout d
        c.sync()
        mids[r] = blocks[r].clone()
        torch.cuda.synchronize()
        c.fr_ntt_dist_dev(p, log_n, d, inverse=True, post=t_un)         # back to coefficients, layout d1
        c.sync()

    group.run(run)
    want_mid = ctx.fr_ntt(ctx.fr_ntt(vec, inverse=True), coset=True)
    assert _assemble(group, mids, log_n, d) == want_mid
    assert _assemble(group, blocks, log_n, d1) == ctx.fr_ntt(want_mid, inverse=True, coset=True)


@pytest.mark.parametrize("world", [2, 8])
def test_dist_msm_total_on_every_rank(ctx, groups, world):
    """point-range sharding: every rank holds its own bases and scalars, the exchange is one affine
    point per rank, every rank ends with the total; checked against the closed form."""
    group = groups(world)
    n = 3000
    seeds = [0xBA5E + r for r in range(world)]
    scal = [ob.random_scalars(n, 900 + r) for r in range(world)]
    want1 = sum(int.from_bytes(ob.fr_dot(scal[r], ob.random_scalars(n, seeds[r]), n), "little") for r in range(world)) % o.R
    outs, hosts = [None] * world, [None] * world
    for g2 in (False, True):
        def run(r, c):
            bases = c.synthetic_bases(n, seeds[r], g2=g2)
            dev = torch.frombuffer(bytearray(scal[r]), dtype=torch.uint8).cuda()
            out = torch.zeros(128, dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            c.msm_dist_dev(bases, dev.data_ptr(), n, out.data_ptr())
            c.sync()
            outs[r] = bytes(out.cpu().numpy())[:128 if g2 else 64]
            buf = bytearray(scal[r])
            hosts[r] = c.msm_dist_raw(bases, torch.frombuffer(buf, dtype=torch.uint8).data_ptr(), n)
            bases.free()

        group.run(run)
        want = (ob.g2_mul_gen if g2 else ob.g1_mul_gen)(want1.to_bytes(32, "little"))
        assert all(x == want for x in outs), g2
        assert all(x == want for x in hosts), g2


@pytest.mark.parametrize("log_n,world", [(6, 1), (8, 2), (10, 4), (12, 8), (16, 2), (16, 8)])
def test_dist_prove_equals_single_gpu(ctx, zkb, groups, log_n, world):
    """zkb_groth16_prove_dist over emulated ranks: proof bytes identical to zkb_groth16_prove with the
    same (key seed, witness, r, s) — layouts, folded tables, sharded SpMV / MSMs and the on-device
    combination all have to be right for that — and pairing-verified."""
    group = groups(world)
    seed = 0x5EED0004 + log_n
    w = zkb.synthetic_witness(log_n, (3).to_bytes(32, "little"), (9).to_bytes(32, "little"))
    r, s = (21).to_bytes(32, "little"), (34).to_bytes(32, "little")
    pk = ctx.pk_synthetic(log_n, seed)
    want, _, _, _ = ctx.groth16_prove(pk, w, r, s)
    group.pk_synthetic(log_n, seed)
    assert group.vk == pk.vk
    proof, t = group.groth16_prove(w, r, s)
    assert proof == want
    assert zkb.groth16_verify(pk.vk, proof, w[32:64])
    # pinned host memory + replay
    wp = zkb.lib.host_alloc(len(w))
    zkb.lib.host_write(wp, w)
    proof2, _ = group.groth16_prove(wp, r, s)
    zkb.lib.host_free(wp)
    assert proof2 == want
    # a non-canonical witness element is seen by one rank only; every rank must refuse
    bad = bytearray(w)
    bad[32 * ((1 << log_n) - 1): 32 * (1 << log_n)] = int(o.R).to_bytes(32, "little")
    with pytest.raises(zkb.ZkbError, match="non-canonical"):
        group.groth16_prove(bytes(bad), r, s)
    assert group.groth16_prove(w, r, s)[0] == want            # and the group still works afterwards
    # the single-GPU entry point refuses a key held in the four-step layout
    with pytest.raises(zkb.ZkbError):
        group.ctxs[0].groth16_prove(group.pks[0], w, r, s)
    for k in group.pks:
        k.free()
    group.pks = None
    pk.free()


def test_batched_ntt_and_tables(ctx):
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
    # mode-1 table: scale * base^((row0 + hi) + lo * mul)
    rows, cols, row0, mul, base, scale = 3, 4, 5, 8, 7, 11
    tab = torch.empty(rows * cols * 32, dtype=torch.uint8, device="cuda")
    ctx.fr_table_dev(tab.data_ptr(), base.to_bytes(32, "little"), scale.to_bytes(32, "little"), rows, cols, row0, mul, 1)
    ctx.fr_convert_dev(tab.data_ptr(), rows * cols, False)
    ctx.sync()
    raw = bytes(tab.cpu().numpy())
    want = [scale * pow(base, (row0 + hi) + lo * mul, o.R) % o.R for hi in range(rows) for lo in range(cols)]
    assert [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(rows * cols)] == want
