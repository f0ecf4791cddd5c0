"""Host-side logic of the multi-GPU MSM (point-range sharding + all-gather of partials) with
world_size 2 over gloo on CPU.  The per-rank MSM is computed by the oracle here — this test checks
ranges, gather order and the combine step, not the CUDA kernels (those are `-m gpu`)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_index_space():
    sys.path.insert(0, ROOT)
    from eigen_zeth_b200.sharding import all_shards, shard_range
    for n in (0, 1, 7, 8, 1 << 20, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            sh = all_shards(n, world)
            assert sh[0][0] == 0 and sum(c for _, c in sh) == n
            for (f0, c0), (f1, _) in zip(sh, sh[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in sh) - min(c for _, c in sh) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n, seed, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_a as o
    import oracle_b as ob
    from eigen_zeth_b200.sharding import sharded_msm
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bases = ob.synthetic_bases(n, seed)
    scalars = ob.random_scalars(n, seed + 1)

    def local_msm(first, count):
        out, _ = ob.msm_g1(bases[64 * first: 64 * (first + count)], scalars[32 * first: 32 * (first + count)])
        return torch.frombuffer(bytearray(out), dtype=torch.uint8)

    def point_sum(gathered):
        acc = None
        for row in gathered:
            acc = o.g1_add(acc, o.g1_from_bytes(bytes(row.numpy())))
        return o.g1_to_bytes(acc)

    got = sharded_msm(local_msm, point_sum, n, world, rank)
    want, _ = ob.msm_g1(bases, scalars)
    q.put((rank, got == want))
    dist.destroy_process_group()


def test_sharded_msm_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 301, 77, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------- four-step NTT layout algebra over gloo
def _four_step_worker(rank, world, port, log_n, q):
    """the steps of the four-step NTT (csrc/dist.cu) with Python-int math for the local transforms and a
    real process group for the exchange (gloo has no all_to_all, so it is built from all_gather)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_a as o
    from eigen_zeth_b200.sharding import local_index_map
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, d = 1 << log_n, log_n // 2
    M, N1 = 1 << d, 1 << (log_n - d)
    mloc, n1loc = M // world, N1 // world
    rng = o.Xoshiro256ss(4242)
    full = [rng.fr() for _ in range(n)]
    want = o.ntt(full)
    idx_in = local_index_map(log_n, d, world, rank)              # [low_loc][high] -> global index
    x = [[full[idx_in[l][h]] for h in range(N1)] for l in range(mloc)]
    w = o.fr_root_of_unity(log_n)
    # (a) size-N1 transforms over `high`, then the inter-step twiddle w^{i2 * k1}
    y = []
    for l in range(mloc):
        row = o.ntt(x[l])
        i2 = rank * mloc + l
        y.append([row[k1] * pow(w, i2 * k1, o.R) % o.R for k1 in range(N1)])
    # (b) pack [Mloc][G][N1loc] -> [G][Mloc][N1loc], (c) exchange
    send = [[y[l][g * n1loc + k] for l in range(mloc) for k in range(n1loc)] for g in range(world)]
    gathered = [None] * world
    dist.all_gather_object(gathered, send)
    recv = [gathered[src][rank] for src in range(world)]         # chunk `rank` of every rank, in rank order
    # (d) unpack [M][N1loc] -> [N1loc][M], (e) size-M transforms
    flat = [v for chunk in recv for v in chunk]                  # [i2][k1_loc]
    out = []
    for k in range(n1loc):
        out.append(o.ntt([flat[i2 * n1loc + k] for i2 in range(M)]))
    idx_out = local_index_map(log_n, log_n - d, world, rank)     # layout L-d of the result
    ok = all(out[l][h] == want[idx_out[l][h]] for l in range(n1loc) for h in range(M))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("log_n", [6, 9])
def test_four_step_layout_world2_gloo(log_n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + log_n) % 2000
    procs = [ctx.Process(target=_four_step_worker, args=(r, 2, port, log_n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
