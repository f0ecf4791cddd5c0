"""Run under torchrun on >= 2 GPUs: the library's own multi-GPU path (zkb_comm_init over NCCL, receive
windows mapped with CUDA IPC, or ncclSend/ncclRecv staging with ZKB_FORCE_STAGING=1) against the
single-GPU results.  torch.distributed only carries the 128-byte unique id and the final verdict.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigen_zeth_b200 as z  # noqa: E402
from eigen_zeth_b200 import launch  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = z.Context(local)
    if os.environ.get("ZKB_FORCE_STAGING") == "1":
        ctx.set("comm_force_staging", 1)
    ctx.comm_init(launch.unique_id_via_torch(dev), rank, world)
    assert (ctx.comm_rank, ctx.comm_world) == (rank, world)
    ok = True
    # ---- 1. four-step NTT with the fused exchange
    for log_n in (12, 20, 22):
        n = 1 << log_n
        g = torch.Generator().manual_seed(1234 + log_n)
        v = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, generator=g)
        v[:, 7] &= 0x0FFFFFFF
        vec = bytes(v.numpy().tobytes())
        full = torch.frombuffer(bytearray(vec), dtype=torch.uint8).to(dev)
        ctx.fr_convert_dev(full.data_ptr(), n, True)
        ctx.sync()
        d = log_n // 2
        n_loc = n // world
        for inverse in (False, True):
            blk = torch.empty(n_loc * 32, dtype=torch.uint8, device=dev)
            ctx.fr_gather_layout_dev(full.data_ptr(), blk.data_ptr(), log_n, d, world, rank)
            ctx.sync()
            torch.cuda.synchronize()
            ctx.fr_ntt_dist_dev(blk.data_ptr(), log_n, d, inverse=inverse)
            ctx.fr_convert_dev(blk.data_ptr(), n_loc, False)
            ctx.sync()
            want_full = torch.frombuffer(bytearray(ctx.fr_ntt(vec, inverse=inverse)), dtype=torch.uint8).to(dev)
            want = torch.empty_like(blk)
            ctx.fr_gather_layout_dev(want_full.data_ptr(), want.data_ptr(), log_n, log_n - d, world, rank)
            ctx.sync()
            good = bool(torch.equal(blk, want))
            ok &= good
            if rank == 0:
                print("four-step NTT 2^%d inverse=%s world=%d [%s]: %s" % (log_n, inverse, world, ctx.comm_transport,
                                                                           "ok" if good else "MISMATCH"), flush=True)
    # ---- 2. sharded MSM: every rank gets the total
    n = 1 << 16
    bases = ctx.synthetic_bases(n, 0xBA5E + rank)
    sc = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, generator=torch.Generator().manual_seed(rank))
    sc[:, 3] &= (1 << 60) - 1
    scd = sc.to(dev)
    out = torch.zeros(64, dtype=torch.uint8, device=dev)
    part = torch.zeros(64, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.msm_dist_dev(bases, scd.data_ptr(), n, out.data_ptr())
    ctx.msm_dev(bases, scd.data_ptr(), n, part.data_ptr())
    ctx.sync()
    parts = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(parts, part)
    want = ctx.point_sum(b"".join(bytes(p.cpu().numpy()) for p in parts))
    good = bytes(out.cpu().numpy()) == want and ctx.msm_dist_raw(bases, sc.data_ptr(), n) == want
    ok &= good
    if rank == 0:
        print("sharded MSM 2^16 x %d ranks: %s" % (world, "ok" if good else "MISMATCH"), flush=True)
    # ---- 3. Groth16 over all ranks == single GPU
    for log_n in (10, 16, 18):
        seed = 0x5EED0004
        w = z.synthetic_witness(log_n, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"))
        r, s = (11).to_bytes(32, "little"), (13).to_bytes(32, "little")
        want = None
        if rank == 0:
            c1 = z.Context(local)
            pk1 = c1.pk_synthetic(log_n, seed)
            want, _, _, _ = c1.groth16_prove(pk1, w, r, s)
            assert z.groth16_verify(pk1.vk, want, w[32:64])
            pk1.free()
            c1.close()
        dpk = ctx.pk_synthetic_dist(log_n, seed)
        proof, t = ctx.groth16_prove_dist(dpk, w, r, s)
        wp = z.lib.host_alloc(len(w))
        z.lib.host_write(wp, w)
        dist.barrier()
        t0 = time.perf_counter()
        proof2, t = ctx.groth16_prove_dist(dpk, wp, r, s)
        dt = time.perf_counter() - t0
        z.lib.host_free(wp)
        dpk.free()
        if rank == 0:
            good = proof == want and proof2 == want
            ok &= good
            print("Groth16 2^%d over %d ranks: %s (%.2f ms; %s)" % (log_n, world, "ok" if good else "MISMATCH", dt * 1e3,
                                                                   {k: round(v, 2) for k, v in t.items() if v}), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU OK", flush=True)


if __name__ == "__main__":
    main()
