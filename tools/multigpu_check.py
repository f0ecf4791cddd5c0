"""Run under torchrun on >= 2 GPUs: checks the NCCL paths against the single-GPU results.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py

 1. four-step NTT (dist_ntt.FourStepNtt.transform, NCCL all-to-all) == single-GPU zkb_fr_ntt
 2. Groth16 prove with every MSM sharded by points (sharding.groth16_prove_sharded)
 3. Groth16 prove with the witness map distributed as well (sharding.FourStepProver)
    both == the single-GPU proof bytes (rank 0 recomputes it) and pairing-verified.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigen_zeth_b200 as z  # noqa: E402
from eigen_zeth_b200.dist_ntt import FourStepNtt  # noqa: E402
from eigen_zeth_b200.sharding import FourStepProver, groth16_prove_sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = z.Context(local)
    ok = True
    # ---- 1. NTT
    for log_n in (12, 20):
        n = 1 << log_n
        g = torch.Generator().manual_seed(1234 + log_n)
        v = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, generator=g)
        v[:, 7] &= 0x0FFFFFFF
        vec = bytes(v.numpy().tobytes())
        full = torch.frombuffer(bytearray(vec), dtype=torch.uint8).to(dev)
        ctx.fr_convert_dev(full.data_ptr(), n, True)
        ctx.sync()
        eng = FourStepNtt(ctx, log_n, world, rank)
        d = log_n // 2
        for inverse in (False, True):
            blk = torch.empty(eng.n_loc * 32, dtype=torch.uint8, device=dev)
            ctx.fr_gather_layout_dev(full.data_ptr(), blk.data_ptr(), log_n, d, world, rank)
            eng.transform(blk, d, inverse=inverse)
            ctx.fr_convert_dev(blk.data_ptr(), eng.n_loc, False)
            ctx.sync()
            want_full = torch.frombuffer(bytearray(ctx.fr_ntt(vec, inverse=inverse)), dtype=torch.uint8).to(dev)
            want = torch.empty_like(blk)
            ctx.fr_gather_layout_dev(want_full.data_ptr(), want.data_ptr(), log_n, log_n - d, world, rank)
            ctx.sync()
            good = bool(torch.equal(blk, want))
            ok &= good
            if rank == 0:
                print("four-step NTT 2^%d inverse=%s world=%d: %s" % (log_n, inverse, world, "ok" if good else "MISMATCH"), flush=True)
    # ---- 2./3. Groth16
    log_n, seed = 16, 0x5EED0004
    w = z.synthetic_witness(log_n, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"))
    r, s = (11).to_bytes(32, "little"), (13).to_bytes(32, "little")
    want = None
    if rank == 0:
        pk1 = ctx.pk_synthetic(log_n, seed)
        want, _, _, _ = ctx.groth16_prove(pk1, w, r, s)
        assert z.groth16_verify(pk1.vk, want, w[32:64])
        pk1.free()
    spk = ctx.pk_synthetic_shard(log_n, seed, rank, world)
    proof, _ = groth16_prove_sharded(ctx, spk, w, r, s, world)
    spk.free()
    prover = FourStepProver(ctx, log_n, world, rank)
    fpk = ctx.pk_synthetic_shard(log_n, seed, rank, world, h_low_bits=prover.h_low_bits)
    proof4 = prover.prove(fpk, w, r, s)
    dist.barrier()
    t0 = time.perf_counter()
    wt16 = torch.frombuffer(bytearray(w), dtype=torch.uint8).pin_memory()
    proof4b = prover.prove(fpk, wt16, r, s)          # witness uploaded in slices + NVLink all-gather
    dist.barrier()
    dt = time.perf_counter() - t0
    if os.environ.get("ZKB_STAGES"):
        glog = int(os.environ["ZKB_STAGES"])
        p2 = FourStepProver(ctx, glog, world, rank)
        k2 = ctx.pk_synthetic_shard(glog, seed, rank, world, h_low_bits=p2.h_low_bits)
        w2 = z.synthetic_witness(glog, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"))
        wt = torch.frombuffer(bytearray(w2), dtype=torch.uint8).pin_memory()
        for i in range(3):
            tm = {}
            dist.barrier()
            p2.prove(k2, wt, r, s, timings=tm)
        if rank == 0:
            print("stages 2^%d world=%d:" % (glog, world), {k: round(v * 1e3, 2) for k, v in tm.items()}, flush=True)
        k2.free()
    if rank == 0:
        g1, g2 = proof == want, proof4 == want and proof4b == want
        ok &= g1 and g2
        print("Groth16 2^%d sharded MSMs: %s; four-step witness map: %s (%.2f ms)" % (
            log_n, "ok" if g1 else "MISMATCH", "ok" if g2 else "MISMATCH", dt * 1e3), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if not int(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("MULTIGPU OK", flush=True)


if __name__ == "__main__":
    main()
