"""Distributed Groth16 prove at a given size under torchrun (BASELINE.json configs[4]: 2^26
constraints on 8 GPUs) through zkb_groth16_prove_dist: four-step NTTs with the exchange fused into
the NTT kernels + every MSM sharded by points.  torch.distributed only carries the communicator id.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/prove_large.py 26
"""
import json, os, resource, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigen_zeth_b200 as z
from eigen_zeth_b200 import launch

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = z.Context(local)
ctx.comm_init(launch.unique_id_via_torch(dev), rank, world)
t0 = time.time()
pk = ctx.pk_synthetic_dist(log_n, 0x5EED0004)
dist.barrier()
setup_s = time.time() - t0
t0 = time.time()
wp = z.lib.host_alloc(32 << log_n)
z.lib.synthetic_witness_into(log_n, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"), wp)
pub = z.lib.host_read(wp + 32, 32)
witness_s = time.time() - t0
r, s = (11).to_bytes(32, "little"), (13).to_bytes(32, "little")
times, stages = [], {}
for i in range(reps + 1):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    proof, tm = ctx.groth16_prove_dist(pk, wp, r, s)
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if i:
        times.append(float(dt.item()))
        stages = {k: round(v, 2) for k, v in tm.items() if v}
free_b, total_b = torch.cuda.mem_get_info()
if rank == 0:
    ok = z.groth16_verify(pk.vk, proof, pub)
    out = {"log_n": log_n, "n_gpus": world, "prove_ms_best": round(min(times) * 1e3, 2),
           "prove_ms_all": [round(t * 1e3, 2) for t in times], "device_stage_ms_rank0": stages,
           "exchange": ctx.comm_transport, "pairing_verified": bool(ok), "setup_s": round(setup_s, 1),
           "witness_gen_s": round(witness_s, 1), "gpu_mem_used_gib_rank0": round((total_b - free_b) / 2**30, 1),
           "host_rss_gib_rank0": round(resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 2**20, 1)}
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "prove_large_%d_n%d.json" % (log_n, world)), "w") as f:
        json.dump(out, f, indent=1)
z.lib.host_free(wp)
pk.free()
ctx.close()
dist.destroy_process_group()
