"""Real NCCL paths (four-step NTT all-to-all, sharded Groth16): needs >= 2 visible GPUs, otherwise
skipped — the same logic is covered with emulated ranks in test_gpu_dist_ntt.py / test_gpu_groth16.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_paths_match_single_gpu():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_four_step_prover_world1_equals_standard_prove(ctx, zkb):
    """the distributed witness-map chain with a single rank (no collective) must reproduce the
    standard prover bit for bit — this pins the layout bookkeeping and the folded tables."""
    from eigen_zeth_b200.sharding import FourStepProver
    for log_n in (6, 9, 12, 16):
        seed = 0x5EED0004 + log_n
        w = zkb.synthetic_witness(log_n, (3).to_bytes(32, "little"), (9).to_bytes(32, "little"))
        r, s = (21).to_bytes(32, "little"), (34).to_bytes(32, "little")
        pk = ctx.pk_synthetic(log_n, seed)
        want, _, h_want, _ = ctx.groth16_prove(pk, w, r, s, want_h=True)
        prover = FourStepProver(ctx, log_n, 1, 0)
        fpk = ctx.pk_synthetic_shard(log_n, seed, 0, 1, h_low_bits=prover.h_low_bits)
        assert prover.prove(fpk, w, r, s) == want
        assert zkb.groth16_verify(pk.vk, want, w[32:64])
        pk.free()
        fpk.free()
