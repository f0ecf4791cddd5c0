"""The NCCL transport between PROCESSES (zkb_comm_init + CUDA-IPC receive windows, and the
ncclSend/ncclRecv staging fallback): needs >= 2 visible GPUs, otherwise skipped — the same library
code runs with emulated ranks in tests/test_gpu_dist.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("staging", [0, 1])
def test_nccl_paths_match_single_gpu(staging):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 1 << (n.bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29533 + staging), os.path.join(ROOT, "tools", "dist_check.py")]
    env = dict(os.environ, ZKB_FORCE_STAGING=str(staging))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "MULTIGPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_device_group_on_all_visible_gpus(zkb):
    """zkb_comm_init_local across real devices (the prover service's --devices mode): peer pointers
    over NVLink, no NCCL"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 1 << (n.bit_length() - 1)
    group = zkb.lib.DeviceGroup(list(range(world)))
    log_n, seed = 16, 0x5EED0004
    w = zkb.synthetic_witness(log_n, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"))
    r, s = (11).to_bytes(32, "little"), (13).to_bytes(32, "little")
    group.pk_synthetic(log_n, seed)
    proof, _ = group.groth16_prove(w, r, s)
    c0 = zkb.Context(0)
    pk = c0.pk_synthetic(log_n, seed)
    want, _, _, _ = c0.groth16_prove(pk, w, r, s)
    assert proof == want and zkb.groth16_verify(pk.vk, proof, w[32:64])
    pk.free()
    c0.close()
    group.close()
