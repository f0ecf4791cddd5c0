import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, eigen_zeth_b200 as z
ctx = z.Context(0)
glog = 22
pk = ctx.pk_synthetic(glog, 0x5EED0004)
w = z.synthetic_witness(glog, (5).to_bytes(32, "little"), (7).to_bytes(32, "little"))
wt = torch.frombuffer(bytearray(w), dtype=torch.uint8).pin_memory()
r, s = (11).to_bytes(32, "little"), (13).to_bytes(32, "little")
for i in range(4):
    t0 = time.perf_counter(); proof, _, _, tm = ctx.groth16_prove(pk, wt.data_ptr(), r, s); dt = time.perf_counter() - t0
print("prove 2^22: %.2f ms" % (dt * 1e3), {k: round(v, 2) for k, v in tm.items()}, z.groth16_verify(pk.vk, proof, w[32:64]), flush=True)
