"""MSM time must not depend on the scalar distribution (chunked walk + iterated partial merge):
uniform vs 0/1 (boolean witness wires) vs all-equal scalars at 2^22."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, eigen_zeth_b200 as z
from bench import uniform_fr_scalars
ctx = z.Context(0); stream = torch.cuda.ExternalStream(ctx.stream)
n = 1 << 22
b = ctx.synthetic_bases(n, 3); b.precompute(0)
res = torch.zeros(64, dtype=torch.uint8, device="cuda")
u = uniform_fr_scalars(n, 1)
cases = {"uniform": u}
bits = np.zeros((n, 4), dtype=np.uint64); bits[:, 0] = np.arange(n, dtype=np.uint64) & np.uint64(1); cases["zero_one"] = bits
same = np.tile(u[:1], (n, 1)); cases["all_equal"] = same
half = u.copy(); half[::2] = 0; half[::2, 0] = 1; cases["half_ones"] = half
for name, arr in cases.items():
    sc = torch.from_numpy(arr.view("int64")).cuda()
    for _ in range(2): ctx.msm_dev(b, sc.data_ptr(), n, res.data_ptr())
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); ctx.msm_dev(b, sc.data_ptr(), n, res.data_ptr()); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%-10s %.3f ms" % (name, sorted(ts)[2]), flush=True)
