"""tiny driver for profiling: a few 2^24 Fr NTTs (forward, full twiddle tables) on device data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eigen_zeth_b200 as z
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ctx = z.Context(0)
n = 1 << log_n
x = torch.randint(0, 2**31 - 1, (n, 8), dtype=torch.int32, device="cuda")
x[:, 7] &= 0x0FFFFFFF
stream = torch.cuda.ExternalStream(ctx.stream)
for _ in range(3):
    ctx.fr_ntt_dev(x.data_ptr(), x.data_ptr(), log_n, montgomery=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5):
    ctx.fr_ntt_dev(x.data_ptr(), x.data_ptr(), log_n, montgomery=True)
e1.record(stream); e1.synchronize()
print("ntt 2^%d: %.3f ms" % (log_n, e0.elapsed_time(e1) / 5))
