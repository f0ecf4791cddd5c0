"""BASELINE.json configs[4] emulation: the ProverService on one B200 serving eigen-zeth's client
state machine (tools/operator_replay.py) with the reference's 1 s inter-step sleeps."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import eigen_zeth_b200 as z
from eigen_zeth_b200 import proof_format as pf, prover_service as ps
from operator_replay import OperatorReplay
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
sleep = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t0 = time.time()
backend = ps.GpuGroth16Backend(log_n=log_n)
setup_s = time.time() - t0
servicer = ps.ProverServicer(backend)
srv, port = ps.make_server(servicer, "127.0.0.1:0")
srv.start()
op = OperatorReplay("127.0.0.1:%d" % port, step_sleep=sleep)
rows = []
for k in range(1, 4):
    t0 = time.time()
    res, t = op.execute(k)
    e2e = time.time() - t0
    ok = z.groth16_verify(backend.vk, pf.proof_from_json(res.proof), pf.public_input_from_json(res.public_input))
    rows.append({"batch": k, "e2e_s": round(e2e, 3), "final_proof_rpc_s": round(t["final"], 4),
                 "device_prove_ms": round(servicer.timings[-1]["ms_total"], 2), "pairing_verified": bool(ok)})
    print(rows[-1], flush=True)
op.close(); srv.stop(0)
out = {"log_n": log_n, "setup_s": round(setup_s, 2), "step_sleep_s": sleep, "batches": rows}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "service_demo.json"), "w"), indent=1)
print(json.dumps(out))
