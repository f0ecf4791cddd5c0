"""BASELINE.json configs[4] emulation: the ProverService (one process, --devices GPUs behind ONE
address) serving eigen-zeth's client state machine (tools/operator_replay.py) with the reference's
1 s inter-step sleeps (src/prover/provider.rs:542), then once more with the sleeps at 0 to expose
the GenFinalProof RPC latency and its host-side breakdown.

  python tools/service_demo.py <log_n> [devices, e.g. 0,1,2,3,4,5,6,7] [step_sleep]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import eigen_zeth_b200 as z
from eigen_zeth_b200 import proof_format as pf, prover_service as ps
from operator_replay import OperatorReplay
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
devices = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
sleep = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
t0 = time.time()
backend = ps.GpuGroth16Backend(log_n=log_n, devices=devices)
setup_s = time.time() - t0
servicer = ps.ProverServicer(backend)
srv, port = ps.make_server(servicer, "127.0.0.1:0")
srv.start()
out = {"log_n": log_n, "devices": devices, "setup_s": round(setup_s, 2), "runs": []}
for step_sleep, batches in ((sleep, 3), (0.0, 3)):
    op = OperatorReplay("127.0.0.1:%d" % port, step_sleep=step_sleep)
    rows = []
    for k in range(1, batches + 1):
        t0 = time.time()
        res, t = op.execute(k, batch_id="demo-%s-%d" % (step_sleep, k))
        e2e = time.time() - t0
        ok = z.groth16_verify(backend.vk, pf.proof_from_json(res.proof), pf.public_input_from_json(res.public_input))
        tm = servicer.timings[-1]
        rows.append({"batch": k, "e2e_s": round(e2e, 3), "final_proof_rpc_ms": round(1e3 * t["final"], 2),
                     "device_prove_ms": round(tm["ms_total"], 2),
                     "host_breakdown_ms": {k2: round(v, 2) for k2, v in tm.items() if k2.endswith("_ms") and not k2.startswith("ms_")},
                     "pairing_verified": bool(ok)})
        print(rows[-1], flush=True)
    op.close()
    out["runs"].append({"step_sleep_s": step_sleep, "batches": rows})
srv.stop(0)
backend.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "service_demo_2p%d_n%d.json" % (log_n, len(devices))), "w"), indent=1)
print(json.dumps(out))
