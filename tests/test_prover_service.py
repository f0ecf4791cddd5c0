"""The outer boundary: prover.v1.ProverService over real gRPC, driven by a replay of eigen-zeth's
client (tools/operator_replay.py, mirroring reference src/prover/provider.rs:243-544).
CPU tests use a stub backend (protocol / state-machine behaviour only); the GPU test proves for real."""
import json
import os
import sys

import pytest

import oracle_a as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")))


class CannedBackend:
    """returns the reference's canned proof (what DEBUG_PROOF does, src/settlement/worker.rs:49-96)."""

    def __init__(self):
        self.calls = 0
        self.fail_next = False

    def final_proof(self, batch_id, recursive_proof, aggregator_addr):
        self.calls += 1
        if self.fail_next:
            self.fail_next = False
            raise RuntimeError("injected failure")
        return GOLDEN["proof_json"], GOLDEN["public_input_json"], {}


@pytest.fixture()
def server():
    from eigen_zeth_b200 import prover_service as ps
    backend = CannedBackend()
    servicer = ps.ProverServicer(backend)
    srv, port = ps.make_server(servicer, "127.0.0.1:0")
    srv.start()
    yield "127.0.0.1:%d" % port, backend, servicer
    srv.stop(0)


def test_wire_format_is_prost_compatible():
    """hand-encoded protobuf bytes of a GenFinalProofRequest (field numbers of prover.proto:13-23,130-135)."""
    from eigen_zeth_b200 import protocol as pb
    inner = b"\x0a\x01b" + b"\x12\x02rp" + b"\x1a\x05BN128" + b"\x22\x011"
    wire = b"\x0a\x01x" + b"\x2a" + bytes([len(inner)]) + inner
    req = pb.ProverRequest.FromString(wire)
    assert req.id == "x" and req.WhichOneof("request_type") == "gen_final_proof"
    assert (req.gen_final_proof.batch_id, req.gen_final_proof.recursive_proof, req.gen_final_proof.curve_name,
            req.gen_final_proof.aggregator_addr) == ("b", "rp", "BN128", "1")
    assert req.SerializeToString() == wire
    # ChunkProof has proof_key = 3 and proof = 2 (prover.proto:107-111)
    cp = pb.ChunkProof(chunk_id=1, proof_key="k", proof="p")
    assert cp.SerializeToString() == b"\x08\x01\x12\x01p\x1a\x01k"
    assert pb.METHOD_PATH == "/prover.v1.ProverService/ProverStream"


def test_full_batch_sequence_and_tolerances(server):
    from operator_replay import OperatorReplay
    from eigen_zeth_b200 import proof_format as pf
    addr, backend, _ = server
    op = OperatorReplay(addr, step_sleep=0.0)
    res, t = op.execute(block_number=7)
    assert res.block_number == 7 and len(res.pre_state_root) == 32 and len(res.post_state_root) == 32
    assert pf.parse_proof(res.proof) == o.parse_proof(GOLDEN["proof_json"])
    assert pf.ProofResult.from_json(res.to_json()) == res
    # replay of the same batch id (client resume path) is served from the cache, bit-identical
    res1, _ = op.execute(block_number=8, batch_id="fixed-batch")
    calls = backend.calls
    res2, _ = op.execute(block_number=8, batch_id="fixed-batch")
    assert backend.calls == calls and res2 == res1
    # GetStatus is answered in-stream and does not disturb the sequence
    st = op.get_status()
    assert st.prover_status.version_server and st.prover_status.number_of_cores > 0
    assert st.prover_status.last_computed_request_id
    op.close()


def test_errors_keep_the_stream_open(server):
    from operator_replay import OperatorReplay
    addr, backend, _ = server
    op = OperatorReplay(addr, step_sleep=0.0, curve="BLS12381")
    with pytest.raises(RuntimeError, match="unsupported curve"):
        op.execute(1)
    op.curve = "BN128"
    backend.fail_next = True
    with pytest.raises(RuntimeError, match="injected failure"):
        op.execute(2)
    res, _ = op.execute(3)          # same stream still works: COMPLETED_ERROR never drops it
    assert res.block_number == 3
    op.close()


def test_batch_ledger_and_concurrent_streams(server):
    """per-batch bookkeeping, idempotent replay of every step, several operator streams
    queueing on one prover."""
    import threading
    import time
    from operator_replay import OperatorReplay
    from eigen_zeth_b200 import protocol as pb
    addr, backend, servicer = server
    op = OperatorReplay(addr, step_sleep=0.0)
    res, _ = op.execute(block_number=11, batch_id="ledger-1")
    row = servicer.ledger.view("ledger-1")
    assert row["steps"] == {"chunks": 1, "chunk_proofs": 1, "aggregated": 1, "final": 1}
    assert row["facts"]["chunk_count"] == 2 and json.loads(row["facts"]["batch_data"])["blocks"] == [11]
    op.execute(block_number=11, batch_id="ledger-1")                 # full replay: every step counted twice, no error
    assert servicer.ledger.view("ledger-1")["steps"]["chunks"] == 2
    # a chunk-proof request that contradicts what GenBatchChunks handed out is refused (and the stream survives)
    bad = pb.ProverRequest(id="x", gen_batch_proof=pb.GenBatchProofRequest(gen_chunk_proof=pb.GenChunkProof(
        batch_id="ledger-1", task_id=row["facts"]["task_id"], chunk_count=5, batch_data=row["facts"]["batch_data"])))
    r = servicer.handle(bad)
    assert r.gen_batch_proof.gen_chunk_proof.result_code == pb.COMPLETED_ERROR
    assert "chunk_count" in r.gen_batch_proof.gen_chunk_proof.error_message
    # ... the same batch id re-submitted with other blocks likewise
    with pytest.raises(AssertionError):          # the replay client asserts COMPLETED_OK on this step
        op.execute(block_number=12, batch_id="ledger-1")
    # a step for a batch this server has never seen (server restart) is accepted
    ok = pb.ProverRequest(id="y", gen_batch_proof=pb.GenBatchProofRequest(gen_chunk_proof=pb.GenChunkProof(
        batch_id="after-restart", task_id="0000000001", chunk_count=3, batch_data="{}")))
    r = servicer.handle(ok)
    assert r.gen_batch_proof.gen_chunk_proof.result_code == pb.COMPLETED_OK
    assert len(r.gen_batch_proof.gen_chunk_proof.batch_proof_result.chunk_proofs) == 3
    op.close()

    # four operator streams at once against one (slow) prover: all served, proofs never overlap
    state = {"inside": 0, "max_inside": 0}
    real = backend.final_proof

    def slow(*a):
        state["inside"] += 1
        state["max_inside"] = max(state["max_inside"], state["inside"])
        time.sleep(0.05)
        state["inside"] -= 1
        return real(*a)

    backend.final_proof = slow
    results, errors = {}, []

    def one(k):
        try:
            o_ = OperatorReplay(addr, step_sleep=0.0)
            results[k] = o_.execute(block_number=100 + k, batch_id="stream-%d" % k)[0]
            o_.close()
        except Exception as ex:
            errors.append(ex)

    threads = [threading.Thread(target=one, args=(k,)) for k in range(4)]
    [t.start() for t in threads]
    [t.join(timeout=30) for t in threads]
    assert not errors and sorted(results) == [0, 1, 2, 3]
    assert [results[k].block_number for k in range(4)] == [100, 101, 102, 103]
    assert state["max_inside"] == 1 and servicer.queued == 0


@pytest.mark.gpu
def test_service_proves_on_gpu_and_verifies():
    from operator_replay import OperatorReplay
    import eigen_zeth_b200 as z
    from eigen_zeth_b200 import proof_format as pf, prover_service as ps
    backend = ps.GpuGroth16Backend(log_n=12)
    srv, port = ps.make_server(ps.ProverServicer(backend), "127.0.0.1:0")
    srv.start()
    try:
        op = OperatorReplay("127.0.0.1:%d" % port, step_sleep=0.0)
        seen = set()
        for k in (1, 2):
            res, t = op.execute(k, batch_id="batch-%d" % k)
            proof = pf.proof_from_json(res.proof)
            pub = pf.public_input_from_json(res.public_input)
            assert z.groth16_verify(backend.vk, proof, pub)
            A, B, C = o.parse_proof(res.proof)
            assert o.g1_is_on_curve(A) and o.g2_is_on_curve(B) and o.g1_is_on_curve(C)
            seen.add(res.proof)
            res_again, _ = op.execute(k, batch_id="batch-%d" % k)
            assert res_again.proof == res.proof
        assert len(seen) == 2
        op.close()
    finally:
        srv.stop(0)


def test_blinding_is_keyed_and_bound_to_the_whole_request():
    """r, s must not be computable from public request fields, and one (r, s) must never
    cover two different witnesses."""
    from eigen_zeth_b200 import prover_service as ps
    s1, s2 = b"\x01" * 32, b"\x02" * 32
    r = ps._blinding(s1, "r", "batch", "proof", "addr")
    assert len(r) == 32 and int.from_bytes(r, "little") < o.R
    assert r == ps._blinding(s1, "r", "batch", "proof", "addr")              # replay: same blinding
    assert r != ps._blinding(s2, "r", "batch", "proof", "addr")              # keyed by the server secret
    assert r != ps._blinding(s1, "s", "batch", "proof", "addr")
    assert r != ps._blinding(s1, "r", "batch", "other proof", "addr")        # another witness -> another r
    assert r != ps._blinding(s1, "r", "batch", "proof", "other addr")
    assert r != ps._blinding(s1, "r", "batc", "hproof", "addr")              # length-prefixed: no concatenation games
    import hashlib
    assert r != (int.from_bytes(hashlib.sha256(b"r" + b"batch").digest(), "little") & ((1 << 253) - 1)).to_bytes(32, "little")


def test_secret_file_round_trip(tmp_path):
    from eigen_zeth_b200 import prover_service as ps
    path = str(tmp_path / "blinding.key")
    a = ps.load_or_create_secret(path)
    assert len(a) == 32 and ps.load_or_create_secret(path) == a
    assert (os.stat(path).st_mode & 0o777) == 0o600
    assert ps.load_or_create_secret(None) != ps.load_or_create_secret(None)


def test_service_refuses_synthetic_setup_without_the_flag(capsys):
    from eigen_zeth_b200 import prover_service as ps
    with pytest.raises(SystemExit):
        ps.main(["--log-n", "8"])
    assert "insecure-synthetic-setup" in capsys.readouterr().err


def test_handler_failure_answers_with_the_request_variant(server):
    """the client forwards only same-variant responses (provider.rs:646-668); a handler bug
    on a gen_batch_proof step must therefore be answered as gen_batch_proof, never as get_status."""
    from eigen_zeth_b200 import protocol as pb
    _, _, servicer = server

    def boom(req):
        raise RuntimeError("ledger exploded")

    servicer._gen_batch_chunks = boom
    servicer._gen_chunk_proof = boom
    r = servicer.handle(pb.ProverRequest(id="1", gen_batch_proof=pb.GenBatchProofRequest(
        gen_batch_chunks=pb.GenBatchChunks(batch_id="b1", batch=pb.Batch(block_number=[1]), chain_id=1, program_name="evm"))))
    assert r.WhichOneof("response_type") == "gen_batch_proof"
    res = r.gen_batch_proof.gen_batch_chunks
    assert r.gen_batch_proof.WhichOneof("step") == "gen_batch_chunks" and res.result_code == pb.COMPLETED_ERROR
    assert res.batch_id == "b1" and "ledger exploded" in res.error_message
    assert len(res.pre_state_root) == 32 and len(res.post_state_root) == 32
    r = servicer.handle(pb.ProverRequest(id="2", gen_batch_proof=pb.GenBatchProofRequest(
        gen_chunk_proof=pb.GenChunkProof(batch_id="b1", task_id="t", chunk_count=1, batch_data="{}"))))
    assert r.WhichOneof("response_type") == "gen_batch_proof"
    assert r.gen_batch_proof.WhichOneof("step") == "gen_chunk_proof"
    assert r.gen_batch_proof.gen_chunk_proof.result_code == pb.COMPLETED_ERROR


def test_aggregate_step_prefetches_the_final_witness(server):
    """the witness of GenFinalProof depends on the string handed out by GenAggregatedProof: the
    servicer starts building it as soon as it hands that string out (it arrives back >= 1 s later,
    provider.rs:542)."""
    from operator_replay import OperatorReplay
    addr, backend, servicer = server
    seen = []
    backend.prefetch = lambda rp, agg: seen.append((rp, agg))
    op = OperatorReplay(addr, step_sleep=0.0)
    res, _ = op.execute(5, batch_id="pf-1")
    op.close()
    # started at the chunk step already (three client sleeps ahead of GenFinalProof), confirmed at the
    # aggregate step: both name the very string the client later sends as recursive_proof
    assert len(seen) == 2 and seen[0] == seen[1]
    rp, agg = seen[0]
    assert agg == op.aggregator and len(rp) == 64
