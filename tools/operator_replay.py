"""Replay of eigen-zeth's prover client (reference src/prover/provider.rs:243-544) as a test driver.

It sends exactly the request sequence ProverChannel::execute sends for one batch, matches responses
by VARIANT only (provider.rs:646-668), enforces the client's own tolerances (32-byte roots
:323-324, at least one chunk proof :384-387, final_proof present :492) and feeds the returned
strings through the parse_proof / parse_public_input restatement.  `step_sleep` stands for the
hard-coded 1 s between transitions (provider.rs:542).
"""
from __future__ import annotations

import os
import queue
import sys
import time
import uuid

import grpc

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eigen_zeth_b200 import proof_format, protocol as pb  # noqa: E402

DEFAULT_AGGREGATOR = "479881985774944702531460751064278034642760119942"    # commands/run.rs:76


class OperatorReplay:
    def __init__(self, address: str, step_sleep: float = 0.0, chain_id: int = 12345, program: str = "evm",
                 curve: str = "BN128", aggregator_addr: str = DEFAULT_AGGREGATOR):
        self.channel = grpc.insecure_channel(address)
        self.call = self.channel.stream_stream("/%s/%s" % (pb.SERVICE, pb.METHOD),
                                               request_serializer=pb.ProverRequest.SerializeToString,
                                               response_deserializer=pb.ProverResponse.FromString)
        self.q = queue.Queue()
        self.responses = self.call(iter(self.q.get, None))
        self.step_sleep, self.chain_id, self.program, self.curve, self.aggregator = step_sleep, chain_id, program, curve, aggregator_addr

    def _rpc(self, req, want):
        self.q.put(req)
        while True:
            resp = next(self.responses)
            kind = resp.WhichOneof("response_type")
            if kind == "get_status":            # logged and dropped by the client (provider.rs:651-654)
                if want == "get_status":
                    return resp
                continue
            assert kind == want, "server answered %s to a %s request" % (kind, want)
            return resp

    def get_status(self):
        return self._rpc(pb.ProverRequest(id=str(uuid.uuid4()), get_status=pb.GetStatusRequest()), "get_status").get_status

    def execute(self, block_number: int, batch_id: str | None = None):
        """-> (ProofResult, dict of per-step seconds)"""
        batch_id = batch_id or str(uuid.uuid4())
        t = {}
        t0 = time.time()
        r = self._rpc(pb.ProverRequest(id=str(uuid.uuid4()), gen_batch_proof=pb.GenBatchProofRequest(
            gen_batch_chunks=pb.GenBatchChunks(batch_id=batch_id, batch=pb.Batch(block_number=[block_number]),
                                               chain_id=self.chain_id, program_name=self.program))), "gen_batch_proof")
        chunks = r.gen_batch_proof.gen_batch_chunks
        assert r.gen_batch_proof.WhichOneof("step") == "gen_batch_chunks" and chunks.result_code == pb.COMPLETED_OK
        assert len(chunks.pre_state_root) == 32 and len(chunks.post_state_root) == 32
        t["chunks"] = time.time() - t0
        time.sleep(self.step_sleep)
        t0 = time.time()
        r = self._rpc(pb.ProverRequest(id=str(uuid.uuid4()), gen_batch_proof=pb.GenBatchProofRequest(
            gen_chunk_proof=pb.GenChunkProof(batch_id=batch_id, task_id=chunks.task_id, chunk_count=chunks.chunk_count,
                                             chain_id=self.chain_id, program_name=self.program,
                                             batch_data=chunks.batch_data))), "gen_batch_proof")
        cp = r.gen_batch_proof.gen_chunk_proof
        assert cp.result_code == pb.COMPLETED_OK and len(cp.batch_proof_result.chunk_proofs) >= 1
        first, last = cp.batch_proof_result.chunk_proofs[0].proof, cp.batch_proof_result.chunk_proofs[-1].proof
        t["chunk_proof"] = time.time() - t0
        time.sleep(self.step_sleep)
        t0 = time.time()
        r = self._rpc(pb.ProverRequest(id=str(uuid.uuid4()), gen_aggregated_proof=pb.GenAggregatedProofRequest(
            batch_id=batch_id, recursive_proof_1=first, recursive_proof_2=last)), "gen_aggregated_proof")
        agg = r.gen_aggregated_proof
        assert agg.result_code == pb.COMPLETED_OK
        t["aggregate"] = time.time() - t0
        time.sleep(self.step_sleep)
        t0 = time.time()
        r = self._rpc(pb.ProverRequest(id=str(uuid.uuid4()), gen_final_proof=pb.GenFinalProofRequest(
            batch_id=batch_id, recursive_proof=agg.result_string, curve_name=self.curve,
            aggregator_addr=self.aggregator)), "gen_final_proof")
        fin = r.gen_final_proof
        t["final"] = time.time() - t0
        if fin.result_code != pb.COMPLETED_OK:
            raise RuntimeError("GenFinalProof error: %s" % fin.error_message)
        assert fin.HasField("final_proof")
        proof_format.parse_proof(fin.final_proof.proof)
        proof_format.parse_public_input(fin.final_proof.public_input)
        result = proof_format.ProofResult(block_number, fin.final_proof.proof, fin.final_proof.public_input,
                                          bytes(chunks.pre_state_root), bytes(chunks.post_state_root))
        return result, t

    def close(self):
        self.q.put(None)
        self.channel.close()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--address", default="127.0.0.1:50061")
    ap.add_argument("--batches", type=int, default=3)
    ap.add_argument("--step-sleep", type=float, default=1.0)
    a = ap.parse_args()
    op = OperatorReplay(a.address, a.step_sleep)
    for k in range(1, a.batches + 1):
        t0 = time.time()
        res, t = op.execute(k)
        print("batch %d: e2e %.3f s, server-side final %.3f s" % (k, time.time() - t0, t["final"]))
    op.close()
