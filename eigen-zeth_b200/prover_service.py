"""gRPC `prover.v1.ProverService` server — the outer drop-in boundary.

eigen-zeth's operator connects to PROVER_ADDR (default http://127.0.0.1:50061, reference
src/config/env.rs:21) and drives one bidirectional stream (src/prover/provider.rs:603-705) through
  GenBatchChunks -> GenChunkProof -> GenAggregatedProof -> GenFinalProof   (provider.rs:276-544).
Obligations this server keeps (each derived from the client code):
  * every request is answered with the SAME response variant; ids are echoed but the client matches
    by variant only (provider.rs:646-668);
  * result_code COMPLETED_OK advances the client, anything else makes it resend (provider.rs:511-516),
    so failures are reported as COMPLETED_ERROR + error_message and the stream is never dropped
    (the client has no restart arm in Aggregate/Final, provider.rs:436-438,486-487);
  * pre/post_state_root are exactly 32 bytes (provider.rs:323-324);
  * chunk_proofs is non-empty (provider.rs:384-387); final_proof is present on OK (provider.rs:492);
  * a replayed step with the same batch_id returns the same bytes (resume path, provider.rs:245-274):
    r, s and the witness are derived from the request, and results are cached per batch_id.
The batch-execution / chunk-STARK / aggregation steps belong to the external zkVM and are out of
scope: they are deterministic synthetic stubs.  GenFinalProof is the
real thing: a BN254 Groth16 proof computed on the B200 through libzkb.so, pairing-checked before it
leaves, and rendered in the proof.json layout of reference proof/proof.json.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import logging
import os
import threading
import time
from collections import OrderedDict
from concurrent import futures

import grpc

from . import proof_format, protocol as pb

log = logging.getLogger("zkb.prover")
R = proof_format.R
VERSION_SERVER = "zkb-0.1.0"
VERSION_PROTO = "v0_0_1"


def _h(*parts) -> bytes:
    m = hashlib.sha256()
    for p in parts:
        b = p if isinstance(p, bytes) else str(p).encode()
        m.update(len(b).to_bytes(8, "little"))
        m.update(b)
    return m.digest()


def _fr_from_hash(*parts) -> bytes:
    """deterministic canonical Fr: SHA-256 -> 253 bits (< r)."""
    v = int.from_bytes(_h(*parts), "little") & ((1 << 253) - 1)
    return v.to_bytes(32, "little")


class GpuGroth16Backend:
    """GenFinalProof on one B200.  No CPU fallback: constructing it without the CUDA library or an
    sm_100 GPU raises."""

    def __init__(self, log_n: int = 16, setup_seed: int = 0x5EED0004, device: int = 0, self_check: bool = True):
        from . import lib
        self.lib = lib
        self.ctx = lib.Context(device)
        self.log_n = log_n
        t0 = time.time()
        self.pk = self.ctx.pk_synthetic(log_n, setup_seed)
        self.self_check = self_check
        log.info("proving key for 2^%d constraints resident on GPU %d (%.1f s)", log_n, device, time.time() - t0)

    @property
    def vk(self) -> bytes:
        return self.pk.vk

    def final_proof(self, batch_id: str, recursive_proof: str, aggregator_addr: str):
        w0 = _fr_from_hash("w0", recursive_proof, aggregator_addr)
        w1 = _fr_from_hash("w1", recursive_proof, aggregator_addr)
        r = _fr_from_hash("r", batch_id)
        s = _fr_from_hash("s", batch_id)
        witness = self.lib.synthetic_witness(self.log_n, w0, w1)
        proof, _, _, timings = self.ctx.groth16_prove(self.pk, witness, r, s)
        public_input = witness[32:64]
        if self.self_check and not self.lib.groth16_verify(self.pk.vk, proof, public_input):
            raise RuntimeError("pairing check failed on a freshly generated proof")
        proof_json, input_json = proof_format.proof_to_json(proof), proof_format.public_input_to_json(public_input)
        proof_format.validate_final_proof(proof_json, input_json)     # what leaves is what verify_batches will accept
        return proof_json, input_json, timings


class BatchLedger:
    """Per-batch bookkeeping across the four steps.  The client carries
    task_id / chunk_count / batch_data from one response into the next request
    (provider.rs:340-377) and, after a restart, re-enters at whatever step it had persisted with the
    same batch_id (provider.rs:245-274).  So: a step for an unknown batch is accepted (this server may
    have restarted too), a step for a known batch must agree with what was handed out earlier, and
    every step is idempotent.  Bounded LRU; thread-safe (several operator streams share it)."""

    STEPS = ("chunks", "chunk_proofs", "aggregated", "final")

    def __init__(self, capacity=1024):
        self._lock = threading.Lock()
        self._rows = OrderedDict()
        self._capacity = capacity

    def record(self, batch_id: str, step: str, **facts):
        """-> None when consistent, else a message naming the mismatch."""
        with self._lock:
            row = self._rows.get(batch_id)
            if row is None:
                row = self._rows[batch_id] = {"steps": {}, "facts": {}}
            self._rows.move_to_end(batch_id)
            for k, v in facts.items():
                old = row["facts"].get(k)
                if old is not None and old != v:
                    return "batch %s: %s was %r, request says %r" % (batch_id, k, old, v)
            row["facts"].update(facts)
            row["steps"][step] = row["steps"].get(step, 0) + 1
            while len(self._rows) > self._capacity:
                self._rows.popitem(last=False)
        return None

    def view(self, batch_id: str):
        with self._lock:
            row = self._rows.get(batch_id)
            return None if row is None else {"steps": dict(row["steps"]), "facts": dict(row["facts"])}

    def __len__(self):
        with self._lock:
            return len(self._rows)


class ProverServicer:
    """the stream handler; `backend` only needs .final_proof(batch_id, recursive_proof, aggregator_addr)."""

    def __init__(self, backend, prover_name="zkb-b200", cache_size=64):
        self.backend = backend
        self.ledger = BatchLedger()
        self.queued = 0                           # GenFinalProof requests waiting for / on the GPU (all streams)
        self._count_lock = threading.Lock()
        self.prover_name = prover_name
        self.prover_id = hashlib.sha256(("%s-%d" % (prover_name, os.getpid())).encode()).hexdigest()[:16]
        self._lock = threading.Lock()            # one proof at a time per server
        self._cache = OrderedDict()
        self._cache_size = cache_size
        self.status = pb.STATUS_IDLE
        self.current_id, self.current_start = "", 0
        self.last_id, self.last_end = "", 0
        self.timings = []

    # ---- stubs for the out-of-scope zkVM steps (deterministic in batch_id / block numbers)
    def _gen_batch_chunks(self, req):
        blocks = list(req.batch.block_number)
        task_id = "%010d" % (blocks[0] if blocks else 0)
        batch_data = json.dumps({"blocks": blocks, "chain_id": req.chain_id, "program": req.program_name})
        bad = self.ledger.record(req.batch_id, "chunks", task_id=task_id, chunk_count=2, batch_data=batch_data)
        if bad:        # the same batch_id re-submitted with other blocks: refuse rather than prove the wrong thing
            return pb.GenBatchProofResponse(gen_batch_chunks=pb.GenBatchChunksResult(
                batch_id=req.batch_id, result_code=pb.COMPLETED_ERROR, error_message=bad,
                pre_state_root=bytes(32), post_state_root=bytes(32)))
        res = pb.GenBatchChunksResult(
            batch_id=req.batch_id, task_id=task_id, result_code=pb.COMPLETED_OK, chunk_count=2, batch_data=batch_data,
            pre_state_root=_h("pre", req.chain_id, *blocks), post_state_root=_h("post", req.chain_id, *blocks))
        return pb.GenBatchProofResponse(gen_batch_chunks=res)

    def _gen_chunk_proof(self, req):
        bad = self.ledger.record(req.batch_id, "chunk_proofs", task_id=req.task_id, chunk_count=int(req.chunk_count),
                                 batch_data=req.batch_data)
        if bad:
            return pb.GenBatchProofResponse(gen_chunk_proof=pb.GenChunkProofResult(
                batch_id=req.batch_id, task_id=req.task_id, result_code=pb.COMPLETED_ERROR, error_message=bad))
        n = max(1, int(req.chunk_count))
        proofs = [pb.ChunkProof(chunk_id=i, proof_key="%s-%d" % (req.task_id, i),
                                proof=_h("chunk", req.batch_id, req.batch_data, i).hex()) for i in range(n)]
        res = pb.GenChunkProofResult(batch_id=req.batch_id, task_id=req.task_id, result_code=pb.COMPLETED_OK,
                                     batch_proof_result=pb.BatchProofResult(task_id=req.task_id, chunk_proofs=proofs))
        return pb.GenBatchProofResponse(gen_chunk_proof=res)

    def _gen_aggregated(self, req):
        self.ledger.record(req.batch_id, "aggregated")
        return pb.GenAggregatedProofResponse(
            batch_id=req.batch_id, result_code=pb.COMPLETED_OK,
            result_string=_h("agg", req.recursive_proof_1, req.recursive_proof_2).hex())

    # ---- the Groth16 step
    def _gen_final(self, req, request_id):
        if req.curve_name.upper() != "BN128":       # CURVE_TYPE default, reference src/config/env.rs:22
            return pb.GenFinalProofResponse(batch_id=req.batch_id, result_code=pb.COMPLETED_ERROR,
                                            error_message="unsupported curve %r (only BN128)" % req.curve_name)
        key = (req.batch_id, req.recursive_proof, req.aggregator_addr)
        # no consistency facts here: a re-aggregated batch may legitimately carry another recursive
        # proof, and an error would make the client resend the same request forever (provider.rs:511-516)
        self.ledger.record(req.batch_id, "final")
        with self._count_lock:
            self.queued += 1                      # streams queue here: one proof at a time on the GPU
        try:
            return self._gen_final_locked(req, request_id, key)
        finally:
            with self._count_lock:
                self.queued -= 1

    def _gen_final_locked(self, req, request_id, key):
        with self._lock:
            hit = self._cache.get(key)
            if hit is None:
                self.status, self.current_id, self.current_start = pb.STATUS_COMPUTING, request_id, int(time.time())
                try:
                    proof_json, input_json, timings = self.backend.final_proof(*key)
                    hit = (proof_json, input_json)
                    self.timings.append(timings)
                    self._cache[key] = hit
                    while len(self._cache) > self._cache_size:
                        self._cache.popitem(last=False)
                except Exception as ex:            # never drop the stream: report and let the client retry
                    log.exception("GenFinalProof failed")
                    return pb.GenFinalProofResponse(batch_id=req.batch_id, result_code=pb.COMPLETED_ERROR,
                                                    error_message="%s: %s" % (type(ex).__name__, ex))
                finally:
                    self.status, self.last_id, self.last_end = pb.STATUS_IDLE, request_id, int(time.time())
                    self.current_id, self.current_start = "", 0
        return pb.GenFinalProofResponse(batch_id=req.batch_id, result_code=pb.COMPLETED_OK, result_string="",
                                        final_proof=pb.FinalProof(proof=hit[0], public_input=hit[1]))

    def _get_status(self, request_id):
        try:
            import psutil
            vm = psutil.virtual_memory()
            total, free = int(vm.total), int(vm.available)
        except Exception:
            total = free = 0
        return pb.GetStatusResponse(
            id=request_id, result_code=0, status=self.status,
            prover_status=pb.ProverStatus(
                last_computed_request_id=self.last_id, last_computed_end_time=self.last_end,
                current_computing_request_id=self.current_id, current_computing_start_time=self.current_start,
                version_proto=VERSION_PROTO, version_server=VERSION_SERVER, prover_name=self.prover_name,
                prover_id=self.prover_id, number_of_cores=os.cpu_count() or 0, total_memory=total, free_memory=free))

    def handle(self, request):
        """one request -> one response of the same variant (never raises)."""
        kind = request.WhichOneof("request_type")
        resp = pb.ProverResponse(id=request.id)
        try:
            if kind == "get_status":
                resp.get_status.CopyFrom(self._get_status(request.id))
            elif kind == "gen_batch_proof":
                step = request.gen_batch_proof.WhichOneof("step")
                if step == "gen_batch_chunks":
                    resp.gen_batch_proof.CopyFrom(self._gen_batch_chunks(request.gen_batch_proof.gen_batch_chunks))
                elif step == "gen_chunk_proof":
                    resp.gen_batch_proof.CopyFrom(self._gen_chunk_proof(request.gen_batch_proof.gen_chunk_proof))
                else:
                    resp.gen_batch_proof.CopyFrom(pb.GenBatchProofResponse(gen_batch_chunks=pb.GenBatchChunksResult(
                        result_code=pb.COMPLETED_ERROR, error_message="empty GenBatchProofRequest",
                        pre_state_root=bytes(32), post_state_root=bytes(32))))
            elif kind == "gen_aggregated_proof":
                resp.gen_aggregated_proof.CopyFrom(self._gen_aggregated(request.gen_aggregated_proof))
            elif kind == "gen_final_proof":
                resp.gen_final_proof.CopyFrom(self._gen_final(request.gen_final_proof, request.id))
            else:
                resp.get_status.CopyFrom(pb.GetStatusResponse(id=request.id, result_code=1, status=self.status,
                                                              error_message="empty ProverRequest"))
        except Exception as ex:                    # belt and braces: a handler bug must not kill the stream
            log.exception("handler error")
            if kind == "gen_final_proof":
                resp.gen_final_proof.CopyFrom(pb.GenFinalProofResponse(result_code=pb.COMPLETED_ERROR, error_message=str(ex)))
            elif kind == "gen_aggregated_proof":
                resp.gen_aggregated_proof.CopyFrom(pb.GenAggregatedProofResponse(result_code=pb.COMPLETED_ERROR, error_message=str(ex)))
            else:
                resp.get_status.CopyFrom(pb.GetStatusResponse(id=request.id, result_code=1, error_message=str(ex)))
        return resp

    def prover_stream(self, request_iterator, context):
        for request in request_iterator:
            yield self.handle(request)


def make_server(servicer: ProverServicer, address: str = "127.0.0.1:50061", max_workers: int = 8):
    handler = grpc.method_handlers_generic_handler(pb.SERVICE, {
        pb.METHOD: grpc.stream_stream_rpc_method_handler(
            servicer.prover_stream, request_deserializer=pb.ProverRequest.FromString,
            response_serializer=pb.ProverResponse.SerializeToString)})
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers),
                         options=[("grpc.max_receive_message_length", 256 << 20), ("grpc.max_send_message_length", 256 << 20)])
    server.add_generic_rpc_handlers((handler,))
    port = server.add_insecure_port(address)       # the client speaks plaintext http:// (provider.rs:606)
    return server, port


def main(argv=None):
    ap = argparse.ArgumentParser(description="zkb ProverService (drop-in for eigen-zeth's PROVER_ADDR)")
    ap.add_argument("--address", default="0.0.0.0:50061")
    ap.add_argument("--log-n", type=int, default=16, help="log2 of the synthetic circuit's QAP domain")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--setup-seed", type=lambda s: int(s, 0), default=0x5EED0004)
    ap.add_argument("--vk-out", default=None, help="write the verifying key (hex, include/zkb.h layout) here")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    backend = GpuGroth16Backend(args.log_n, args.setup_seed, args.device)
    if args.vk_out:
        with open(args.vk_out, "w") as f:
            f.write(backend.vk.hex())
    server, port = make_server(ProverServicer(backend), args.address)
    server.start()
    log.info("ProverService listening on %s (port %d)", args.address, port)
    server.wait_for_termination()


if __name__ == "__main__":
    main()
