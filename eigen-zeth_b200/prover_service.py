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
import hmac
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
DEFAULT_AGGREGATOR = "479881985774944702531460751064278034642760119942"    # reference src/commands/run.rs:76


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


def _blinding(secret: bytes, tag: str, batch_id: str, recursive_proof: str, aggregator_addr: str) -> bytes:
    """Groth16 blinding scalar: HMAC-SHA256 under a server-side secret over EVERYTHING the proof is a
    function of (batch id, witness source, aggregator).  Never derived from public request fields
    alone — anyone who can recompute r, s strips the blinding (A - r*delta) and zero-knowledge is
    gone; and one (r, s) is never reused on two different witnesses."""
    m = hmac.new(secret, digestmod=hashlib.sha256)
    for part in (tag, batch_id, recursive_proof, aggregator_addr):
        b = part.encode()
        m.update(len(b).to_bytes(8, "little"))
        m.update(b)
    return (int.from_bytes(m.digest(), "little") & ((1 << 253) - 1)).to_bytes(32, "little")


def load_or_create_secret(path: str | None) -> bytes:
    """32 random bytes; persisted (mode 0600) when a path is given so that a restarted server replays
    a batch with the same blinding (idempotent GenFinalProof, provider.rs:245-274)."""
    if path and os.path.exists(path):
        with open(path, "rb") as f:
            sec = f.read()
        if len(sec) < 32:
            raise ValueError("secret file %s holds fewer than 32 bytes" % path)
        return sec
    sec = os.urandom(32)
    if path:
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_EXCL, 0o600)
        with os.fdopen(fd, "wb") as f:
            f.write(sec)
    return sec


class _PinnedWitness:
    """witness staging in page-locked host memory (zkb_host_alloc): the H2D copy inside
    zkb_groth16_prove then runs at PCIe speed instead of through a pageable bounce buffer."""

    def __init__(self, lib, nbytes):
        self.lib, self.nbytes = lib, nbytes
        self.ptr = lib.host_alloc(nbytes)

    def free(self):
        if self.ptr:
            self.lib.host_free(self.ptr)
            self.ptr = 0


class GpuGroth16Backend:
    """GenFinalProof on the B200s named in `devices`.  No CPU fallback: constructing it without the
    CUDA library or an sm_100 GPU raises.

    The synthetic circuit's setup is TEST-ONLY: its toxic waste comes from `setup_seed`, so whoever
    knows the seed can forge proofs for the exported verifying key.  A production deployment loads a
    ceremony key through zkb_pk_load instead."""

    def __init__(self, log_n: int = 16, setup_seed: int | None = None, device: int = 0, self_check: bool = True,
                 secret: bytes | None = None, devices=None):
        from . import lib
        self.lib = lib
        self.log_n = log_n
        self.self_check = self_check
        self.secret = secret if secret is not None else os.urandom(32)
        if setup_seed is None:                                     # never logged, never derivable
            setup_seed = int.from_bytes(os.urandom(8), "little")
        devices = list(devices) if devices else [device]
        t0 = time.time()
        if len(devices) == 1:
            self.group = None
            self.ctx = lib.Context(devices[0])
            self.pk = self.ctx.pk_synthetic(log_n, setup_seed)
            self._vk = self.pk.vk
        else:
            self.group = lib.DeviceGroup(devices)                   # one context + NCCL rank per GPU
            self.group.pk_synthetic(log_n, setup_seed)
            self._vk = self.group.vk
        # witness staging: a small pool of pinned buffers, each either free, being filled by a
        # generator thread for one (recursive_proof, aggregator) key, or in use by a running proof
        self._slots = [{"buf": _PinnedWitness(lib, 32 << log_n), "key": None, "thread": None, "in_use": False, "age": 0}
                       for _ in range(2)]
        self._age = 0
        self._pre_lock = threading.Lock()
        log.info("proving key for 2^%d constraints resident on GPU(s) %s (%.1f s)", log_n, devices, time.time() - t0)

    @property
    def vk(self) -> bytes:
        return self._vk

    # ---- witness: generated ahead of the request whenever its inputs are already known
    def _witness_seeds(self, recursive_proof, aggregator_addr):
        return _fr_from_hash("w0", recursive_proof, aggregator_addr), _fr_from_hash("w1", recursive_proof, aggregator_addr)

    def _claim_slot(self, key):
        """under the lock: the slot already holding `key`, else a reusable one (idle generator, no proof
        running on it) re-targeted to `key` with its generator started; None when every slot is busy"""
        for sl in self._slots:
            if sl["key"] == key:
                return sl
        idle = [sl for sl in self._slots if not sl["in_use"] and (sl["thread"] is None or not sl["thread"].is_alive())]
        if not idle:
            return None
        sl = min(idle, key=lambda x: (x["key"] is not None, x["age"]))       # a free one first, else the oldest
        w0, w1 = self._witness_seeds(*key)
        sl["key"] = key
        self._age += 1
        sl["age"] = self._age
        sl["thread"] = threading.Thread(target=self.lib.synthetic_witness_into, args=(self.log_n, w0, w1, sl["buf"].ptr),
                                        daemon=True)
        sl["thread"].start()
        return sl

    def prefetch(self, recursive_proof: str, aggregator_addr: str):
        """Start generating the witness for (recursive_proof, aggregator_addr) in the background (no-op
        when it is already under way or every staging buffer is busy).  The servicer calls this as
        soon as the string is known — at the chunk step, three client sleeps (provider.rs:542) ahead
        of the GenFinalProof request that carries it."""
        with self._pre_lock:
            self._claim_slot((recursive_proof, aggregator_addr))

    def _witness(self, recursive_proof, aggregator_addr):
        """-> the slot holding this witness, generator finished, marked in use"""
        key = (recursive_proof, aggregator_addr)
        while True:
            with self._pre_lock:
                sl = self._claim_slot(key)
                if sl is not None:
                    sl["in_use"] = True
                    break
            time.sleep(0.001)            # every buffer is busy with another stream's proof: wait for one
        sl["thread"].join()
        return sl

    def _release(self, sl):
        with self._pre_lock:
            sl["in_use"] = False
            sl["key"] = None             # a replayed batch is answered from the servicer's cache, not from here

    def final_proof(self, batch_id: str, recursive_proof: str, aggregator_addr: str):
        t = {}
        t0 = time.perf_counter()
        slot = self._witness(recursive_proof, aggregator_addr)
        try:
            return self._final_proof_on(slot["buf"], t, t0, batch_id, recursive_proof, aggregator_addr)
        finally:
            self._release(slot)

    def _final_proof_on(self, wbuf, t, t0, batch_id, recursive_proof, aggregator_addr):
        t["witness_wait_ms"] = 1e3 * (time.perf_counter() - t0)
        r = _blinding(self.secret, "r", batch_id, recursive_proof, aggregator_addr)
        s = _blinding(self.secret, "s", batch_id, recursive_proof, aggregator_addr)
        t0 = time.perf_counter()
        if self.group is None:
            proof, _, _, dev_t = self.ctx.groth16_prove(self.pk, wbuf.ptr, r, s)
        else:
            proof, dev_t = self.group.groth16_prove(wbuf.ptr, r, s)
        t["prove_call_ms"] = 1e3 * (time.perf_counter() - t0)
        public_input = self.lib.host_read(wbuf.ptr + 32, 32)
        t0 = time.perf_counter()
        if self.self_check and not self.lib.groth16_verify(self._vk, proof, public_input):
            raise RuntimeError("pairing check failed on a freshly generated proof")
        t["pairing_check_ms"] = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        proof_json, input_json = proof_format.proof_to_json(proof), proof_format.public_input_to_json(public_input)
        proof_format.validate_final_proof(proof_json, input_json)     # what leaves is what verify_batches will accept
        t["format_and_check_ms"] = 1e3 * (time.perf_counter() - t0)
        t.update(dev_t)
        return proof_json, input_json, t

    def close(self):
        for sl in self._slots:
            if sl["thread"] is not None:
                sl["thread"].join()
            sl["buf"].free()


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
        self.last_aggregator = DEFAULT_AGGREGATOR

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
        # Everything the later steps hand out is a deterministic function of what is known now, so the
        # string that will come back as GenFinalProofRequest.recursive_proof three steps (>= 3 s of
        # client sleeps, provider.rs:542) from now is already known: start the witness generation of
        # the final circuit here and it is off the GenFinalProof critical path whatever its size.
        self._prefetch(self._aggregated(self._chunk_proof_strings(req.batch_id, batch_data, 2)))
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
        strings = self._chunk_proof_strings(req.batch_id, req.batch_data, n)
        proofs = [pb.ChunkProof(chunk_id=i, proof_key="%s-%d" % (req.task_id, i), proof=strings[i]) for i in range(n)]
        res = pb.GenChunkProofResult(batch_id=req.batch_id, task_id=req.task_id, result_code=pb.COMPLETED_OK,
                                     batch_proof_result=pb.BatchProofResult(task_id=req.task_id, chunk_proofs=proofs))
        return pb.GenBatchProofResponse(gen_chunk_proof=res)

    @staticmethod
    def _chunk_proof_strings(batch_id, batch_data, n):
        return [_h("chunk", batch_id, batch_data, i).hex() for i in range(n)]

    @staticmethod
    def _aggregated(strings):
        """what GenAggregatedProof answers when the client sends the first and last chunk proof
        (provider.rs:386-387, :422-433)"""
        return _h("agg", strings[0], strings[-1]).hex()

    def _prefetch(self, recursive_proof):
        """start building the witness of the final circuit for `recursive_proof`, with the aggregator
        address this operator used last (a constant of the operator, commands/run.rs:76)"""
        pre = getattr(self.backend, "prefetch", None)
        if pre is not None and self.last_aggregator is not None:
            try:
                pre(recursive_proof, self.last_aggregator)
            except Exception:
                log.exception("witness prefetch failed (the final step will generate it)")

    def _gen_aggregated(self, req):
        self.ledger.record(req.batch_id, "aggregated")
        result = _h("agg", req.recursive_proof_1, req.recursive_proof_2).hex()
        # the client sends `result` back as GenFinalProofRequest.recursive_proof after its 1 s step
        # sleep (provider.rs:472-483, :542); normally this witness was already started at the chunk
        # step (no-op then), a client that aggregates other proofs gets it started here
        self._prefetch(result)
        return pb.GenAggregatedProofResponse(batch_id=req.batch_id, result_code=pb.COMPLETED_OK, result_string=result)

    # ---- the Groth16 step
    def _gen_final(self, req, request_id):
        if req.curve_name.upper() != "BN128":       # CURVE_TYPE default, reference src/config/env.rs:22
            return pb.GenFinalProofResponse(batch_id=req.batch_id, result_code=pb.COMPLETED_ERROR,
                                            error_message="unsupported curve %r (only BN128)" % req.curve_name)
        key = (req.batch_id, req.recursive_proof, req.aggregator_addr)
        self.last_aggregator = req.aggregator_addr
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
            elif kind == "gen_batch_proof":
                # the client forwards only same-variant responses to ProverChannel (provider.rs:646-668);
                # a GetStatus answer here would leave it blocked on response_receiver.recv() forever
                if request.gen_batch_proof.WhichOneof("step") == "gen_chunk_proof":
                    resp.gen_batch_proof.CopyFrom(pb.GenBatchProofResponse(gen_chunk_proof=pb.GenChunkProofResult(
                        batch_id=request.gen_batch_proof.gen_chunk_proof.batch_id, result_code=pb.COMPLETED_ERROR,
                        error_message=str(ex))))
                else:
                    resp.gen_batch_proof.CopyFrom(pb.GenBatchProofResponse(gen_batch_chunks=pb.GenBatchChunksResult(
                        batch_id=request.gen_batch_proof.gen_batch_chunks.batch_id, result_code=pb.COMPLETED_ERROR,
                        error_message=str(ex), pre_state_root=bytes(32), post_state_root=bytes(32))))
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
    ap.add_argument("--devices", default=None, help="comma-separated GPU list, e.g. 0,1,2,3,4,5,6,7: one proof is "
                    "spread over all of them (four-step NTTs + every MSM sharded by point range)")
    ap.add_argument("--insecure-synthetic-setup", action="store_true",
                    help="REQUIRED to serve the synthetic test circuit: its trusted setup is generated locally, "
                         "so whoever holds the seed can forge proofs for the exported verifying key")
    ap.add_argument("--setup-seed", type=lambda s: int(s, 0), default=None,
                    help="reproducible (publicly known!) toxic waste for tests; default: fresh random, never logged")
    ap.add_argument("--secret-file", default=None, help="32-byte blinding secret (created if absent, mode 0600)")
    ap.add_argument("--vk-out", default=None, help="write the verifying key (hex, include/zkb.h layout) here")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    if not args.insecure_synthetic_setup:
        ap.error("this build serves only the synthetic test circuit, whose trusted setup is NOT a ceremony: "
                 "pass --insecure-synthetic-setup to acknowledge that its proofs are forgeable by the key holder "
                 "(production keys enter through zkb_pk_load, see INTEGRATION.md)")
    log.warning("TEST-ONLY trusted setup: the synthetic circuit's toxic waste is generated on this host; "
                "do not point a production verifier at the exported verifying key")
    devices = [int(x) for x in args.devices.split(",")] if args.devices else [args.device]
    backend = GpuGroth16Backend(args.log_n, args.setup_seed, devices=devices, secret=load_or_create_secret(args.secret_file))
    if args.vk_out:
        with open(args.vk_out, "w") as f:
            f.write(backend.vk.hex())
    server, port = make_server(ProverServicer(backend), args.address)
    server.start()
    log.info("ProverService listening on %s (port %d), GPUs %s", args.address, port, devices)
    server.wait_for_termination()


if __name__ == "__main__":
    main()
