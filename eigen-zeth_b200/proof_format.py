"""proof.json / public_input.json writer and reader — the output layout eigen-zeth consumes.

Normative layout: reference proof/proof.json:1-22 and proof/public_input.json:1-3 (read at
src/settlement/worker.rs:56-59).  The reader restates `parse_proof` / `parse_public_input`
(src/settlement/ethereum/mod.rs:445-481): coordinates are decimal *strings*, pi_b is copied
unswapped ([c0, c1]), `protocol` / `curve` are ignored, only element 0 of the input array is read,
and any missing key or non-string value is the error "invalid json data".
`ProofResult` mirrors src/db/mod.rs:63-71 (serde_json of [u8; 32] is an array of 32 numbers).
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ERR = "invalid json data"


def _le(b: bytes) -> int:
    return int.from_bytes(b, "little")


def proof_to_json(proof: bytes) -> str:
    """256-byte proof (A 64 | B 128 | C 64, include/zkb.h) -> proof.json text."""
    if len(proof) != 256:
        raise ValueError("proof must be 256 bytes")
    c = [_le(proof[i:i + 32]) for i in range(0, 256, 32)]
    if any(v >= Q for v in c):
        raise ValueError("coordinate not reduced modulo q")
    doc = {
        "pi_a": {"x": str(c[0]), "y": str(c[1])},
        "pi_b": {"x": [str(c[2]), str(c[3])], "y": [str(c[4]), str(c[5])]},
        "pi_c": {"x": str(c[6]), "y": str(c[7])},
        "protocol": "groth16",
        "curve": "BN128",
    }
    return json.dumps(doc, indent=4)


def public_input_to_json(x: bytes) -> str:
    """32-byte canonical Fr -> public_input.json text (array of exactly one decimal string)."""
    v = _le(x)
    if len(x) != 32 or v >= R:
        raise ValueError("public input must be a canonical 32-byte Fr element")
    return json.dumps([str(v)], indent=4)


def _u256_from_dec_str(v) -> int:
    # U256::from_dec_str: ASCII digits only, value < 2^256; `.as_str()` fails on non-strings
    if not isinstance(v, str):
        raise ValueError(ERR)
    if not v or not v.isascii() or not v.isdigit():
        raise ValueError("invalid character in decimal string")
    n = int(v, 10)
    if n >> 256:
        raise ValueError("decimal string overflows U256")
    return n


def _get(v, *path):
    for p in path:
        try:
            v = v[p]
        except (KeyError, IndexError, TypeError):
            raise ValueError(ERR)
    return v


def parse_proof(json_str: str):
    """-> ((ax, ay), ((bx0, bx1), (by0, by1)), (cx, cy)) exactly as the Solidity Proof struct gets them."""
    v = json.loads(json_str)
    a = (_u256_from_dec_str(_get(v, "pi_a", "x")), _u256_from_dec_str(_get(v, "pi_a", "y")))
    b = ((_u256_from_dec_str(_get(v, "pi_b", "x", 0)), _u256_from_dec_str(_get(v, "pi_b", "x", 1))),
         (_u256_from_dec_str(_get(v, "pi_b", "y", 0)), _u256_from_dec_str(_get(v, "pi_b", "y", 1))))
    c = (_u256_from_dec_str(_get(v, "pi_c", "x")), _u256_from_dec_str(_get(v, "pi_c", "y")))
    return a, b, c


def parse_public_input(json_str: str):
    v = json.loads(json_str)
    return [_u256_from_dec_str(_get(v, 0))]


def proof_from_json(json_str: str) -> bytes:
    return _coords_to_bytes(parse_proof(json_str))


def public_input_from_json(json_str: str) -> bytes:
    (x,) = parse_public_input(json_str)
    if x >= R:
        raise ValueError("public input not reduced modulo r")   # NewStateRootNotInsidePrime on-chain
    return x.to_bytes(32, "little")


def validate_final_proof(proof_json: str, public_input_json: str) -> None:
    """Strict gate in front of `verify_batches`.  The reference's own readers
    (ethereum/mod.rs:445-481) accept any U256 and only element 0 of the input array; the contract then
    reverts with InvalidProof / NewStateRootNotInsidePrime (interfaces/zkvm.rs:82-130) after the L1
    transaction was sent.  This check raises ValueError before that: the JSON must parse the way
    parse_proof parses it, hold exactly one public input (`uint256[1]`), every coordinate < q, the
    input < r, pi_a / pi_c on the curve, pi_b on the twist and in the order-r subgroup, and no
    element at infinity.  The group checks run in libzkb.so (zkb_proof_check, host code)."""
    v = json.loads(public_input_json)
    if not isinstance(v, list) or len(v) != 1:
        raise ValueError("public_input.json must be an array of exactly one decimal string")
    from . import lib
    reasons = lib.proof_check(_coords_to_bytes(parse_proof(proof_json), clamp=True),
                              min(parse_public_input(public_input_json)[0], (1 << 256) - 1).to_bytes(32, "little"))
    if reasons:
        raise ValueError("malformed final proof: " + "; ".join(reasons))


def _coords_to_bytes(parsed, clamp=False) -> bytes:
    a, b, c = parsed
    vals = [a[0], a[1], b[0][0], b[0][1], b[1][0], b[1][1], c[0], c[1]]
    if not clamp and any(x >= Q for x in vals):
        raise ValueError("coordinate not reduced modulo q")
    return b"".join(x.to_bytes(32, "little") for x in vals)


def vk_to_json(vk: bytes) -> str:
    """576-byte verifying key (include/zkb.h layout) -> the JSON a Groth16Verifier.sol generator
    takes.  G2 coordinates are written [c0, c1] — the order the contract behind
    reference contracts/EigenZkVM.json stores them in and the order of
    `pi_b` in proof.json; its pairing library performs the swap the 0x08 precompile needs."""
    if len(vk) != 576:
        raise ValueError("verifying key must be 576 bytes")
    c = [_le(vk[i:i + 32]) for i in range(0, 576, 32)]
    if any(v >= Q for v in c):
        raise ValueError("coordinate not reduced modulo q")
    g1 = lambda i: [str(c[i]), str(c[i + 1])]
    g2 = lambda i: [[str(c[i]), str(c[i + 1])], [str(c[i + 2]), str(c[i + 3])]]
    return json.dumps({"protocol": "groth16", "curve": "BN128", "nPublic": 1, "vk_alpha_1": g1(0), "vk_beta_2": g2(2),
                       "vk_gamma_2": g2(6), "vk_delta_2": g2(10), "IC": [g1(14), g1(16)]}, indent=1)


def vk_from_json(json_str: str) -> bytes:
    v = json.loads(json_str)
    flat = list(v["vk_alpha_1"]) + [x for pair in v["vk_beta_2"] for x in pair] + [x for pair in v["vk_gamma_2"] for x in pair] \
        + [x for pair in v["vk_delta_2"] for x in pair] + [x for pt in v["IC"] for x in pt]
    vals = [_u256_from_dec_str(x) for x in flat]
    if len(vals) != 18 or any(x >= Q for x in vals):
        raise ValueError("malformed verifying key")
    return b"".join(x.to_bytes(32, "little") for x in vals)


@dataclass
class ProofResult:
    """src/db/mod.rs:63-71; stored under BATCH_PROOF_{n} (src/settlement/worker.rs:179-182)."""
    block_number: int = 0
    proof: str = ""
    public_input: str = ""
    pre_state_root: bytes = field(default_factory=lambda: bytes(32))
    post_state_root: bytes = field(default_factory=lambda: bytes(32))

    def to_json(self) -> str:
        if len(self.pre_state_root) != 32 or len(self.post_state_root) != 32:
            raise ValueError("state roots must be 32 bytes")       # provider.rs:323-324
        return json.dumps({"block_number": self.block_number, "proof": self.proof,
                           "public_input": self.public_input, "pre_state_root": list(self.pre_state_root),
                           "post_state_root": list(self.post_state_root)})

    @staticmethod
    def from_json(s: str) -> "ProofResult":
        d = json.loads(s)
        return ProofResult(int(d["block_number"]), d["proof"], d["public_input"], bytes(d["pre_state_root"]),
                           bytes(d["post_state_root"]))


PREFIX_BATCH_PROOF = "BATCH_PROOF_"          # src/db/mod.rs:42-45


def batch_proof_key(block_no: int) -> bytes:
    """the rollup-DB key a finished proof is stored under (src/settlement/worker.rs:179-182) and
    looked up by (src/custom_reth/eigen.rs:77-82)."""
    return (PREFIX_BATCH_PROOF + str(int(block_no))).encode()


@dataclass
class BatchProofInfo:
    """The `eigenrpc_getBatchProof` view of a stored ProofResult (src/custom_reth/eigen.rs:76-97,
    struct at :108-117): proof / public_input verbatim, state roots as "0x" + lower-case hex."""
    block_number: int
    proof: str
    public_input: str
    pre_state_root: str
    post_state_root: str

    def to_json(self) -> str:
        return json.dumps({"block_number": self.block_number, "proof": self.proof, "public_input": self.public_input,
                           "pre_state_root": self.pre_state_root, "post_state_root": self.post_state_root})


def get_batch_proof(db_get, block_no: int):
    """eigenrpc_getBatchProof over any `db_get(key: bytes) -> bytes | None` (the reference's
    Database::get, src/db/mod.rs:12-16): None when no proof is stored for the block, else the
    BatchProofInfo the RPC returns.  A stored value that is not a ProofResult raises ValueError (the
    reference answers with an internal RPC error, eigen.rs:84-85)."""
    raw = db_get(batch_proof_key(block_no))
    if raw is None:
        return None
    try:
        pr = ProofResult.from_json(raw.decode() if isinstance(raw, (bytes, bytearray)) else raw)
    except (KeyError, TypeError, json.JSONDecodeError) as ex:
        raise ValueError("stored batch proof is not a ProofResult: %s" % ex)
    if len(pr.pre_state_root) != 32 or len(pr.post_state_root) != 32:
        raise ValueError("stored state roots must be 32 bytes")
    return BatchProofInfo(pr.block_number, pr.proof, pr.public_input, "0x" + pr.pre_state_root.hex(),
                          "0x" + pr.post_state_root.hex())
