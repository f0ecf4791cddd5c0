"""CPU-only checks of the product: the C-ABI library loads and exports every symbol include/zkb.h
declares; host-side logic (witness generation, pairing verifier, proof.json writer/reader, error
behaviour) matches the oracles; no compute entry point works without a GPU (no CPU fallback)."""
import json
import os
import re

import pytest

import oracle_a as o
import oracle_b as ob
from helpers import fr_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_fixtures.json")))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "zkb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(zkb):
    lib = zkb.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(zkb.lib.EXPORTS) == declared
    assert "sm_100a" in zkb.version()


def test_library_is_sm100a_only():
    """the shipped cubin targets sm_100a and nothing else (no multi-arch fallback)."""
    import subprocess
    so = os.path.join(ROOT, "eigen-zeth_b200", "libzkb.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(zkb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(zkb.ZkbError) as e:
        zkb.Context(0)
    assert e.value.code == zkb.lib.ZKB_ENOGPU


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "eigen-zeth_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_a" not in text and "oracle_b" not in text and "oracle/" not in text, f


def test_host_witness_matches_oracles(zkb):
    rng = o.Xoshiro256ss(4)
    for log_n in (3, 6, 10):
        w0, w1 = fr_bytes([rng.fr()]), fr_bytes([rng.fr()])
        got = zkb.synthetic_witness(log_n, w0, w1)
        assert got == ob.synthetic_witness(log_n, w0, w1)
        z = [int.from_bytes(got[i:i + 32], "little") for i in range(0, len(got), 32)]
        assert o.r1cs_is_satisfied(o.synthetic_r1cs(log_n), z) if log_n <= 6 else True
    with pytest.raises(zkb.ZkbError):
        zkb.synthetic_witness(4, int(o.R).to_bytes(32, "little"), bytes(32))


def test_host_pairing_verifier_against_oracle_proofs(zkb):
    log_n = 4
    pk = ob.ProvingKey(log_n, 4242)
    rng = o.Xoshiro256ss(5)
    w0, w1, r, s = (fr_bytes([rng.fr()]) for _ in range(4))
    wit = ob.synthetic_witness(log_n, w0, w1)
    proof, _, _, _ = pk.prove(wit, r, s)
    vk = pk.vk()
    assert zkb.groth16_verify(vk, proof, wit[32:64])
    bad = bytearray(wit[32:64])
    bad[0] ^= 1
    assert not zkb.groth16_verify(vk, proof, bytes(bad))
    # a well-formed proof of another circuit (the reference fixture) is rejected, not malformed
    A, B, C = o.parse_proof(GOLDEN["proof_json"])
    other = o.g1_to_bytes(A) + o.g2_to_bytes(B) + o.g1_to_bytes(C)
    assert not zkb.groth16_verify(vk, other, wit[32:64])
    # malformed: off-curve point, non-canonical coordinate, public input >= r
    off = bytearray(proof)
    off[0] ^= 1
    with pytest.raises(zkb.ZkbError):
        zkb.groth16_verify(vk, bytes(off), wit[32:64])
    with pytest.raises(zkb.ZkbError):
        zkb.groth16_verify(vk, int(o.Q).to_bytes(32, "little") + proof[32:], wit[32:64])
    with pytest.raises(zkb.ZkbError):
        zkb.groth16_verify(vk, proof, int(o.R).to_bytes(32, "little"))
    # swapping c0/c1 of B leaves the twist
    sw = proof[:64] + proof[96:128] + proof[64:96] + proof[160:192] + proof[128:160] + proof[192:]
    with pytest.raises(zkb.ZkbError):
        zkb.groth16_verify(vk, sw, wit[32:64])


def test_proof_format_matches_reference_layout():
    from eigen_zeth_b200 import proof_format as pf
    # reader == the reference's test_parse_proof known answers
    a, b, c = pf.parse_proof(GOLDEN["test_parse_proof_json"])
    got = [a[0], a[1], b[0][0], b[0][1], b[1][0], b[1][1], c[0], c[1]]
    assert got == [int(v) for v in GOLDEN["test_parse_proof_expected"]]
    assert pf.parse_public_input(GOLDEN["test_parse_public_input_json"]) == [int(GOLDEN["test_parse_public_input_expected"][0])]
    # writer reproduces the reference's proof/proof.json byte for byte from the parsed coordinates
    raw = pf.proof_from_json(GOLDEN["proof_json"])
    assert pf.proof_to_json(raw) == GOLDEN["proof_json"].rstrip("\n")
    assert json.loads(pf.proof_to_json(raw)) == json.loads(GOLDEN["proof_json"])
    pub = pf.public_input_from_json(GOLDEN["public_input_json"])
    assert pf.public_input_to_json(pub) == GOLDEN["public_input_json"].rstrip("\n")
    # compact form (worker.rs:760-761) parses too; whitespace is free-form
    assert pf.parse_proof(GOLDEN["worker_compact_proof_json"]) == o.parse_proof(GOLDEN["worker_compact_proof_json"])
    # error behaviour of parse_proof: missing key / non-string -> "invalid json data"
    for broken in ('{"pi_a":{"x":"1"}}', '{"pi_a":{"x":1,"y":"2"}}', '[]'):
        with pytest.raises(ValueError):
            pf.parse_proof(broken)
    with pytest.raises(ValueError):
        pf.parse_public_input("[]")
    with pytest.raises(ValueError):
        pf.parse_public_input('["12x"]')
    with pytest.raises(ValueError):
        pf.proof_to_json(int(o.Q).to_bytes(32, "little") + bytes(224))
    # verifying-key export for a Groth16Verifier.sol generator, [c0, c1] order
    vk = ob.ProvingKey(3, 5).vk()
    js = pf.vk_to_json(vk)
    assert pf.vk_from_json(js) == vk
    doc = json.loads(js)
    assert doc["vk_beta_2"][0] == [str(int.from_bytes(vk[64:96], "little")), str(int.from_bytes(vk[96:128], "little"))]
    assert len(doc["IC"]) == 2 and doc["nPublic"] == 1
    pr = pf.ProofResult(7, "p", "i", bytes(range(32)), bytes(32))
    assert pf.ProofResult.from_json(pr.to_json()) == pr
    assert json.loads(pr.to_json())["pre_state_root"] == list(range(32))


def _f2_sqrt(a):
    """square root in Fq[u]/(u^2+1), q = 3 mod 4 (test helper)."""
    q = o.Q
    rt = lambda v: (lambda r: r if r * r % q == v % q else None)(pow(v, (q + 1) // 4, q))
    if a[1] == 0:
        r = rt(a[0])
        return (r, 0) if r is not None else (0, rt(-a[0] % q))
    s = rt((a[0] * a[0] + a[1] * a[1]) % q)
    if s is None:
        return None
    inv2 = pow(2, q - 2, q)
    x0 = rt((a[0] + s) * inv2 % q)
    if x0 is None:
        x0 = rt((a[0] - s) * inv2 % q)
    if x0 is None:
        return None
    y = (x0, a[1] * pow(2 * x0, q - 2, q) % q)
    return y if o.f2_sqr(y) == (a[0] % q, a[1] % q) else None


def test_strict_final_proof_validator(zkb):
    """the gate in front of verify_batches.  All three reference fixtures are well-formed
    group elements (Appendix C); every class of malformation is named."""
    from eigen_zeth_b200 import proof_format as pf
    for key in ("proof_json", "test_parse_proof_json", "worker_compact_proof_json"):
        pf.validate_final_proof(GOLDEN[key], GOLDEN["public_input_json"])
    raw = pf.proof_from_json(GOLDEN["proof_json"])
    pub = pf.public_input_from_json(GOLDEN["public_input_json"])
    assert zkb.proof_check(raw, pub) == []
    le = lambda v: int(v).to_bytes(32, "little")
    # pi_a: y + 1 leaves the curve; x >= q is a range error (parse_proof itself would accept both)
    bad = raw[:32] + le(int.from_bytes(raw[32:64], "little") + 1) + raw[64:]
    assert zkb.proof_check(bad, pub) == ["pi_a not on the curve"]
    bad = le(int.from_bytes(raw[:32], "little") + o.Q) + raw[32:]
    assert zkb.proof_check(bad, pub) == ["pi_a coordinate >= q"]
    # pi_b with c0 / c1 swapped (the snarkjs-vs-contract ordering mistake): off the twist
    sw = raw[:64] + raw[96:128] + raw[64:96] + raw[160:192] + raw[128:160] + raw[192:]
    assert zkb.proof_check(sw, pub) == ["pi_b not on the twist"]
    # a twist point outside the order-r subgroup: x = 1 + 0u gives one after clearing nothing
    x = 0
    while True:
        x += 1
        y = _f2_sqrt(o.f2_add(o.f2_mul(o.f2_sqr((x, 0)), (x, 0)), o.TWIST_B))
        if y is not None and not o.g2_in_subgroup(((x, 0), y)):
            break
    cof = raw[:64] + le(x) + le(0) + le(y[0]) + le(y[1]) + raw[192:]
    assert zkb.proof_check(cof, pub) == ["pi_b outside the order-r subgroup"]
    # pi_c at infinity, input >= r, both reported together
    assert zkb.proof_check(raw[:192] + bytes(64), le(o.R)) == ["public input >= r", "proof element is the point at infinity"]
    # the JSON-level gate: two inputs, oversized input, tampered coordinate
    with pytest.raises(ValueError, match="exactly one"):
        pf.validate_final_proof(GOLDEN["proof_json"], '["1","2"]')
    with pytest.raises(ValueError, match="public input >= r"):
        pf.validate_final_proof(GOLDEN["proof_json"], '["%d"]' % o.R)
    doc = json.loads(GOLDEN["proof_json"])
    doc["pi_c"]["y"] = str(int(doc["pi_c"]["y"]) ^ 1)
    with pytest.raises(ValueError, match="pi_c not on the curve"):
        pf.validate_final_proof(json.dumps(doc), GOLDEN["public_input_json"])
    doc["pi_c"]["y"] = str((1 << 256) - 1)
    with pytest.raises(ValueError, match="pi_c coordinate >= q"):
        pf.validate_final_proof(json.dumps(doc), GOLDEN["public_input_json"])


def test_rust_sys_crate_is_in_sync_with_the_header():
    """rust/zkb-sys cannot be built here (no cargo); it must at least bind every symbol of zkb.h."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"])
    assert r.returncode == 0, "run `python tools/gen_rust_sys.py` after changing include/zkb.h"
    text = open(os.path.join(ROOT, "rust", "zkb-sys", "src", "lib.rs")).read()
    for sym in _declared_symbols():
        assert "pub fn %s(" % sym in text, sym
    import re
    hdr = open(os.path.join(ROOT, "include", "zkb.h")).read()
    for name in re.findall(r"typedef struct (zkb_[a-z0-9_]+)", hdr):
        assert "pub struct %s " % name in text, name
    for name in re.findall(r"#define (ZKB_[A-Z0-9_]+) ", hdr):
        assert "pub const %s:" % name in text, name
    assert 'cargo:rustc-link-lib=dylib=zkb' in open(os.path.join(ROOT, "rust", "zkb-sys", "build.rs")).read()


def _build_example(tmp_path, name="prove_cli"):
    import subprocess
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "eigen-zeth_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", name + ".cpp"), "-L" + libdir, "-lzkb",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    return exe


def _zkey_case(tmp_path):
    """a small circuit as the files a circom / snarkjs flow leaves behind: circuit.zkey + witness.wtns"""
    import oracle_a as o
    import oracle_zkey as zkm
    from eigen_zeth_b200 import circom_io as cio
    from helpers import small_circuit
    cs, z = small_circuit(6)
    rng = o.Xoshiro256ss(0x2E71)
    tau, alpha, beta, gamma, delta = [rng.fr() or 1 for _ in range(5)]
    pk, vk = o.groth16_setup(cs, tau, alpha, beta, gamma, delta)
    zpath, wpath = str(tmp_path / "circuit.zkey"), str(tmp_path / "witness.wtns")
    open(zpath, "wb").write(zkm.zkey_bytes(cs, pk, vk, zkm.h_points_snarkjs(tau, delta, pk["n"])))
    open(wpath, "wb").write(cio.write_wtns(z))
    return cs, z, vk, zpath, wpath


def test_zkey_prover_example_parses_before_touching_the_gpu(tmp_path):
    """examples/zkey_prover.cpp (the rapidsnark-style command line over the C ABI): malformed inputs
    exit 1 whatever the machine, a good pair exits 2 on a machine without a GPU"""
    import subprocess
    import torch
    from eigen_zeth_b200 import circom_io as cio
    exe = _build_example(tmp_path, "zkey_prover")
    cs, z, vk, zpath, wpath = _zkey_case(tmp_path)
    out = [str(tmp_path / "proof.json"), str(tmp_path / "public.json")]
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True, timeout=120)
    r = run(zpath)
    assert r.returncode == 1 and "usage" in r.stderr
    junk = str(tmp_path / "junk.zkey")
    open(junk, "wb").write(open(zpath, "rb").read()[:300])
    r = run(junk, wpath, *out)
    assert r.returncode == 1 and "not a BN254 Groth16 .zkey" in r.stderr
    short = str(tmp_path / "short.wtns")
    open(short, "wb").write(cio.write_wtns(z[:-1]))
    r = run(zpath, short, *out)
    assert r.returncode == 1 and "the key has %d variables" % cs["m"] in r.stderr
    assert not os.path.exists(out[0])
    if not torch.cuda.is_available():
        r = run(zpath, wpath, *out)
        assert r.returncode == 2 and "no CPU fallback" in r.stderr and not os.path.exists(out[0])


@pytest.mark.gpu
def test_zkey_prover_example_writes_a_proof_the_oracle_verifies(tmp_path):
    import json
    import subprocess
    import oracle_a as o
    exe = _build_example(tmp_path, "zkey_prover")
    cs, z, vk, zpath, wpath = _zkey_case(tmp_path)
    out = [str(tmp_path / "proof.json"), str(tmp_path / "public.json")]
    r = subprocess.run([exe, zpath, wpath, *out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "pairing check passed" in r.stderr, r.stderr
    proof = o.parse_proof(open(out[0]).read())
    public = [int(x) for x in json.load(open(out[1]))]
    assert public == [z[1]]
    assert o.groth16_verify(vk, proof, public)
    # fresh blinding every run: another proof of the same statement, still valid
    r = subprocess.run([exe, zpath, wpath, *out], capture_output=True, text=True, timeout=300)
    proof2 = o.parse_proof(open(out[0]).read())
    assert r.returncode == 0 and proof2 != proof and o.groth16_verify(vk, proof2, public)


def test_cpp_example_builds_against_the_c_abi_and_fails_loudly_without_gpu(tmp_path):
    import subprocess
    import torch
    exe = _build_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "8"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_example_proves_and_prints_reference_layout(tmp_path):
    import subprocess
    from eigen_zeth_b200 import proof_format as pf
    exe = _build_example(tmp_path)
    r = subprocess.run([exe, "10"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "pairing check passed" in r.stderr
    proof_txt, input_txt = r.stdout.split("}\n[", 1)
    proof_txt += "}"
    input_txt = "[" + input_txt
    a, b, c = pf.parse_proof(proof_txt)
    assert o.g1_is_on_curve(a) and o.g2_is_on_curve(b) and o.g1_is_on_curve(c)
    assert pf.parse_public_input(input_txt)[0] < o.R
    # byte-identical to the library's own writer
    assert pf.proof_to_json(pf.proof_from_json(proof_txt)) == proof_txt.rstrip("\n")


def test_get_batch_proof_view_matches_reference_rpc():
    """eigenrpc_getBatchProof (reference src/custom_reth/eigen.rs:76-97): BATCH_PROOF_{n} -> BatchProofInfo
    with the proof strings verbatim and 0x-hex state roots; a missing key is None."""
    from eigen_zeth_b200 import proof_format as pf
    pre, post = bytes(range(32)), bytes(range(255, 223, -1))
    pr = pf.ProofResult(42, GOLDEN["proof_json"], GOLDEN["public_input_json"], pre, post)
    db = {b"BATCH_PROOF_42": pr.to_json().encode(), b"BATCH_PROOF_43": b"{not a proof result"}
    assert pf.batch_proof_key(42) == b"BATCH_PROOF_42"
    info = pf.get_batch_proof(db.get, 42)
    assert info.block_number == 42 and info.proof == GOLDEN["proof_json"] and info.public_input == GOLDEN["public_input_json"]
    assert info.pre_state_root == "0x000102030405060708090a0b0c0d0e0f101112131415161718191a1b1c1d1e1f"
    assert info.post_state_root == "0x" + post.hex() and len(info.post_state_root) == 66
    assert json.loads(info.to_json())["pre_state_root"] == info.pre_state_root
    assert list(json.loads(info.to_json())) == ["block_number", "proof", "public_input", "pre_state_root", "post_state_root"]
    assert pf.get_batch_proof(db.get, 7) is None
    with pytest.raises(ValueError):
        pf.get_batch_proof(db.get, 43)
    # the canned debug proof of the reference (worker.rs:49-96: post_state_root[24..] = batch le bytes) views cleanly
    post_dbg = bytes(24) + (9).to_bytes(8, "little")
    db[b"BATCH_PROOF_9"] = pf.ProofResult(9, "p", "i", bytes(32), post_dbg).to_json().encode()
    assert pf.get_batch_proof(db.get, 9).post_state_root == "0x" + "00" * 24 + "0900000000000000"


def test_python_binding_checks_buffer_lengths(zkb):
    """the C ABI takes bare pointers; the binding refuses short buffers before they reach cudaMemcpy"""
    class FakeCtx:
        lib = zkb.load_library()
    pk = zkb.lib.ProvingKey.__new__(zkb.lib.ProvingKey)
    pk.n_vars = 16
    with pytest.raises(ValueError, match="witness holds"):
        pk.witness_ptr(bytes(32 * 15))
    ptr, keep = pk.witness_ptr(bytes(32 * 16))
    assert ptr.value and keep is not None


def test_fast_final_exponentiation_equals_square_and_multiply(zkb):
    """the verifier's hard part uses the base-q decomposition with three powers by the BN parameter and
    Frobenius maps; it must give exactly what plain square-and-multiply on (q^12 - 1) / r gives"""
    log_n, seed = 5, 31
    opk = ob.ProvingKey(log_n, seed)
    wit = ob.synthetic_witness(log_n, fr_bytes([3]), fr_bytes([4]))
    proof, _, _, _ = opk.prove(wit, fr_bytes([5]), fr_bytes([6]))
    assert zkb.load_library().zkb_selftest_final_exp(proof) == 0
    assert zkb.groth16_verify(opk.vk(), proof, wit[32:64])
    assert not zkb.groth16_verify(opk.vk(), proof, wit[64:96])
    # same key again (memoised key validation), another key (memo replaced), the first key again
    opk2 = ob.ProvingKey(log_n, seed + 1)
    proof2, _, _, _ = opk2.prove(wit, fr_bytes([5]), fr_bytes([6]))
    assert zkb.groth16_verify(opk2.vk(), proof2, wit[32:64])
    assert not zkb.groth16_verify(opk2.vk(), proof, wit[32:64])
    assert zkb.groth16_verify(opk.vk(), proof, wit[32:64])
    bad_vk = bytearray(opk.vk())
    bad_vk[64] ^= 1                                        # beta_g2 leaves the twist
    with pytest.raises(zkb.ZkbError):
        zkb.groth16_verify(bytes(bad_vk), proof, wit[32:64])
    assert zkb.groth16_verify(opk.vk(), proof, wit[32:64])


def test_host_proof_assembly_matches_oracle(zkb):
    """K15 on the host: the six scalar multiplications and additions that turn the five MSM sums into
    (A, B, C) — against oracle B's proof bytes for full-size and edge blinding scalars."""
    import time
    log_n, seed = 6, 404
    opk = ob.ProvingKey(log_n, seed)
    wit = ob.synthetic_witness(log_n, fr_bytes([12]), fr_bytes([34]))
    rng = o.Xoshiro256ss(9)
    cases = [(rng.fr(), rng.fr()), (0, 0), (1, o.R - 1), (o.R - 1, 1), (5, 0)]
    for r, s in cases:
        rb, sb = fr_bytes([r]), fr_bytes([s])
        proof, inter, _, _ = opk.prove(wit, rb, sb)
        t0 = time.perf_counter()
        got = zkb.lib.groth16_assemble_host(opk.fixed(), inter, rb, sb)
        dt = time.perf_counter() - t0
        assert got == proof, (r, s)
        assert zkb.groth16_verify(opk.vk(), got, wit[32:64])
        f = opk.fixed()
        blob = f["alpha_g1"] + f["beta_g1"] + f["delta_g1"] + f["beta_g2"] + f["delta_g2"]
        assert zkb.load_library().zkb_selftest_assemble_tables(blob, inter, rb, sb) == 0, (r, s)
    assert dt < 0.05
    with pytest.raises(zkb.ZkbError):
        zkb.lib.groth16_assemble_host(opk.fixed(), inter, fr_bytes([1]), int(o.R).to_bytes(32, "little"))


def test_groth16_verify_ex_any_number_of_public_inputs(zkb):
    """f2: the host verifier for keys with 0, 1 or several public inputs, against Oracle A proofs"""
    import oracle_a as o
    from helpers import small_circuit
    g1s = lambda pts: b"".join(o.g1_to_bytes(P) for P in pts)
    for n_pub in (0, 1, 3):
        cs, z = small_circuit(5)
        cs["n_pub"] = n_pub                        # wires 1..3 = out, x, y: any prefix can be public
        rng = o.Xoshiro256ss(0xF2 + n_pub)
        tau, alpha, beta, gamma, delta, r, s = [rng.fr() or 1 for _ in range(7)]
        pk, vk = o.groth16_setup(cs, tau, alpha, beta, gamma, delta)
        (A, B, C), _ = o.groth16_prove(pk, cs, z, r, s)
        pub = z[1:1 + n_pub]
        assert o.groth16_verify(vk, (A, B, C), pub)
        fixed = o.g1_to_bytes(vk["alpha_g1"]) + o.g2_to_bytes(vk["beta_g2"]) + o.g2_to_bytes(vk["gamma_g2"]) + o.g2_to_bytes(vk["delta_g2"])
        ic = g1s(vk["gamma_abc_g1"])
        proof = o.g1_to_bytes(A) + o.g2_to_bytes(B) + o.g1_to_bytes(C)
        inputs = b"".join(v.to_bytes(32, "little") for v in pub)
        assert zkb.groth16_verify_ex(fixed, ic, proof, inputs)
        if n_pub == 1:
            assert zkb.groth16_verify(fixed + ic, proof, inputs)
        for i in range(n_pub):                     # any single wrong input fails
            bad = bytearray(inputs)
            bad[32 * i] ^= 1
            assert not zkb.groth16_verify_ex(fixed, ic, proof, bytes(bad))
        other = o.g1_to_bytes(o.g1_add(A, o.G1_GEN)) + proof[64:]
        assert not zkb.groth16_verify_ex(fixed, ic, other, inputs)
        if n_pub:
            with pytest.raises(zkb.ZkbError):      # an input >= r is malformed, not merely invalid
                zkb.groth16_verify_ex(fixed, ic, proof, int(o.R).to_bytes(32, "little") + inputs[32:])
            with pytest.raises(ValueError):
                zkb.groth16_verify_ex(fixed, ic, proof, inputs[32:])
        off = bytearray(ic)
        off[0] ^= 1
        with pytest.raises(zkb.ZkbError):          # ic0 off the curve
            zkb.groth16_verify_ex(fixed, bytes(off), proof, inputs)
