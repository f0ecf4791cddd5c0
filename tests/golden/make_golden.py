"""Regenerates tests/golden/reference_fixtures.json from the read-only reference checkout.

The reference pins only the proof.json *layout* (SURVEY.md §8c): this script extracts the three
proof fixtures and the public-input fixture the reference's own tests hold
  - proof/proof.json + proof/public_input.json                (read at src/settlement/worker.rs:56-59)
  - test_parse_proof / test_parse_public_input                (src/settlement/ethereum/mod.rs:487-589)
  - the compact proof string of the ignored worker test       (src/settlement/worker.rs:760-761)
so the tests can run where /root/reference does not exist (the GPU box).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")


def main():
    fx = {}
    fx["proof_json"] = open(os.path.join(REF, "proof/proof.json")).read()
    fx["public_input_json"] = open(os.path.join(REF, "proof/public_input.json")).read()
    src = open(os.path.join(REF, "src/settlement/ethereum/mod.rs")).read()
    m = re.search(r'fn test_parse_proof\(\) \{\s*let json_str = r#"(.*?)"#;', src, re.S)
    fx["test_parse_proof_json"] = m.group(1)
    m = re.search(r'fn test_parse_public_input\(\) \{\s*let json_str = r#"(.*?)"#;', src, re.S)
    fx["test_parse_public_input_json"] = m.group(1)
    # the values the reference test asserts, in its order: a.x a.y b.x[0] b.x[1] b.y[0] b.y[1] c.x c.y
    body = src[src.index("fn test_parse_proof()"):src.index("fn test_parse_public_input()")]
    body = body[body.index("let proof = parse_proof"):]
    fx["test_parse_proof_expected"] = re.findall(r'"(\d{60,})"', body)
    body2 = src[src.index("fn test_parse_public_input()"):]
    body2 = body2[body2.index("let input = parse_public_input"):]
    fx["test_parse_public_input_expected"] = re.findall(r'"(\d{60,})"', body2)[:1]
    w = open(os.path.join(REF, "src/settlement/worker.rs")).read()
    m = re.search(r'let proof = r#"(\{.*?\})"#', w, re.S) or re.search(r'"(\{\\"pi_a\\".*?\})"', w, re.S)
    if m:
        fx["worker_compact_proof_json"] = m.group(1).replace('\\"', '"')
    json.dump(fx, open(OUT, "w"), indent=1)
    print("wrote", OUT, {k: len(v) for k, v in fx.items()})


if __name__ == "__main__":
    main()
