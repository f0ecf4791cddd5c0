"""Regenerates the extern "C" block of rust/zkb-sys/src/lib.rs from include/zkb.h.
The crate cannot be built in this image (no Rust toolchain); tests/test_abi_host.py checks that the
generated file is in sync with the header.  Usage: python tools/gen_rust_sys.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TMAP = {"int": "c_int", "int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "double": "f64",
        "float": "f32", "void": "c_void", "char": "c_char", "uint8_t": "u8", "size_t": "usize"}


def rust_type(t):
    t = re.sub(r"\s+", " ", t.strip())
    const = "const" in t.split("*")[0].split() if "*" in t else False
    base = t.replace("const", "").strip()
    stars = base.count("*")
    base = base.replace("*", "").strip()
    if base.startswith("struct "):
        base = base[7:]
    r = TMAP.get(base, base)
    for i in range(stars):
        r = ("*const " if (const and i == stars - 1) else "*mut ") + r
    return r


def extern_block():
    hdr = open(os.path.join(ROOT, "include", "zkb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    src = re.sub(r"#.*", "", src)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(zkb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    out = []
    for ret, name, args in protos:
        params = []
        args = re.sub(r"\s+", " ", args.strip())
        if args and args != "void":
            for a in args.split(","):
                m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)(\[[^\]]*\])?$", a.strip())
                ty, nm, arr = m.group(1).strip(), m.group(2), m.group(3)
                if arr:
                    ty += " *"
                ty = ty.replace("* const", "*")
                if nm in ("type", "in", "out", "ref", "box", "move", "as"):
                    nm += "_"
                params.append("%s: %s" % (nm, rust_type(ty)))
        rr = rust_type(ret.strip())
        sig = "    pub fn %s(%s)" % (name, ", ".join(params))
        if rr != "c_void":
            sig += " -> %s" % rr
        out.append(sig + ";")
    return "\n".join(out), [n for _, n, _ in protos]


def main():
    path = os.path.join(ROOT, "rust", "zkb-sys", "src", "lib.rs")
    text = open(path).read()
    block, _ = extern_block()
    head, rest = text.split('extern "C" {\n', 1)
    new = head + 'extern "C" {\n' + block + "\n}\n"
    if "--check" in sys.argv:
        sys.exit(0 if new == text else 1)
    open(path, "w").write(new)
    print("wrote", path)


if __name__ == "__main__":
    main()
