"""shared helpers for the parity tests (oracle side only)."""
import oracle_a as o


def fr_bytes(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def fr_list(buf):
    return [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]


def g1_bytes(pts):
    return b"".join(o.g1_to_bytes(p) for p in pts)


def g1_list(buf):
    return [o.g1_from_bytes(buf[i:i + 64]) for i in range(0, len(buf), 64)]


def g2_bytes(pts):
    return b"".join(o.g2_to_bytes(p) for p in pts)


def g2_list(buf):
    return [o.g2_from_bytes(buf[i:i + 128]) for i in range(0, len(buf), 128)]


def rand_fr(rng, n):
    return [rng.fr() for _ in range(n)]


def rand_fq(rng, n):
    out = []
    while len(out) < n:
        v = rng.next() | (rng.next() << 64) | (rng.next() << 128) | ((rng.next() & ((1 << 62) - 1)) << 192)
        if v < o.Q:
            out.append(v)
    return out
