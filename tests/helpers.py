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


def small_circuit(n_constraints: int):
    """a circom-shaped circuit that is NOT the synthetic chain: wires (1, out, x, y, t_0..t_{k-1}),
    one public output, two private inputs;  t_0 = x*y,  t_{j+1} = (t_j + 3x + j) * (t_j - y),
    last row binds out = t_last * 1.  Returns (cs dict in oracle A's form, satisfying assignment)."""
    import oracle_a as o
    k = n_constraints - 1
    assert k >= 1
    m = 4 + k
    A, B, C = [[(2, 1)]], [[(3, 1)]], [[(4, 1)]]
    for j in range(k - 1):
        A.append([(4 + j, 1), (2, 3), (0, j)] if j else [(4 + j, 1), (2, 3)])
        B.append([(4 + j, 1), (3, o.R - 1)])
        C.append([(5 + j, 1)])
    A.append([(4 + k - 1, 1)])
    B.append([(0, 1)])
    C.append([(1, 1)])
    x, y = 0x1234567, o.R - 5
    t = [x * y % o.R]
    for j in range(k - 1):
        t.append((t[-1] + 3 * x + j) * (t[-1] - y) % o.R)
    z = [1, t[-1], x, y] + t
    return {"A": A, "B": B, "C": C, "m": m, "n_pub": 1}, z
