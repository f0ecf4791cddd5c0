"""Oracle A — pure-Python big-integer restatement of the BN254 Groth16 proving path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this file; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

What it restates: the work eigen-zeth hands to its external prover at
`GenFinalProofRequest` (reference proto/prover/v1/prover.proto:130-148, sent from
src/prover/provider.rs:472-483) and whose result is consumed by `parse_proof` /
`parse_public_input` (src/settlement/ethereum/mod.rs:445-481).  The arithmetic itself is
NOT in the reference tree (SURVEY.md finding 1); it follows the published algorithms of
arkworks 0.4 (ark-groth16 `create_proof_with_reduction_and_matrices` with the
LibsnarkReduction, ark-ec `VariableBaseMSM`, ark-poly `Radix2EvaluationDomain`), restated
from SURVEY.md Appendix B.  PARITY UNPINNED for proof values: the reference holds no
(vk, proof, input) triple; this oracle is pinned by (i) the BN254
constants of Appendix A, (ii) group-membership of the reference's three proof fixtures,
(iii) pairing bilinearity, (iv) Groth16 completeness/soundness on its own proofs.

Everything is plain Python ints: slow, obviously correct, for n <= 2^12.
"""
from __future__ import annotations

# --------------------------------------------------------------------------- constants
X_BN = 4965661367192848881
Q = 36 * X_BN**4 + 36 * X_BN**3 + 24 * X_BN**2 + 6 * X_BN + 1   # base field modulus
R = 36 * X_BN**4 + 36 * X_BN**3 + 18 * X_BN**2 + 6 * X_BN + 1   # scalar field modulus
assert Q == 21888242871839275222246405745257275088696311157297823662689037894645226208583
assert R == 21888242871839275222246405745257275088548364400416034343698204186575808495617
ATE_LOOP = 6 * X_BN + 2
FR_GEN = 5                     # multiplicative generator of Fr, also the coset shift
FR_TWO_ADICITY = 28
FR_ROOT_2_28 = pow(FR_GEN, (R - 1) >> FR_TWO_ADICITY, R)
assert FR_ROOT_2_28 == 19103219067921713944291392827692070036145651957329286315305642004821462161904

G1_GEN = (1, 2)
G2_GEN = (
    (10857046999023057135944570762232829481370756359578518086990519993285655852781,
     11559732032986387107991004021392285783925812861821192530917403151452391805634),
    (8495653923123431417604973247489272438418190587263600148770280649306958101930,
     4082367875863433681332203403145435568316851327593401208105741076214120093531),
)


def inv(a: int, p: int) -> int:
    return pow(a, p - 2, p)


# --------------------------------------------------------------------------- PRNG
MASK64 = (1 << 64) - 1


class Xoshiro256ss:
    """xoshiro256** seeded through splitmix64 — shared by Python, C++ and the product host code."""

    def __init__(self, seed: int):
        s = seed & MASK64
        st = []
        for _ in range(4):
            s = (s + 0x9E3779B97F4A7C15) & MASK64
            z = s
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
            st.append(z ^ (z >> 31))
        self.s = st

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (64 - k))) & MASK64

    def next(self) -> int:
        s = self.s
        result = (self._rotl((s[1] * 5) & MASK64, 7) * 9) & MASK64
        t = (s[1] << 17) & MASK64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 45)
        return result

    def fr(self) -> int:
        """uniform in [0, R): 4 words little-endian, masked to 254 bits, rejection."""
        while True:
            v = self.next() | (self.next() << 64) | (self.next() << 128) | (self.next() << 192)
            v &= (1 << 254) - 1
            if v < R:
                return v


# --------------------------------------------------------------------------- Fq2 / Fq6 / Fq12
# Fq2 = Fq[u]/(u^2+1);  Fq6 = Fq2[v]/(v^3 - xi), xi = 9+u;  Fq12 = Fq6[w]/(w^2 - v)
XI = (9, 1)


def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
def f2_conj(a): return (a[0], (-a[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_sqr(a): return f2_mul(a, a)
def f2_scale(a, k): return (a[0] * k % Q, a[1] * k % Q)


def f2_inv(a):
    d = inv((a[0] * a[0] + a[1] * a[1]) % Q, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)


def f2_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


F2_ZERO, F2_ONE = (0, 0), (1, 0)
TWIST_B = f2_mul((3, 0), f2_inv(XI))      # y^2 = x^3 + 3/xi on the twist


def f6_add(a, b): return tuple(f2_add(x, y) for x, y in zip(a, b))
def f6_sub(a, b): return tuple(f2_sub(x, y) for x, y in zip(a, b))
def f6_neg(a): return tuple(f2_neg(x) for x in a)


def f6_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    c0 = f2_add(f2_mul(a0, b0), f2_mul(XI, f2_add(f2_mul(a1, b2), f2_mul(a2, b1))))
    c1 = f2_add(f2_add(f2_mul(a0, b1), f2_mul(a1, b0)), f2_mul(XI, f2_mul(a2, b2)))
    c2 = f2_add(f2_add(f2_mul(a0, b2), f2_mul(a1, b1)), f2_mul(a2, b0))
    return (c0, c1, c2)


def f6_mul_v(a):  # multiply by v
    return (f2_mul(XI, a[2]), a[0], a[1])


def f6_inv(a):
    a0, a1, a2 = a
    t0 = f2_sub(f2_sqr(a0), f2_mul(XI, f2_mul(a1, a2)))
    t1 = f2_sub(f2_mul(XI, f2_sqr(a2)), f2_mul(a0, a1))
    t2 = f2_sub(f2_sqr(a1), f2_mul(a0, a2))
    d = f2_add(f2_mul(a0, t0), f2_mul(XI, f2_add(f2_mul(a2, t1), f2_mul(a1, t2))))
    di = f2_inv(d)
    return (f2_mul(t0, di), f2_mul(t1, di), f2_mul(t2, di))


F6_ZERO = (F2_ZERO, F2_ZERO, F2_ZERO)
F6_ONE = (F2_ONE, F2_ZERO, F2_ZERO)
F12_ONE = (F6_ONE, F6_ZERO)


def f12_mul(a, b):
    a0, a1 = a
    b0, b1 = b
    t0 = f6_mul(a0, b0)
    t1 = f6_mul(a1, b1)
    c0 = f6_add(t0, f6_mul_v(t1))
    c1 = f6_sub(f6_sub(f6_mul(f6_add(a0, a1), f6_add(b0, b1)), t0), t1)
    return (c0, c1)


def f12_inv(a):
    a0, a1 = a
    d = f6_inv(f6_sub(f6_mul(a0, a0), f6_mul_v(f6_mul(a1, a1))))
    return (f6_mul(a0, d), f6_neg(f6_mul(a1, d)))


def f12_pow(a, e):
    r = F12_ONE
    while e:
        if e & 1:
            r = f12_mul(r, a)
        a = f12_mul(a, a)
        e >>= 1
    return r


# --------------------------------------------------------------------------- curves (affine, None = infinity)
def g1_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - 3) % Q == 0


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q)


def g1_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = 3 * x1 * x1 * inv(2 * y1, Q) % Q
    else:
        lam = (y2 - y1) * inv((x2 - x1) % Q, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)


def g1_mul(P, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


def g2_is_on_curve(P):
    if P is None:
        return True
    x, y = P
    return f2_sub(f2_sqr(y), f2_add(f2_mul(f2_sqr(x), x), TWIST_B)) == F2_ZERO


def g2_neg(P):
    return None if P is None else (P[0], f2_neg(P[1]))


def g2_add(P, S):
    if P is None:
        return S
    if S is None:
        return P
    x1, y1 = P
    x2, y2 = S
    if x1 == x2:
        if f2_add(y1, y2) == F2_ZERO:
            return None
        lam = f2_mul(f2_scale(f2_sqr(x1), 3), f2_inv(f2_scale(y1, 2)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(P, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, P)
        P = g2_add(P, P)
        k >>= 1
    return acc


def g2_in_subgroup(P):
    if P is None:
        return True
    acc, k, T = None, R, P            # must not reduce k mod R here
    while k:
        if k & 1:
            acc = g2_add(acc, T)
        T = g2_add(T, T)
        k >>= 1
    return acc is None


# --------------------------------------------------------------------------- pairing (optimal ate)
FROB_GAMMA_X = f2_pow(XI, (Q - 1) // 3)
FROB_GAMMA_Y = f2_pow(XI, (Q - 1) // 2)


def g2_frobenius(P):
    return (f2_mul(f2_conj(P[0]), FROB_GAMMA_X), f2_mul(f2_conj(P[1]), FROB_GAMMA_Y))


def _line(T, S, P):
    """line through T,S (twist coords; tangent if T==S) evaluated at P in G1 -> (Fq12 value, T+S)."""
    xt, yt = T
    xs, ys = S
    if T == S:
        lam = f2_mul(f2_scale(f2_sqr(xt), 3), f2_inv(f2_scale(yt, 2)))
    else:
        lam = f2_mul(f2_sub(ys, yt), f2_inv(f2_sub(xs, xt)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), xt), xs)
    y3 = f2_sub(f2_mul(lam, f2_sub(xt, x3)), yt)
    xp, yp = P
    # l = yP - lam*xP*w + (lam*xT - yT)*w^3,   w^3 = v*w
    c0 = ((yp % Q, 0), F2_ZERO, F2_ZERO)
    c1 = (f2_neg(f2_scale(lam, xp)), f2_sub(f2_mul(lam, xt), yt), F2_ZERO)
    return (c0, c1), (x3, y3)


def miller_loop(P, Qt):
    """P in G1 affine, Qt in G2 affine (twist coords). Infinity -> 1."""
    if P is None or Qt is None:
        return F12_ONE
    f = F12_ONE
    T = Qt
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        l, T2 = _line(T, T, P)
        f = f12_mul(f12_mul(f, f), l)
        T = T2
        if (ATE_LOOP >> i) & 1:
            l, T2 = _line(T, Qt, P)
            f = f12_mul(f, l)
            T = T2
    Q1 = g2_frobenius(Qt)
    Q2 = g2_neg(g2_frobenius(Q1))
    l, T = _line(T, Q1, P)
    f = f12_mul(f, l)
    l, T = _line(T, Q2, P)
    f = f12_mul(f, l)
    return f


FINAL_EXP = (Q**12 - 1) // R


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def pairing(P, Qt):
    return final_exponentiation(miller_loop(P, Qt))


def pairing_product_is_one(pairs):
    f = F12_ONE
    for P, Qt in pairs:
        f = f12_mul(f, miller_loop(P, Qt))
    return final_exponentiation(f) == F12_ONE


# --------------------------------------------------------------------------- NTT over Fr
def fr_root_of_unity(log_n: int) -> int:
    assert 0 <= log_n <= FR_TWO_ADICITY
    return pow(FR_ROOT_2_28, 1 << (FR_TWO_ADICITY - log_n), R)


def ntt(a, inverse=False):
    """natural-order in, natural-order out radix-2 NTT; inverse includes 1/n."""
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    a = list(a)
    for i in range(n):                       # bit reversal
        j = int(format(i, "0%db" % log_n)[::-1], 2) if log_n else 0
        if i < j:
            a[i], a[j] = a[j], a[i]
    w_n = fr_root_of_unity(log_n)
    if inverse:
        w_n = inv(w_n, R)
    m = 1
    while m < n:
        w_m = pow(w_n, n // (2 * m), R)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % R
                u = a[k + j]
                a[k + j] = (u + t) % R
                a[k + j + m] = (u - t) % R
                w = w * w_m % R
        m *= 2
    if inverse:
        ninv = inv(n, R)
        a = [x * ninv % R for x in a]
    return a


def dft_naive(a, inverse=False):
    n = len(a)
    w = fr_root_of_unity(n.bit_length() - 1)
    if inverse:
        w = inv(w, R)
    out = [sum(a[j] * pow(w, i * j, R) for j in range(n)) % R for i in range(n)]
    if inverse:
        ninv = inv(n, R)
        out = [x * ninv % R for x in out]
    return out


def coset_ntt(a, shift=FR_GEN):
    """evaluate polynomial with coefficients a on shift*<w>  (ark-poly coset fft)."""
    g = 1
    b = []
    for x in a:
        b.append(x * g % R)
        g = g * shift % R
    return ntt(b)


def coset_intt(a, shift=FR_GEN):
    b = ntt(a, inverse=True)
    gi = inv(shift, R)
    g = 1
    out = []
    for x in b:
        out.append(x * g % R)
        g = g * gi % R
    return out


# --------------------------------------------------------------------------- MSM
def msm_g1_naive(bases, scalars):
    acc = None
    for P, k in zip(bases, scalars):
        acc = g1_add(acc, g1_mul(P, k))
    return acc


def msm_g2_naive(bases, scalars):
    acc = None
    for P, k in zip(bases, scalars):
        acc = g2_add(acc, g2_mul(P, k))
    return acc


def _msm_pippenger(bases, scalars, add, c=None):
    """unsigned-window Pippenger as in ark-ec VariableBaseMSM (c = 3 if N<32 else ceil(ln N)+2)."""
    import math
    n = len(bases)
    if c is None:
        c = 3 if n < 32 else int(math.ceil(math.log(n))) + 2
    nwin = (254 + c - 1) // c
    total = None
    for w in range(nwin - 1, -1, -1):
        if total is not None:
            for _ in range(c):
                total = add(total, total)
        buckets = [None] * ((1 << c) - 1)
        for P, k in zip(bases, scalars):
            d = (k >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = add(buckets[d - 1], P)
        run = None
        acc = None
        for b in reversed(buckets):
            run = add(run, b)
            acc = add(acc, run)
        total = add(total, acc)
    return total


def msm_g1(bases, scalars, c=None):
    return _msm_pippenger(bases, scalars, g1_add, c)


def msm_g2(bases, scalars, c=None):
    return _msm_pippenger(bases, scalars, g2_add, c)


# --------------------------------------------------------------------------- synthetic R1CS (DESIGN.md "synthetic chain circuit")
def synthetic_r1cs(log_n: int):
    """m = 2^log_n variables z = (1, x_pub, w...), 2^log_n - 2 constraints, 1 public input.

    rows are lists of (col, coeff).  Constraint j (0 <= j <= m-5):
        (z[j+2] + (j%7+1)*z[0]) * z[j+3] = z[j+4]
    constraint m-4:   z[m-1] * z[m-2] = z[1]
    constraint m-3:   (z[1] + z[2]) * z[0] = z[1] + z[2]
    """
    m = 1 << log_n
    assert m >= 8
    A, B, C = [], [], []
    for j in range(m - 4):
        A.append([(0, j % 7 + 1), (j + 2, 1)])
        B.append([(j + 3, 1)])
        C.append([(j + 4, 1)])
    A.append([(m - 1, 1)]); B.append([(m - 2, 1)]); C.append([(1, 1)])
    A.append([(1, 1), (2, 1)]); B.append([(0, 1)]); C.append([(1, 1), (2, 1)])
    return {"m": m, "n_pub": 1, "A": A, "B": B, "C": C}


def synthetic_witness(log_n: int, w0: int, w1: int):
    m = 1 << log_n
    z = [0] * m
    z[0] = 1
    z[2], z[3] = w0 % R, w1 % R
    for j in range(m - 4):
        z[j + 4] = (z[j + 2] + (j % 7 + 1)) * z[j + 3] % R
    z[1] = z[m - 1] * z[m - 2] % R
    return z


def r1cs_is_satisfied(cs, z):
    dot = lambda row: sum(c * z[i] for i, c in row) % R
    return all(dot(a) * dot(b) % R == dot(c) for a, b, c in zip(cs["A"], cs["B"], cs["C"]))


# --------------------------------------------------------------------------- Groth16 (arkworks / libsnark reduction)
def domain_size(cs):
    n = len(cs["A"]) + cs["n_pub"] + 1
    return 1 << (n - 1).bit_length()


def lagrange_at(tau: int, n: int):
    """L_j(tau) for the size-n radix-2 domain."""
    w = fr_root_of_unity(n.bit_length() - 1)
    zt = (pow(tau, n, R) - 1) % R
    out = []
    wj = 1
    ninv = inv(n, R)
    for _ in range(n):
        out.append(zt * ninv % R * wj % R * inv((tau - wj) % R, R) % R)
        wj = wj * w % R
    return out


def groth16_setup(cs, tau, alpha, beta, gamma, delta):
    """returns (pk, vk).  Query layout follows ark-groth16 ProvingKey:
    a_query[m], b_g1_query[m], b_g2_query[m], l_query[m-1-n_pub], h_query[n-1]."""
    m, n_pub = cs["m"], cs["n_pub"]
    n_inst = n_pub + 1
    n_cons = len(cs["A"])
    n = domain_size(cs)
    L = lagrange_at(tau, n)
    u = [0] * m
    v = [0] * m
    w = [0] * m
    for j in range(n_cons):
        for i, c in cs["A"][j]:
            u[i] = (u[i] + c * L[j]) % R
        for i, c in cs["B"][j]:
            v[i] = (v[i] + c * L[j]) % R
        for i, c in cs["C"][j]:
            w[i] = (w[i] + c * L[j]) % R
    for i in range(n_inst):                   # input-consistency rows appended to A
        u[i] = (u[i] + L[n_cons + i]) % R
    zt = (pow(tau, n, R) - 1) % R
    gi, di = inv(gamma, R), inv(delta, R)
    G1, G2 = G1_GEN, G2_GEN
    pk = {
        "n": n, "m": m, "n_pub": n_pub,
        "alpha_g1": g1_mul(G1, alpha), "beta_g1": g1_mul(G1, beta), "delta_g1": g1_mul(G1, delta),
        "beta_g2": g2_mul(G2, beta), "delta_g2": g2_mul(G2, delta),
        "a_query": [g1_mul(G1, x) for x in u],
        "b_g1_query": [g1_mul(G1, x) for x in v],
        "b_g2_query": [g2_mul(G2, x) for x in v],
        "l_query": [g1_mul(G1, (beta * u[i] + alpha * v[i] + w[i]) % R * di % R) for i in range(n_inst, m)],
        "h_query": [g1_mul(G1, pow(tau, i, R) * zt % R * di % R) for i in range(n - 1)],
    }
    vk = {
        "alpha_g1": pk["alpha_g1"], "beta_g2": pk["beta_g2"],
        "gamma_g2": g2_mul(G2, gamma), "delta_g2": pk["delta_g2"],
        "gamma_abc_g1": [g1_mul(G1, (beta * u[i] + alpha * v[i] + w[i]) % R * gi % R) for i in range(n_inst)],
    }
    return pk, vk


def qap_witness_map(cs, z):
    """h = coefficients of (A.B - C)/Z, length n-1 (LibsnarkReduction::witness_map_from_matrices)."""
    n = domain_size(cs)
    n_cons = len(cs["A"])
    dot = lambda row: sum(c * z[i] for i, c in row) % R
    a = [dot(r) for r in cs["A"]] + [0] * (n - n_cons)
    b = [dot(r) for r in cs["B"]] + [0] * (n - n_cons)
    c = [dot(r) for r in cs["C"]] + [0] * (n - n_cons)
    for i in range(cs["n_pub"] + 1):
        a[n_cons + i] = z[i]
    evals = (list(a), list(b), list(c))
    a = coset_ntt(ntt(a, inverse=True))
    b = coset_ntt(ntt(b, inverse=True))
    c = coset_ntt(ntt(c, inverse=True))
    zinv = inv((pow(FR_GEN, n, R) - 1) % R, R)
    ab = [(x * y - t) % R * zinv % R for x, y, t in zip(a, b, c)]
    h = coset_intt(ab)
    assert h[n - 1] == 0
    return h[: n - 1], evals


def groth16_prove(pk, cs, z, r, s, msm1=msm_g1, msm2=msm_g2):
    """returns (A, B, C) affine + dict of intermediates (the five MSM results and h)."""
    n_inst = pk["n_pub"] + 1
    h, evals = qap_witness_map(cs, z)
    msm_a = msm1(pk["a_query"], z)
    msm_b1 = msm1(pk["b_g1_query"], z)
    msm_b2 = msm2(pk["b_g2_query"], z)
    msm_l = msm1(pk["l_query"], z[n_inst:])
    msm_h = msm1(pk["h_query"], h)
    A = g1_add(g1_add(pk["alpha_g1"], msm_a), g1_mul(pk["delta_g1"], r))
    B1 = g1_add(g1_add(pk["beta_g1"], msm_b1), g1_mul(pk["delta_g1"], s))
    B2 = g2_add(g2_add(pk["beta_g2"], msm_b2), g2_mul(pk["delta_g2"], s))
    C = g1_add(g1_add(msm_l, msm_h), g1_add(g1_mul(A, s), g1_mul(B1, r)))
    C = g1_add(C, g1_neg(g1_mul(pk["delta_g1"], r * s % R)))
    inter = {"h": h, "a_evals": evals[0], "b_evals": evals[1], "c_evals": evals[2],
             "msm_a": msm_a, "msm_b1": msm_b1, "msm_b2": msm_b2, "msm_l": msm_l, "msm_h": msm_h}
    return (A, B2, C), inter


def groth16_verify(vk, proof, public_inputs):
    A, B, C = proof
    if not (g1_is_on_curve(A) and g1_is_on_curve(C) and g2_is_on_curve(B)):
        return False
    acc = vk["gamma_abc_g1"][0]
    for x, P in zip(public_inputs, vk["gamma_abc_g1"][1:]):
        acc = g1_add(acc, g1_mul(P, x))
    return pairing_product_is_one([
        (A, B), (g1_neg(vk["alpha_g1"]), vk["beta_g2"]),
        (g1_neg(acc), vk["gamma_g2"]), (g1_neg(C), vk["delta_g2"]),
    ])


# --------------------------------------------------------------------------- byte encodings (canonical LE)
def fr_to_bytes(x: int) -> bytes:
    return int(x).to_bytes(32, "little")


def g1_to_bytes(P) -> bytes:
    """affine x||y, 32-byte little-endian canonical; infinity = 64 zero bytes."""
    if P is None:
        return bytes(64)
    return P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little")


def g1_from_bytes(b: bytes):
    if b == bytes(64):
        return None
    return (int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little"))


def g2_to_bytes(P) -> bytes:
    """x.c0||x.c1||y.c0||y.c1; infinity = 128 zero bytes."""
    if P is None:
        return bytes(128)
    (x0, x1), (y0, y1) = P
    return b"".join(v.to_bytes(32, "little") for v in (x0, x1, y0, y1))


def g2_from_bytes(b: bytes):
    if b == bytes(128):
        return None
    v = [int.from_bytes(b[i * 32:(i + 1) * 32], "little") for i in range(4)]
    return ((v[0], v[1]), (v[2], v[3]))


# --------------------------------------------------------------------------- proof.json layout (reference proof/proof.json:1-22)
def proof_to_json(proof) -> str:
    import json
    A, B, C = proof
    return json.dumps({
        "pi_a": {"x": str(A[0]), "y": str(A[1])},
        "pi_b": {"x": [str(B[0][0]), str(B[0][1])], "y": [str(B[1][0]), str(B[1][1])]},
        "pi_c": {"x": str(C[0]), "y": str(C[1])},
        "protocol": "groth16", "curve": "BN128",
    }, indent=4)


def parse_proof(json_str: str):
    """restatement of reference src/settlement/ethereum/mod.rs:445-473 (no Fp2 swap, strings only)."""
    import json
    v = json.loads(json_str)

    def dec(x):
        if not isinstance(x, str):
            raise ValueError("invalid json data")
        val = int(x, 10)
        if not (0 <= val < 1 << 256) or not x.isdigit():
            raise ValueError("invalid decimal")
        return val
    A = (dec(v["pi_a"]["x"]), dec(v["pi_a"]["y"]))
    B = ((dec(v["pi_b"]["x"][0]), dec(v["pi_b"]["x"][1])), (dec(v["pi_b"]["y"][0]), dec(v["pi_b"]["y"][1])))
    C = (dec(v["pi_c"]["x"]), dec(v["pi_c"]["y"]))
    return A, B, C


def parse_public_input(json_str: str):
    """restatement of reference src/settlement/ethereum/mod.rs:475-481 (element 0 only)."""
    import json
    v = json.loads(json_str)
    if not isinstance(v[0], str) or not v[0].isdigit():
        raise ValueError("invalid json data")
    return [int(v[0], 10)]
