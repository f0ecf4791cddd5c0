"""Word-level model (32-bit limbs, explicit carry flag) of the Montgomery squaring in csrc/ff.cuh
(mont_sqr): used to validate the carry-chain schedule before it was transcribed to PTX."""
import random, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
from oracle_a import Q, R

M32 = 0xFFFFFFFF


class Cpu:
    def __init__(self): self.cf = 0
    def mul_lo(self, a, b): return (a * b) & M32
    def mul_hi(self, a, b): return (a * b) >> 32
    def _add(self, x, y, cin, setcc):
        s = x + y + cin
        if setcc: self.cf = s >> 32
        return s & M32
    def add_cc(self, x, y): return self._add(x, y, 0, True)
    def addc_cc(self, x, y): return self._add(x, y, self.cf, True)
    def addc(self, x, y): return self._add(x, y, self.cf, False)
    def mad_lo_cc(self, a, b, c): return self._add(self.mul_lo(a, b), c, 0, True)
    def madc_lo_cc(self, a, b, c): return self._add(self.mul_lo(a, b), c, self.cf, True)
    def madc_hi_cc(self, a, b, c): return self._add(self.mul_hi(a, b), c, self.cf, True)
    def madc_hi(self, a, b, c): return self._add(self.mul_hi(a, b), c, self.cf, False)
    def mad_hi_cc(self, a, b, c): return self._add(self.mul_hi(a, b), c, 0, True)


def limbs(x, n=8): return [(x >> (32 * i)) & M32 for i in range(n)]
def val(ws): return sum(w << (32 * i) for i, w in enumerate(ws))


def square_wide(c, a):
    """16-word square of an 8-limb number: off-diagonal products in two lanes, doubled, + diagonal."""
    O = [0] * 16
    E = [0] * 16
    def init(L, prods):           # disjoint 64-bit slots, no carries
        for (i, j, pos) in prods:
            L[pos], L[pos + 1] = c.mul_lo(a[i], a[j]), c.mul_hi(a[i], a[j])
    def chain(L, prods, top):     # consecutive slots, then ripple the carry up to `top`
        first = True
        for (i, j, pos) in prods:
            L[pos] = c.mad_lo_cc(a[i], a[j], L[pos]) if first else c.madc_lo_cc(a[i], a[j], L[pos])
            L[pos + 1] = c.madc_hi_cc(a[i], a[j], L[pos + 1])
            first = False
        k = prods[-1][2] + 2
        while k < top:
            L[k] = c.addc_cc(L[k], 0); k += 1
        L[top] = c.addc(L[top], 0)
    init(O, [(0, 1, 1), (0, 3, 3), (0, 5, 5), (0, 7, 7), (2, 7, 9), (4, 7, 11), (6, 7, 13)])
    chain(O, [(1, 2, 3), (1, 4, 5), (1, 6, 7), (3, 6, 9), (5, 6, 11)], 14)
    chain(O, [(2, 3, 5), (2, 5, 7), (4, 5, 9)], 14)
    chain(O, [(3, 4, 7)], 14)
    init(E, [(0, 2, 2), (0, 4, 4), (0, 6, 6), (1, 7, 8), (3, 7, 10), (5, 7, 12)])
    chain(E, [(1, 3, 4), (1, 5, 6), (2, 6, 8), (4, 6, 10)], 13)
    chain(E, [(2, 4, 6), (3, 5, 8)], 13)
    # S = O + E  (words 1..14)
    S = [0] * 16
    S[1] = O[1]
    S[2] = c.add_cc(O[2], E[2])
    for k in range(3, 14):
        S[k] = c.addc_cc(O[k], E[k])
    S[14] = c.addc(O[14], 0)
    # T = 2 S (funnel shifts) + diagonal
    T = [0] * 16
    for k in range(1, 16):
        T[k] = ((S[k] << 1) | (S[k - 1] >> 31)) & M32
    T[0] = c.mul_lo(a[0], a[0])
    T[1] = c.mad_hi_cc(a[0], a[0], T[1])
    for i in range(1, 8):
        T[2 * i] = c.madc_lo_cc(a[i], a[i], T[2 * i])
        T[2 * i + 1] = c.madc_hi_cc(a[i], a[i], T[2 * i + 1]) if i < 7 else c.madc_hi(a[i], a[i], T[2 * i + 1])
    return T


def redc_lo(c, t, p, m0):
    """U = (t + m p) / 2^256 for an 8-word t, two-lane interleaved (the mont_mul schedule without a*b)."""
    pe, po = p[0::2], p[1::2]
    e, o = list(t), [0] * 8
    def lane_mad(L, q, m):        # L += q (x) m, no carry out
        L[0] = c.mad_lo_cc(q[0], m, L[0]); L[1] = c.madc_hi_cc(q[0], m, L[1])
        for k in range(1, 4):
            L[2 * k] = c.madc_lo_cc(q[k], m, L[2 * k])
            L[2 * k + 1] = c.madc_hi_cc(q[k], m, L[2 * k + 1]) if k < 3 else c.madc_hi(q[k], m, L[2 * k + 1])
    def lane_mad_spill(L, other, q, m):
        L[0] = c.mad_lo_cc(q[0], m, L[0]); L[1] = c.madc_hi_cc(q[0], m, L[1])
        for k in range(1, 4):
            L[2 * k] = c.madc_lo_cc(q[k], m, L[2 * k]); L[2 * k + 1] = c.madc_hi_cc(q[k], m, L[2 * k + 1])
        other[7] = c.addc(other[7], 0)
    def lane_shift_mad(x, y, q, m):
        x[0] = c.add_cc(x[0], y[1])
        y0 = c.madc_lo_cc(q[0], m, y[2]); y1 = c.madc_hi_cc(q[0], m, y[3])
        y2 = c.madc_lo_cc(q[1], m, y[4]); y3 = c.madc_hi_cc(q[1], m, y[5])
        y4 = c.madc_lo_cc(q[2], m, y[6]); y5 = c.madc_hi_cc(q[2], m, y[7])
        y6 = c.madc_lo_cc(q[3], m, 0); y7 = c.madc_hi(q[3], m, 0)
        y[:] = [y0, y1, y2, y3, y4, y5, y6, y7]
    m = (e[0] * m0) & M32
    lane_mad(o, po, m); lane_mad_spill(e, o, pe, m)
    x, y = o, e
    for _ in range(7):
        m = (((x[0] + y[1]) & M32) * m0) & M32
        lane_shift_mad(x, y, po, m); lane_mad_spill(x, y, pe, m)
        x, y = y, x
    # after 8 rounds the lane with a zero low word is y (= the last x); result = x + (y >> 32)
    assert y[0] == 0
    r = [0] * 8
    r[0] = c.add_cc(x[0], y[1])
    for k in range(1, 7):
        r[k] = c.addc_cc(x[k], y[k + 1])
    r[7] = c.addc(x[7], 0)
    return r


def mont_sqr(a_int, p_int):
    c = Cpu()
    m0 = (-pow(p_int, -1, 1 << 32)) % (1 << 32)
    T = square_wide(c, limbs(a_int))
    assert val(T) == a_int * a_int
    U = redc_lo(c, T[:8], limbs(p_int), m0)
    assert (val(U) << 256) == val(T[:8]) + ((val(T[:8]) * ((-pow(p_int, -1, 1 << 256)) % (1 << 256))) % (1 << 256)) * p_int
    r = [0] * 8
    r[0] = c.add_cc(U[0], T[8])
    for k in range(1, 7):
        r[k] = c.addc_cc(U[k], T[8 + k])
    r[7] = c.addc(U[7], T[15])
    return val(r)


if __name__ == "__main__":
    rnd = random.Random(1)
    for p in (Q, R):
        Rinv = pow(1 << 256, -1, p)
        cases = [0, 1, p - 1, p, p + 1, 2 * p - 1, (1 << 255) - 1 if (1 << 255) - 1 < 2 * p else 2 * p - 2, 0xFFFFFFFF, (1 << 224) - 1]
        cases += [rnd.randrange(2 * p) for _ in range(3000)]
        for a in cases:
            r = mont_sqr(a, p)
            assert r < 2 * p and r % p == a * a * Rinv % p, hex(a)
    print("mont_sqr model ok")
