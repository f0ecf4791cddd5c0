"""Word-level model of mont_mul2 (csrc/ff.cuh): r = (a*b + c*d) / 2^256 mod p in ONE two-lane CIOS pass
(128 product + 72 reduction IMAD.WIDE instead of 2 x 136).  Checks, for worst-case operands (< 2p and
all-ones words), that no lane outgrows its 8-limb frame, that the carry-outs the schedule drops are
zero, and that the result is the Montgomery product sum and stays below 2.51 p (so one fold2p brings
it back under 2p).  Usage: python tools/mul2_model.py"""
import random

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
W = 1 << 32


def limbs(x):
    return [(x >> (32 * i)) & (W - 1) for i in range(8)]


def lane_val(l):
    return sum(v << (32 * i) for i, v in enumerate(l))


def set_lane(x):
    assert 0 <= x < 1 << 256, "lane overflow"
    return limbs(x)


def mont_mul2(p, a, b, c, d):
    m0 = (-pow(p, -1, W)) % W
    A, B, C, D, Pm = limbs(a), limbs(b), limbs(c), limbs(d), limbs(p)
    ev = lambda X: sum(X[j] << (32 * j) for j in (0, 2, 4, 6))          # even limbs at their positions
    od = lambda X: sum(X[j] << (32 * (j - 1)) for j in (1, 3, 5, 7))    # odd limbs, lane frame starts at position 1
    # step 0: e @ positions 0..7, o @ positions 1..8
    e = set_lane(ev(A) * B[0])
    o = set_lane(od(A) * B[0])
    o = set_lane(lane_val(o) + od(C) * D[0])                             # lane_mad: must fit
    t = lane_val(e) + ev(C) * D[0]                                       # lane_mad_spill: carry -> o[7]
    e = limbs(t & ((1 << 256) - 1)); o[7] += t >> 256; assert o[7] < W
    def redc(x, y):                                                      # x even-aligned lane, y the other
        m = (x[0] * m0) % W
        y2 = set_lane(lane_val(y) + od(Pm) * m)
        t = lane_val(x) + ev(Pm) * m
        x2 = limbs(t & ((1 << 256) - 1)); y2[7] += t >> 256; assert y2[7] < W
        assert x2[0] == 0
        return x2, y2
    e, o = redc(e, o)
    x, y = o, e          # x: lane that becomes even-aligned after the division by 2^32, y: the one shifted
    for i in range(1, 8):
        # lane_shift_mad: x[0] += y[1] (carry into the chain), y_new = (y[2..7],0,0) + a_odd*b_i  [positions 1..8]
        s = x[0] + y[1]
        x[0] = s & (W - 1)
        y = set_lane(lane_val(y[2:] + [0, 0]) + od(A) * B[i] + (s >> 32))
        t = lane_val(x) + ev(A) * B[i]
        x = limbs(t & ((1 << 256) - 1)); y[7] += t >> 256; assert y[7] < W
        y = set_lane(lane_val(y) + od(C) * D[i])
        t = lane_val(x) + ev(C) * D[i]
        x = limbs(t & ((1 << 256) - 1)); y[7] += t >> 256; assert y[7] < W
        x, y = redc(x, y)
        x, y = y, x
    # after the last swap: y is the lane with the zero low limb (positions 0..7), x sits at positions 1..8
    assert y[0] == 0
    r = lane_val(x) + lane_val(y[1:] + [0])
    assert r < 1 << 256
    return r


def main():
    rng = random.Random(5)
    for p in (Q, R):
        inv = pow(1 << 256, -1, p)
        worst = 0
        cases = [(2 * p - 1,) * 4, (2 * p - 1, 2 * p - 1, 2 * p, 2 * p - 1), (0, 0, 0, 0), (1, 1, 1, 1),
                 (2 * p, 2 * p, 2 * p, 2 * p), (2 * p - 1, (1 << 32) - 1, 2 * p - 1, (1 << 64) - 1)]
        cases += [tuple(rng.randrange(2 * p) for _ in range(4)) for _ in range(20000)]
        for a, b, c, d in cases:
            r = mont_mul2(p, a, b, c, d)
            assert r % p == (a * b + c * d) * inv % p
            worst = max(worst, r / p)
            assert r < 2.52 * p
        print("p = %s...: %d cases ok, largest result %.4f p" % (str(p)[:8], len(cases), worst))


if __name__ == "__main__":
    main()
