"""K1-K3 parity: device Montgomery field library and XYZZ curve ops vs Python big ints (oracle A)."""
import pytest

import oracle_a as o
from helpers import fr_bytes, fr_list, g1_bytes, g1_list, g2_bytes, g2_list, rand_fq, rand_fr

pytestmark = pytest.mark.gpu

OPS = {0: lambda a, b, p: (a + b) % p, 1: lambda a, b, p: (a - b) % p, 2: lambda a, b, p: a * b % p,
       3: lambda a, b, p: a * a % p, 4: lambda a, b, p: pow(a, p - 2, p), 5: lambda a, b, p: (-a) % p,
       6: lambda a, b, p: a * a % p,
       # fused dual products (mont_mul2: two products, one Montgomery reduction)
       7: lambda a, b, p: (a * b + (a + b) * (a - b)) % p, 8: lambda a, b, p: (a * b - (a + b) * (a - b)) % p,
       # op 4 = binary extended Euclid (what the kernels call), op 9 = Fermat exponentiation
       9: lambda a, b, p: pow(a, p - 2, p)}


@pytest.mark.parametrize("field", [0, 1])
def test_field_ops_bit_exact(ctx, field):
    p = o.Q if field == 0 else o.R
    rng = o.Xoshiro256ss(100 + field)
    edge = [0, 1, 2, p - 1, p - 2, (1 << 256) % p, (p + 1) // 2, (1 << 253), (1 << 32) - 1, (1 << 64), p >> 1]
    rnd = rand_fq(rng, 2000) if field == 0 else rand_fr(rng, 2000)
    a = edge + rnd + [x for x in edge for _ in edge]
    b = edge[::-1] + rnd[::-1] + [y for _ in edge for y in edge]
    assert len(a) == len(b)
    for op, f in OPS.items():
        got = fr_list(ctx.field_op(field, op, fr_bytes(a), fr_bytes(b)))
        want = [f(x, y, p) for x, y in zip(a, b)]
        assert got == want, "field %d op %d" % (field, op)


def test_field_rejects_nothing_silently(ctx, zkb):
    # a non-canonical scalar must be rejected at the ABI (error behaviour mirrors U256 range checks)
    bases = ctx.upload_bases(g1_bytes([o.G1_GEN]))
    with pytest.raises(zkb.ZkbError):
        ctx.msm(bases, int(o.R).to_bytes(32, "little"))
    with pytest.raises(zkb.ZkbError):
        ctx.upload_bases(int(o.Q).to_bytes(32, "little") + bytes(32))


def test_g1_ops(ctx):
    rng = o.Xoshiro256ss(7)
    ks = [rng.fr() for _ in range(24)]
    P = [o.g1_mul(o.G1_GEN, k) for k in ks]
    S = P[1:] + P[:1]
    # generic add, doubling (P+P), cancellation (P + -P), infinity operands
    a = P + P[:4] + P[:4] + [None, P[0], None]
    b = S + P[:4] + [o.g1_neg(x) for x in P[:4]] + [P[1], None, None]
    got = g1_list(ctx.curve_op(0, 0, g1_bytes(a), g1_bytes(b)))
    assert got == [o.g1_add(x, y) for x, y in zip(a, b)]
    got = g1_list(ctx.curve_op(0, 1, g1_bytes(P + [None])))
    assert got == [o.g1_add(x, x) for x in P] + [None]
    sc = [0, 1, 2, o.R - 1, o.R - 2] + [rng.fr() for _ in range(11)]
    got = g1_list(ctx.curve_op(0, 2, g1_bytes(P[:16]), fr_bytes(sc)))
    assert got == [o.g1_mul(x, k) for x, k in zip(P[:16], sc)]


def test_g2_ops(ctx):
    rng = o.Xoshiro256ss(8)
    P = [o.g2_mul(o.G2_GEN, rng.fr()) for _ in range(10)]
    S = P[1:] + P[:1]
    a = P + P[:3] + P[:3] + [None, P[0]]
    b = S + P[:3] + [o.g2_neg(x) for x in P[:3]] + [P[1], None]
    got = g2_list(ctx.curve_op(1, 0, g2_bytes(a), g2_bytes(b)))
    assert got == [o.g2_add(x, y) for x, y in zip(a, b)]
    got = g2_list(ctx.curve_op(1, 1, g2_bytes(P)))
    assert got == [o.g2_add(x, x) for x in P]
    sc = [0, 1, o.R - 1] + [rng.fr() for _ in range(5)]
    got = g2_list(ctx.curve_op(1, 2, g2_bytes(P[:8]), fr_bytes(sc)))
    assert got == [o.g2_mul(x, k) for x, k in zip(P[:8], sc)]


def test_abi_error_behaviour(ctx, zkb):
    """argument validation at the C ABI: wrong group, out-of-range sub-ranges, unknown tunables,
    one call at a time per context (ZKB_EBUSY)."""
    import ctypes as C
    import threading
    g1 = ctx.upload_bases(g1_bytes([o.G1_GEN] * 4))
    g2 = ctx.upload_bases(g2_bytes([o.G2_GEN] * 2), g2=True)
    out = C.create_string_buffer(128)
    sc = C.create_string_buffer(fr_bytes([1, 2, 3, 4]))
    lib = ctx.lib
    assert lib.zkb_msm_g2(ctx.h, g1.handle, 0, sc, 4, out) == zkb.lib.ZKB_EINVAL          # G1 bases to the G2 entry
    assert lib.zkb_msm_g1(ctx.h, g2.handle, 0, sc, 2, out) == zkb.lib.ZKB_EINVAL
    assert lib.zkb_msm_g1(ctx.h, g1.handle, 3, sc, 2, out) == zkb.lib.ZKB_EINVAL          # range past the end
    assert lib.zkb_msm_g1(ctx.h, g1.handle, 0, None, 2, out) == zkb.lib.ZKB_EINVAL
    assert lib.zkb_ctx_set(ctx.h, b"no_such_knob", 1) == zkb.lib.ZKB_EINVAL
    assert b"unknown tunable" in lib.zkb_last_error(ctx.h)
    assert lib.zkb_fr_ntt(ctx.h, sc, 0, 0, 0) == zkb.lib.ZKB_EINVAL
    assert lib.zkb_fr_ntt(ctx.h, sc, 29, 0, 0) == zkb.lib.ZKB_EINVAL
    with pytest.raises(zkb.ZkbError):
        g1.precompute(1)
    g1.precompute(12)
    with pytest.raises(zkb.ZkbError):
        g1.precompute(12)                                                                 # already expanded
    # re-entrancy: a second thread calling into the same context while a long call runs gets EBUSY
    big = ctx.synthetic_bases(1 << 18, 9)
    scal = fr_bytes(rand_fr(o.Xoshiro256ss(3), 1 << 18))
    seen = []

    def hammer():
        for _ in range(2000):
            rc = lib.zkb_ctx_sync(ctx.h)       # cheap entry point without the guard
            rc = lib.zkb_bench_int_pipe(ctx.h, 1, (C.c_double * 3)())
            seen.append(rc)
            if rc == zkb.lib.ZKB_EBUSY:
                return

    t = threading.Thread(target=hammer)
    t.start()
    for _ in range(3):
        ctx.msm(big, scal)
    t.join()
    assert zkb.lib.ZKB_EBUSY in seen or all(rc == 0 for rc in seen)
    assert o.g1_from_bytes(ctx.msm(g1, fr_bytes([1, 2, 3, 4]))) == o.g1_mul(o.G1_GEN, 10)   # context still healthy
