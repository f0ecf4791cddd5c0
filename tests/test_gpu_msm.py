"""K4-K8 parity: GPU Pippenger MSM (G1, G2) vs oracle A on small sizes + algebraic properties."""
import pytest

import oracle_a as o
from helpers import fr_bytes, g1_bytes, g1_list, g2_bytes, g2_list, rand_fr

pytestmark = pytest.mark.gpu


def _g1_points(rng, n):
    """n distinct points cheaply: running sums of a few random multiples of G."""
    base = [o.g1_mul(o.G1_GEN, rng.fr()) for _ in range(4)]
    pts, acc = [], base[0]
    for i in range(n):
        acc = o.g1_add(acc, base[i & 3])
        pts.append(acc)
    return pts


def _g2_points(rng, n):
    base = [o.g2_mul(o.G2_GEN, rng.fr()) for _ in range(4)]
    pts, acc = [], base[0]
    for i in range(n):
        acc = o.g2_add(acc, base[i & 3])
        pts.append(acc)
    return pts


@pytest.mark.parametrize("n", [1, 2, 3, 17, 64, 300])
@pytest.mark.parametrize("window", [0, 2, 5, 13, 16])
def test_msm_g1_matches_oracle(ctx, n, window):
    rng = o.Xoshiro256ss(n * 31 + window)
    pts = _g1_points(rng, n)
    sc = rand_fr(rng, n)
    ctx.set("msm_window", window)
    try:
        bases = ctx.upload_bases(g1_bytes(pts))
        got = o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc)))
    finally:
        ctx.set("msm_window", 0)
    assert got == o.msm_g1(pts, sc)


def test_msm_g1_edge_scalars(ctx):
    rng = o.Xoshiro256ss(5)
    n = 200
    pts = _g1_points(rng, n)
    pts[3] = None                       # infinity base
    pts[10] = pts[11]                   # equal bases in one bucket (doubling path)
    pts[20] = o.g1_neg(pts[21])         # P and -P in one bucket (cancellation path)
    bases = ctx.upload_bases(g1_bytes(pts))
    cases = {
        "zero": [0] * n,
        "one": [1] * n,
        "r-1": [o.R - 1] * n,
        "half_digits": [sum((1 << 15) << (16 * w) for w in range(15))] * n,
        "sparse": [rng.fr() if i % 2 else 0 for i in range(n)],
        "bits": [i & 1 for i in range(n)],
        "same": [rng.fr()] * n,
    }
    for name, sc in cases.items():
        sc = [s % o.R for s in sc]
        for w in (0, 16):
            ctx.set("msm_window", w)
            got = o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc)))
            ctx.set("msm_window", 0)
            assert got == o.msm_g1(pts, sc), (name, w)
    assert ctx.msm(bases, b"") == bytes(64)


@pytest.mark.parametrize("n", [1, 5, 130])
def test_msm_g2_matches_oracle(ctx, n):
    rng = o.Xoshiro256ss(n + 1000)
    pts = _g2_points(rng, n)
    if n > 4:
        pts[2] = None
        pts[3] = pts[4]
    sc = rand_fr(rng, n)
    sc[0] = o.R - 1
    bases = ctx.upload_bases(g2_bytes(pts), g2=True)
    assert o.g2_from_bytes(ctx.msm(bases, fr_bytes(sc))) == o.msm_g2(pts, sc)


def test_synthetic_bases_and_large_properties(ctx):
    """bases P_i = k_i G are generated on the GPU; spot-check them, then use the closed form
    sum s_i P_i = (sum s_i k_i) G to check a 2^18 MSM without a big CPU computation."""
    n, seed = 1 << 18, 0x5EED0001
    bases = ctx.synthetic_bases(n, seed)
    rng = o.Xoshiro256ss(seed)
    ks = [rng.fr() for _ in range(n)]
    head = g1_list(bases.download(0, 8))
    assert head == [o.g1_mul(o.G1_GEN, k) for k in ks[:8]]
    tail = g1_list(bases.download(n - 4, 4))
    assert tail == [o.g1_mul(o.G1_GEN, k) for k in ks[-4:]]
    srng = o.Xoshiro256ss(0x5EED0000 + 18)
    sc = [srng.fr() for _ in range(n)]
    want = o.g1_mul(o.G1_GEN, sum(s * k for s, k in zip(sc, ks)) % o.R)
    assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc))) == want
    # skewed: witness-like 0/1 scalars and all-equal scalars (one giant bucket)
    sc01 = [(i * 7) & 1 for i in range(n)]
    want = o.g1_mul(o.G1_GEN, sum(k for s, k in zip(sc01, ks) if s) % o.R)
    assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc01))) == want
    want = o.g1_mul(o.G1_GEN, 12345 * sum(ks) % o.R)
    assert o.g1_from_bytes(ctx.msm(bases, fr_bytes([12345] * n))) == want
    # sub-range + linearity: MSM over [a, b) pieces sums to the whole
    parts = [ctx.msm(bases, fr_bytes(sc[a:a + n // 4]), first=a) for a in range(0, n, n // 4)]
    assert ctx.point_sum(b"".join(parts)) == ctx.msm(bases, fr_bytes(sc))


def test_synthetic_g2_closed_form(ctx):
    n, seed = 1 << 12, 0x5EED0002
    bases = ctx.synthetic_bases(n, seed, g2=True)
    rng = o.Xoshiro256ss(seed)
    ks = [rng.fr() for _ in range(n)]
    assert g2_list(bases.download(0, 2)) == [o.g2_mul(o.G2_GEN, k) for k in ks[:2]]
    srng = o.Xoshiro256ss(99)
    sc = [srng.fr() for _ in range(n)]
    want = o.g2_mul(o.G2_GEN, sum(s * k for s, k in zip(sc, ks)) % o.R)
    assert o.g2_from_bytes(ctx.msm(bases, fr_bytes(sc))) == want


@pytest.mark.parametrize("n,window", [(1, 0), (37, 8), (300, 0), (300, 13), (300, 20)])
def test_precomputed_window_multiples_g1(ctx, n, window):
    """bases expanded to 2^{cw} P_i in HBM (zkb_bases_precompute): one shared bucket set, no Horner
    pass; results must not change."""
    rng = o.Xoshiro256ss(4000 + n + window)
    pts = _g1_points(rng, n)
    if n > 10:
        pts[3] = None
        pts[5] = pts[6]
        pts[7] = o.g1_neg(pts[8])
    bases = ctx.upload_bases(g1_bytes(pts))
    c = bases.precompute(window)
    assert c == (window or bases.window) and c >= 8
    assert g1_list(bases.download()) == pts                  # block 0 is still the original points
    for sc in (rand_fr(rng, n), [o.R - 1] * n, [1] * n, [0] * n, [(i * 5) & 1 for i in range(n)]):
        assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc))) == o.msm_g1(pts, sc)
    if n >= 150:   # sub-range of precomputed bases
        sc = rand_fr(rng, 50)
        assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc), first=100)) == o.msm_g1(pts[100:150], sc)


def test_precomputed_g2_and_large_closed_form(ctx):
    rng = o.Xoshiro256ss(4242)
    pts = _g2_points(rng, 70)
    pts[1] = None
    sc = rand_fr(rng, 70)
    bases = ctx.upload_bases(g2_bytes(pts), g2=True)
    bases.precompute(0)
    assert o.g2_from_bytes(ctx.msm(bases, fr_bytes(sc))) == o.msm_g2(pts, sc)
    n, seed = 1 << 17, 0x5EED0009
    big = ctx.synthetic_bases(n, seed)
    krng = o.Xoshiro256ss(seed)
    ks = [krng.fr() for _ in range(n)]
    sc = rand_fr(rng, n)
    want = o.g1_mul(o.G1_GEN, sum(s * k for s, k in zip(sc, ks)) % o.R)
    assert o.g1_from_bytes(ctx.msm(big, fr_bytes(sc))) == want
    big.precompute(0)
    assert o.g1_from_bytes(ctx.msm(big, fr_bytes(sc))) == want
    ctx.set("msm_host_chunks", 3)
    try:
        assert o.g1_from_bytes(ctx.msm(big, fr_bytes(sc))) == want
    finally:
        ctx.set("msm_host_chunks", 0)
    # skewed inputs on the precomputed path: every point lands in the same few buckets (one run per
    # window spans every chunk of the walk, so the whole result goes through the iterated partial
    # merge), a half-and-half mix, and r - 1 (all digits negative after the signed recoding)
    ksum = sum(ks) % o.R
    same = rng.fr()
    for chunks in (0, 3):
        ctx.set("msm_host_chunks", chunks)
        try:
            assert o.g1_from_bytes(ctx.msm(big, fr_bytes([same] * n))) == o.g1_mul(o.G1_GEN, same * ksum % o.R)
            mix = [same if i & 1 else sc[i] for i in range(n)]
            want_mix = o.g1_mul(o.G1_GEN, sum(s * k for s, k in zip(mix, ks)) % o.R)
            assert o.g1_from_bytes(ctx.msm(big, fr_bytes(mix))) == want_mix
            assert o.g1_from_bytes(ctx.msm(big, fr_bytes([o.R - 1] * n))) == o.g1_mul(o.G1_GEN, (o.R - ksum) % o.R)
        finally:
            ctx.set("msm_host_chunks", 0)


# ---------------------------------------------------------------------------------------------
# The benchmarked configuration (bench.py: 2^24 points, precomputed window multiples, c = 22) has
# 2^21 buckets in ONE shared set, which takes the two-level bucket reduction (red_levels = 2,
# msm_common.cu) with the `mul_small` fold of the last level (msm_impl.cuh).  Windows >= 21 reach
# that path; small n keeps the oracle-A comparison cheap.
@pytest.mark.parametrize("window", [21, 22, 23, 24])
def test_wide_windows_two_level_bucket_reduction_g1(ctx, window):
    n = 300
    rng = o.Xoshiro256ss(7000 + window)
    pts = _g1_points(rng, n)
    pts[3] = None
    pts[5] = pts[6]
    pts[7] = o.g1_neg(pts[8])
    bases = ctx.upload_bases(g1_bytes(pts))
    assert bases.precompute(window) == window
    top = (1 << (window - 1))            # the largest digit magnitude: last bucket of the set
    cases = [rand_fr(rng, n), [o.R - 1] * n, [1] * n, [top] * n, [top - 1] * n, [(top << window) | top] * n,
             [(i * 5) & 1 for i in range(n)], [0] * n]
    for sc in cases:
        sc = [s % o.R for s in sc]
        assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc))) == o.msm_g1(pts, sc), window
    for chunks in (2, 3):                # the chunked host-scalar path merges bucket sets
        ctx.set("msm_host_chunks", chunks)
        try:
            assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(cases[0]))) == o.msm_g1(pts, cases[0])
        finally:
            ctx.set("msm_host_chunks", 0)
    bases.free()


@pytest.mark.parametrize("window", [21, 23])
def test_wide_windows_two_level_bucket_reduction_g2(ctx, window):
    n = 90
    rng = o.Xoshiro256ss(7100 + window)
    pts = _g2_points(rng, n)
    pts[1] = None
    pts[4] = pts[5]
    bases = ctx.upload_bases(g2_bytes(pts), g2=True)
    assert bases.precompute(window) == window
    top = 1 << (window - 1)
    for sc in (rand_fr(rng, n), [o.R - 1] * n, [top] * n, [(i * 3) & 1 for i in range(n)]):
        assert o.g2_from_bytes(ctx.msm(bases, fr_bytes(sc))) == o.msm_g2(pts, sc), window
    bases.free()


@pytest.mark.parametrize("log_n,window", [(22, 22), (24, 0)])
def test_headline_msm_closed_form(ctx, log_n, window):
    """The exact path bench.py times — 2^24 synthetic bases, zkb_bases_precompute (auto window = 22),
    uniform scalars, device-resident AND host-scalar entry points — checked against the closed form
    (sum s_i k_i mod r) * G computed by oracle B (one dot product over Fr + one scalar multiplication)."""
    import torch
    import oracle_b as ob
    n, seed = 1 << log_n, 0x5EED0001 + log_n
    bases = ctx.synthetic_bases(n, seed)
    c = bases.precompute(window)
    assert c == 22
    sc = ob.random_scalars(n, 0x5EED0000 + log_n)
    want = ob.msm_closed_form(sc, n, seed)
    host = torch.frombuffer(bytearray(sc), dtype=torch.uint8).pin_memory()
    dev = host.cuda()
    out = torch.zeros(64, dtype=torch.uint8, device="cuda")
    ctx.msm_dev(bases, dev.data_ptr(), n, out.data_ptr())
    ctx.sync()
    assert bytes(out.cpu().numpy()) == want                     # bench.py `value` path
    assert ctx.msm_raw(bases, host.data_ptr(), n) == want       # bench.py `e2e` path (3 overlapped chunks)
    # witness-like skew on the same bases: 0/1 scalars and one giant bucket
    ks = ob.random_scalars(n, seed)
    bits = bytearray(32 * n)
    bits[0::64] = b"\x01" * (n // 2)                            # every even-indexed scalar is 1
    t_bits = torch.frombuffer(bits, dtype=torch.uint8)
    assert ctx.msm_raw(bases, t_bits.data_ptr(), n) == ob.msm_closed_form(bytes(bits), n, seed)
    same = (0x1234567 << 200 | 99).to_bytes(32, "little")
    ksum = ob.fr_sum(ks, n)
    want_same = ob.g1_mul_gen(ob.fr_dot(same, ksum, 1))
    all_same = bytearray(same * n)
    t_same = torch.frombuffer(all_same, dtype=torch.uint8)
    assert ctx.msm_raw(bases, t_same.data_ptr(), n) == want_same
    bases.free()


def test_g2_msm_closed_form_2p20(ctx):
    import oracle_b as ob
    n, seed = 1 << 20, 0x5EED0777
    bases = ctx.synthetic_bases(n, seed, g2=True)
    sc = ob.random_scalars(n, 4242)
    want = ob.msm_closed_form(sc, n, seed, g2=True)
    assert ctx.msm(bases, sc) == want
    bases.precompute(0)
    assert ctx.msm(bases, sc) == want
    bases.free()


@pytest.mark.parametrize("mode", [0, 1])
def test_both_sort_implementations_agree(ctx, mode):
    """msm_sort = 0: the library's own bucket sort fused with the digit extraction (default);
    msm_sort = 1: digits kernel + cub::DeviceRadixSort (comparison baseline).  Same results on uniform,
    0/1, all-equal (every entry in a handful of buckets -> the multi-segment path of the bucket sort)
    and r - 1 scalars, classic and precomputed bases, G1 and G2."""
    import oracle_b as ob
    ctx.set("msm_sort", mode)
    try:
        rng = o.Xoshiro256ss(8800)
        pts = _g1_points(rng, 257)
        pts[9] = None
        sc = rand_fr(rng, 257)
        bases = ctx.upload_bases(g1_bytes(pts))
        for w in (0, 4, 11, 16):
            ctx.set("msm_window", w)
            assert o.g1_from_bytes(ctx.msm(bases, fr_bytes(sc))) == o.msm_g1(pts, sc), w
        ctx.set("msm_window", 0)
        n, seed = 1 << 18, 0x5EED0123
        big = ctx.synthetic_bases(n, seed)
        cases = {"uniform": ob.random_scalars(n, 5), "bits": b"".join(((i * 3) & 1).to_bytes(32, "little") for i in range(n)),
                 "same": (0xABCDEF << 180 | 77).to_bytes(32, "little") * n, "r-1": int(o.R - 1).to_bytes(32, "little") * n}
        for precomp in (False, True):
            if precomp:
                big.precompute(0)
            for name, s in cases.items():
                assert ctx.msm(big, s) == ob.msm_closed_form(s, n, seed), (name, precomp)
        g2 = ctx.synthetic_bases(1 << 12, seed + 1, g2=True)
        s = ob.random_scalars(1 << 12, 6)
        assert ctx.msm(g2, s) == ob.msm_closed_form(s, 1 << 12, seed + 1, g2=True)
    finally:
        ctx.set("msm_sort", 0)
        ctx.set("msm_window", 0)
