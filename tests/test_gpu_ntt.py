"""K9-K11 parity: GPU Fr NTT vs oracle A (ark-poly Radix2EvaluationDomain semantics)."""
import pytest

import oracle_a as o
from helpers import fr_bytes, fr_list, rand_fr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13])
def test_ntt_matches_oracle(ctx, log_n):
    rng = o.Xoshiro256ss(0x5EED0000 + log_n)
    v = rand_fr(rng, 1 << log_n)
    if log_n >= 3:
        v[0], v[1], v[2] = 0, o.R - 1, 1
    buf = fr_bytes(v)
    assert fr_list(ctx.fr_ntt(buf)) == o.ntt(v)
    assert fr_list(ctx.fr_ntt(buf, inverse=True)) == o.ntt(v, inverse=True)
    assert fr_list(ctx.fr_ntt(buf, coset=True)) == o.coset_ntt(v)
    assert fr_list(ctx.fr_ntt(buf, inverse=True, coset=True)) == o.coset_intt(v)
    if log_n >= 9:          # multi-pass sizes: also the two-level inter-pass twiddle path
        ctx.set("ntt_full_twiddles", 0)
        try:
            assert fr_list(ctx.fr_ntt(buf)) == o.ntt(v)
            assert fr_list(ctx.fr_ntt(buf, inverse=True)) == o.ntt(v, inverse=True)
        finally:
            ctx.set("ntt_full_twiddles", 1)


@pytest.mark.parametrize("full", [0, 1])
@pytest.mark.parametrize("log_n", [14, 16, 17, 19, 20, 22])
def test_ntt_properties_large(ctx, log_n, full):
    """size-independent checks: round trips, linearity against a delta, the constant vector."""
    ctx.set("ntt_full_twiddles", full)
    try:
        n = 1 << log_n
        rng = o.Xoshiro256ss(log_n)
        v = rand_fr(rng, 257) * (n // 257 + 1)
        v = v[:n]
        buf = fr_bytes(v)
        f = ctx.fr_ntt(buf)
        assert ctx.fr_ntt(f, inverse=True) == buf
        c = ctx.fr_ntt(buf, coset=True)
        assert ctx.fr_ntt(c, inverse=True, coset=True) == buf
        # X[0] = sum x[i];  X[n/2] = sum (-1)^i x[i]
        assert fr_list(f[:32]) == [sum(v) % o.R]
        alt = (sum(v[0::2]) - sum(v[1::2])) % o.R
        assert fr_list(f[32 * (n // 2): 32 * (n // 2) + 32]) == [alt]
        # delta at position j -> powers of w^j
        j = 12345 % n
        d = bytearray(32 * n)
        d[32 * j: 32 * j + 32] = (1).to_bytes(32, "little")
        fd = ctx.fr_ntt(bytes(d))
        w = pow(o.fr_root_of_unity(log_n), j, o.R)
        for k in (0, 1, 2, 3, n // 2 + 1, n - 1, 77 % n):
            assert fr_list(fd[32 * k: 32 * k + 32]) == [pow(w, k, o.R)]
    finally:
        ctx.set("ntt_full_twiddles", 1)


def test_ntt_rejects_bad_input(ctx, zkb):
    with pytest.raises(zkb.ZkbError):
        ctx.fr_ntt(int(o.R).to_bytes(32, "little") * 4)
    with pytest.raises(ValueError):
        ctx.fr_ntt(bytes(32 * 3))


def test_launch_counter_counts_every_pass(ctx):
    """zkb_ctx_launches feeds bench.py's gpu_launches claim: a resident Montgomery-form transform of
    2^16 is exactly ceil(16/8) = 2 pass kernels once the tables exist."""
    import torch
    n = 1 << 16
    x = torch.zeros(n * 32, dtype=torch.uint8, device="cuda")
    ctx.fr_ntt_dev(x.data_ptr(), x.data_ptr(), 16, montgomery=True)     # builds the tables
    ctx.sync()
    before = ctx.launches
    ctx.fr_ntt_dev(x.data_ptr(), x.data_ptr(), 16, montgomery=True)
    ctx.sync()
    assert ctx.launches - before == 2


@pytest.mark.parametrize("log_n", [16, 20, 24, 25, 26])
def test_ntt_vs_oracle_b_large(ctx, log_n):
    """The metric size (2^24, three passes 8+8+8) and the four-pass plans (2^25 = 7+6+6+6, 2^26) against
    oracle B's ark-poly restatement, byte for byte; the inverse coset transform closes the loop."""
    import oracle_b as ob
    n = 1 << log_n
    buf = ob.random_scalars(n, 0x5EED0002 + log_n)
    want = ob.fr_ntt(buf)
    got = ctx.fr_ntt(buf)
    assert got == want
    del want
    if log_n <= 24:
        assert ctx.fr_ntt(buf, inverse=True, coset=True) == ob.fr_ntt(buf, inverse=True, coset=True)
        ctx.set("ntt_full_twiddles", 0)
        try:
            assert ctx.fr_ntt(buf) == got
        finally:
            ctx.set("ntt_full_twiddles", 1)
    else:
        assert ctx.fr_ntt(got, inverse=True) == buf
