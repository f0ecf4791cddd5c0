"""Index-level Python model of csrc/ntt.cu (pass plan, tile maps, stage twiddles, digit reversal).

Development aid only: it lets the kernel's index arithmetic be checked against oracle_a.ntt on a
machine without a GPU.  Not imported by the product.
"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import oracle_a as o

R = o.R


def plan(L):
    if L <= 8:
        return [L]
    P = (L + 7) // 8
    base, rem = divmod(L, P)
    return [base + 1 if i < rem else base for i in range(P)]


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def warp_ntt(slots, k, root256):
    """slots: 256 values, u = g*2^k + t ; DIF stages b = k-1..0 on bit b ; then bitrev store."""
    x = list(slots)
    for b in range(7, -1, -1):
        if b >= k:
            continue
        wb = pow(root256, 1 << (7 - b), R)     # w_{2^{b+1}}
        for u in range(256):
            if not (u >> b) & 1:
                i = u & ((1 << b) - 1)
                a, c = x[u], x[u | (1 << b)]
                x[u] = (a + c) % R
                x[u | (1 << b)] = (a - c) * pow(wb, i, R) % R
    out = [0] * 256
    tmask = (1 << k) - 1
    for u in range(256):
        out[(u & ~tmask) | bitrev(u & tmask, k)] = x[u]
    return out


def ntt_model(a, inverse=False):
    n = len(a)
    L = n.bit_length() - 1
    ks = plan(L)
    P = len(ks)
    wn = o.fr_root_of_unity(L)
    root256 = o.fr_root_of_unity(8)
    if inverse:
        wn, root256 = o.inv(wn, R), o.inv(root256, R)
    buf = list(a)
    M = n
    for p in range(P - 1):
        k = ks[p]
        S = M >> k
        CT = 8 << (8 - k)
        gb = 8 - k
        tiles_per_blk = (S + CT - 1) // CT
        for tile in range((n // M) * tiles_per_blk):
            blk, tt = divmod(tile, tiles_per_blk)
            r_base = tt * CT
            base = blk * M + r_base
            ncols = min(CT, S - r_base)
            warps = [[0] * 256 for _ in range(8)]
            for q in range(2048):
                c, t = q % CT, q // CT
                if c < ncols:
                    warps[c >> gb][((c & ((1 << gb) - 1)) << k) | t] = buf[base + t * S + c]
            warps = [warp_ntt(w, k, root256) for w in warps]
            for q in range(2048):
                c, K = q % CT, q // CT
                if c >= ncols:
                    continue
                y = warps[c >> gb][((c & ((1 << gb) - 1)) << k) | K]
                e = ((r_base + c) * K) << (L - (M.bit_length() - 1))
                buf[base + K * S + c] = y * pow(wn, e, R) % R
        M >>= k
    # final pass
    k = ks[-1]
    row_bits = L - k
    first = ks[0] if P > 1 else 0
    mid_bits = row_bits - first
    CT = 8 << (8 - k)
    gb = 8 - k
    out = [0] * n
    k0_range = 1 << first
    tiles_per_mid = (k0_range + CT - 1) // CT
    digits = ks[:-1]

    def rowrev(rho):
        res, sh_in, sh_out = 0, row_bits, 0
        for w in digits:
            sh_in -= w
            res |= ((rho >> sh_in) & ((1 << w) - 1)) << sh_out
            sh_out += w
        return res
    for tile in range(tiles_per_mid << mid_bits):
        mid, tt = divmod(tile, tiles_per_mid)
        c_base = tt * CT
        ncols = min(CT, k0_range - c_base)
        warps = [[0] * 256 for _ in range(8)]
        for q in range(2048):
            c, t = q % CT, q // CT
            if c < ncols:
                warps[c >> gb][((c & ((1 << gb) - 1)) << k) | t] = buf[(((c_base + c) << mid_bits) + mid) * (1 << k) + t]
        warps = [warp_ntt(w, k, root256) for w in warps]
        for q in range(2048):
            c, K = q % CT, q // CT
            if c >= ncols:
                continue
            rho = ((c_base + c) << mid_bits) + mid
            out[rowrev(rho) + (K << row_bits)] = warps[c >> gb][((c & ((1 << gb) - 1)) << k) | K]
    if inverse:
        ninv = o.inv(n, R)
        out = [x * ninv % R for x in out]
    return out


if __name__ == "__main__":
    rng = o.Xoshiro256ss(7)
    for L in [1, 2, 3, 5, 8, 9, 10, 12]:
        v = [rng.fr() for _ in range(1 << L)]
        assert ntt_model(v) == o.ntt(v), L
        assert ntt_model(v, True) == o.ntt(v, True), L
        print("L", L, plan(L), "ok")
