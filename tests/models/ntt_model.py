"""Executable model (python ints) of the tiled multi-pass NTT implemented in csrc/ntt.cu.

Used by tests/test_ntt_model.py to pin the *index and twiddle algebra* of the CUDA kernels
(pass plan, tile addressing, DIF levels, bit-reversed write-back, inter-pass twiddles, digit
reversal of the last pass) against the oracle's best_fft, on the CPU.
"""
from oracle import bn254 as B

R_ = B.R
TILE_LOG = 11   # 2048 elements per tile
RMAX_LOG = 8


def plan(log_n):
    """pass radices (bits), most-significant digit first -- mirrors ntt.cu make_plan()."""
    if log_n <= TILE_LOG:
        return [log_n]
    p = -(-log_n // RMAX_LOG)
    base, extra = divmod(log_n, p)
    return [base + 1] * extra + [base] * (p - extra)


def brev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def tile_dif(tile, r, g, w_r):
    """in-tile DIF over axis j (size 2^r) for 2^g columns; tile[j][col]; output bit-reversed in j."""
    Rn, G = 1 << r, 1 << g
    for l in range(r):
        h = 1 << (r - 1 - l)
        for b in range((Rn // 2) * G):
            col = b & (G - 1)
            q = b >> g
            i = ((q >> (r - 1 - l)) << (r - l)) | (q & (h - 1))
            e = (q & (h - 1)) << l
            u, v = tile[i][col], tile[i + h][col]
            tile[i][col] = (u + v) % R_
            d = (u - v) % R_
            tile[i + h][col] = d if l == r - 1 else d * w_r[e] % R_


def ntt(a, omega, log_n):
    n = 1 << log_n
    radices = plan(log_n)
    P = len(radices)
    src = list(a)
    dst = [0] * n
    log_a = 0
    for t, r in enumerate(radices):
        last = t == P - 1
        Rn = 1 << r
        log_c = log_n - log_a - r
        C, A = 1 << log_c, 1 << log_a
        w_r = [pow(omega, (n // Rn) * e, R_) for e in range(max(1, Rn // 2))]
        if P == 1:
            g = 0
        else:
            g = TILE_LOG - r
        G = 1 << g
        if not last:
            ntiles = A * (C // G)
            for tid in range(ntiles):
                a_idx, cg = divmod(tid, C // G)
                c0 = cg * G
                base = a_idx * Rn * C
                tile = [[src[base + j * C + c0 + gg] for gg in range(G)] for j in range(Rn)]
                tile_dif(tile, r, g, w_r)
                for i in range(Rn):
                    k = brev(i, r)
                    for gg in range(G):
                        E = ((c0 + gg) * k) << log_a
                        assert E < n
                        dst[base + k * C + c0 + gg] = tile[i][gg] * pow(omega, E, R_) % R_
        else:
            if P == 1:
                tile = [[src[j]] for j in range(Rn)]
                tile_dif(tile, r, 0, w_r)
                for i in range(Rn):
                    dst[brev(i, r)] = tile[i][0]
            else:
                r1 = radices[0]
                rest_n = A >> r1
                ntiles = A // G
                for tid in range(ntiles):
                    rest = tid % rest_n
                    k1_0 = (tid // rest_n) * G
                    # digit-reverse `rest` = (k2, ..., k_{P-1}) msd first -> k2 + R2*k3 + ...
                    rev, tmp, shift = 0, rest, 0
                    digs = []
                    for rr in reversed(radices[1:P - 1]):
                        digs.append((tmp & ((1 << rr) - 1), rr))
                        tmp >>= rr
                    for d, rr in reversed(digs):
                        rev |= d << shift
                        shift += rr
                    tile = [[src[((k1_0 + gg) * rest_n + rest) * Rn + j] for gg in range(G)] for j in range(Rn)]
                    tile_dif(tile, r, g, w_r)
                    for i in range(Rn):
                        k = brev(i, r)
                        for gg in range(G):
                            dst[(k1_0 + gg) + ((rev + (k << (log_a - r1))) << r1)] = tile[i][gg]
        src, dst = dst, [0] * n
        log_a += r
    return src
