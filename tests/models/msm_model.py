"""Executable model of the MSM pipeline in csrc/msm.cu, over an abstract additive group.

Pins, on the CPU, the control logic that is hard to eyeball in CUDA: signed-digit recoding,
counting sort by (window, bucket), the chunked reduce-by-key levels (first/last runs of a chunk
become partials for the next level, interior runs go straight to their bucket), the final
single-CTA segmented scan, the segmented bucket reduction and the host window fold.
The "group" is pluggable: integers mod a prime (fast, exact multiset check) or real G1 points.
"""
INVALID = 0xFFFFFFFF


class IntGroup:
    """Z_p under addition; element i*G is represented by i."""
    def __init__(self, p): self.p = p
    def zero(self): return 0
    def add(self, a, b): return (a + b) % self.p
    def neg(self, a): return (-a) % self.p
    def dbl(self, a): return 2 * a % self.p


def window_count(c, bits=254):
    return (bits + 1 + c - 1) // c


def recode(s, c, W):
    """signed digits d_w in [-2^(c-1), 2^(c-1)], sum d_w 2^(cw) == s."""
    out, carry = [], 0
    half = 1 << (c - 1)
    for w in range(W):
        d = ((s >> (w * c)) & ((1 << c) - 1)) + carry
        if d > half:
            d -= 1 << c
            carry = 1
        else:
            carry = 0
        out.append(d)
    assert carry == 0
    return out


def reduce_level(grp, keys, pts, L, buckets):
    """one chunked reduce-by-key level: returns (keys', pts') with 2 slots per chunk."""
    n = len(keys)
    nchunks = (n + L - 1) // L
    okeys, opts = [INVALID] * (2 * nchunks), [grp.zero()] * (2 * nchunks)
    for t in range(nchunks):
        cur, acc, nruns = INVALID, grp.zero(), 0
        for pos in range(t * L, min(n, (t + 1) * L)):
            k = keys[pos]
            if k == INVALID:
                continue
            if k != cur:
                if cur != INVALID:
                    if nruns == 1:           # first run of the chunk -> partial slot 0
                        okeys[2 * t], opts[2 * t] = cur, acc
                    else:                    # interior run: complete, sole owner of its bucket
                        buckets[cur] = grp.add(buckets[cur], acc)
                cur, acc, nruns = k, grp.zero(), nruns + 1
            acc = grp.add(acc, pts[pos])
        if cur != INVALID:                   # last run (or the only run)
            slot = 2 * t if nruns == 1 else 2 * t + 1
            okeys[slot], opts[slot] = cur, acc
    return okeys, opts


def final_level(grp, keys, pts, buckets):
    """single CTA: compact valid slots, Hillis-Steele segmented inclusive scan, tails write buckets."""
    dense = [(k, p) for k, p in zip(keys, pts) if k != INVALID]
    m = len(dense)
    ks = [k for k, _ in dense]
    ps = [p for _, p in dense]
    d = 1
    while d < m:
        nxt = list(ps)
        for i in range(m):
            if i >= d and ks[i - d] == ks[i]:
                nxt[i] = grp.add(ps[i - d], ps[i])
        ps = nxt
        d *= 2
    for i in range(m):
        if i == m - 1 or ks[i + 1] != ks[i]:
            buckets[ks[i]] = grp.add(buckets[ks[i]], ps[i])


def msm(grp, scalars, bases, c, L1=8, LK=4, final_max=16, seg_log=2):
    n = len(scalars)
    W = window_count(c)
    B = 1 << (c - 1)
    # counting sort by key = w * B + |d| - 1
    counts = [0] * (W * B)
    digs = [recode(s, c, W) for s in scalars]
    for i in range(n):
        for w, d in enumerate(digs[i]):
            if d:
                counts[w * B + abs(d) - 1] += 1
    offs, run = [], 0
    for x in counts:
        offs.append(run)
        run += x
    T = run
    cursor = list(offs)
    skeys, svals = [INVALID] * (W * n), [0] * (W * n)
    for i in range(n):
        for w, d in enumerate(digs[i]):
            if d:
                k = w * B + abs(d) - 1
                pos = cursor[k]
                cursor[k] += 1
                skeys[pos] = k
                svals[pos] = i | ((1 << 31) if d < 0 else 0)
    assert all(k == INVALID for k in skeys[T:])
    buckets = [grp.zero()] * (W * B)
    # level 1: gather (signed) bases
    pts = [grp.zero() if k == INVALID else (grp.neg(bases[v & 0x7FFFFFFF]) if v >> 31 else bases[v & 0x7FFFFFFF])
           for k, v in zip(skeys, svals)]
    keys, pts = reduce_level(grp, skeys, pts, L1, buckets)
    while len(keys) > final_max:
        keys, pts = reduce_level(grp, keys, pts, LK, buckets)
    final_level(grp, keys, pts, buckets)
    # segmented bucket reduction: per window, segments of S buckets
    S = 1 << seg_log
    if S > B:
        S = B
    wins = []
    for w in range(W):
        total_w = grp.zero()
        for s in range(B // S):
            runs, acc = grp.zero(), grp.zero()
            for j in reversed(range(S)):
                runs = grp.add(runs, buckets[w * B + s * S + j])
                acc = grp.add(acc, runs)
            # + (s*S) * runs  via double-and-add
            m, t, add = s * S, runs, grp.zero()
            while m:
                if m & 1:
                    add = grp.add(add, t)
                t = grp.dbl(t)
                m >>= 1
            total_w = grp.add(total_w, grp.add(acc, add))
        wins.append(total_w)
    # host fold (Horner over windows)
    acc = grp.zero()
    for w in reversed(range(W)):
        for _ in range(c):
            acc = grp.dbl(acc)
        acc = grp.add(acc, wins[w])
    return acc
