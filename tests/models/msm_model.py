"""Executable model of the MSM pipeline in csrc/msm.cu, over an abstract additive group.

Pins, on the CPU, the control logic that is hard to eyeball in CUDA: signed-digit recoding,
counting sort by (window, bucket), the key-less level 1 (bucket of a position from the sort's offsets), the chunked reduce-by-key levels
(first/last runs of a chunk become partials for the next level, interior runs go straight to their bucket),
the final single-CTA segmented scan, the hierarchical bucket reduction, fixed-base window tables with batches of
scalar vectors, window ranges (multi-GPU sharding) and the host window fold.
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


def cta_level(grp, keys, pts, width, buckets):
    """msm_reduce_cta_kernel: one CTA per `width` slots; compaction + segmented scan; runs strictly inside the CTA's valid slots are complete
    and go to their bucket, the runs touching its first / last valid slot become its two partial slots."""
    n = len(keys)
    nct = (n + width - 1) // width
    okeys, opts = [INVALID] * (2 * nct), [grp.zero()] * (2 * nct)
    for b in range(nct):
        dense = [(k, p) for k, p in zip(keys[b * width:(b + 1) * width], pts[b * width:(b + 1) * width]) if k != INVALID]
        m = len(dense)
        i = 0
        while i < m:
            j, acc = i, grp.zero()
            while j < m and dense[j][0] == dense[i][0]:
                acc = grp.add(acc, dense[j][1])
                j += 1
            key = dense[i][0]
            if key == dense[0][0]:
                okeys[2 * b], opts[2 * b] = key, acc
            elif j == m:
                okeys[2 * b + 1], opts[2 * b + 1] = key, acc
            else:
                buckets[key] = grp.add(buckets[key], acc)
            i = j
    return okeys, opts


def bucket_of(offsets, nb, pos):
    """largest b in [0, nb) with offsets[b] <= pos (msm.cu::bucket_of)"""
    lo, hi = 0, nb
    while hi - lo > 1:
        mid = (lo + hi) >> 1
        if offsets[mid] <= pos:
            lo = mid
        else:
            hi = mid
    return lo


def reduce_first(grp, offsets, nb, vals, points, L, buckets):
    """level 1 without a key array (msm_reduce_first_kernel): the bucket of a sorted position follows from the exclusive offsets of
    the counting sort; neighbour check first, binary search when the walk crosses empty buckets."""
    n_valid = offsets[nb]
    nchunks = (max(n_valid, 1) + L - 1) // L
    okeys, opts = [INVALID] * (2 * nchunks), [grp.zero()] * (2 * nchunks)
    for t in range(nchunks):
        start = t * L
        if start >= n_valid:
            continue
        end = min(n_valid, start + L)
        cur = bucket_of(offsets, nb, start)
        next_off = offsets[cur + 1]
        nruns, acc = 1, grp.zero()
        for pos in range(start, end):
            if pos >= next_off:
                if nruns == 1:
                    okeys[2 * t], opts[2 * t] = cur, acc
                else:
                    buckets[cur] = grp.add(buckets[cur], acc)
                cur += 1
                next_off = offsets[cur + 1]
                if pos >= next_off:
                    cur = bucket_of(offsets, nb, pos)
                    next_off = offsets[cur + 1]
                nruns += 1
                acc = grp.zero()
            v = vals[pos]
            p = points[v & 0x7FFFFFFF]
            acc = grp.add(acc, grp.neg(p) if v >> 31 else p)
        slot = 2 * t if nruns == 1 else 2 * t + 1
        okeys[slot], opts[slot] = cur, acc
    return okeys, opts


def bucket_hierarchy(grp, buckets, n_sets, B, seg_log0, finish_at=8):
    """sum_i (i + 1) * B_i per bucket set without scalar multiplication until few elements are left
    (msm_bucket_level_kernel<0/1>, msm_bucket_finish_kernel): R = sum_i [Q_i + lambda * i * P_i]."""
    out = []
    for w in range(n_sets):
        P = list(buckets[w * B:(w + 1) * B])
        Q = None
        m, lam_log, level = B, 0, 0
        while level < 2 and m > finish_at:
            sl = (seg_log0 if seg_log0 else 1) if level == 0 else 3
            while (1 << sl) > m:
                sl -= 1
            S = 1 << sl
            nP, nQ = [], []
            for sg in range(m >> sl):
                run, acc = grp.zero(), grp.zero()
                for j in range(S - 1, 0, -1):
                    run = grp.add(run, P[sg * S + j])
                    acc = grp.add(acc, run)
                run = grp.add(run, P[sg * S])
                for _ in range(lam_log):
                    acc = grp.dbl(acc)
                if Q is None:
                    acc = grp.add(acc, run)
                else:
                    for j in range(S):
                        acc = grp.add(acc, Q[sg * S + j])
                nP.append(run)
                nQ.append(acc)
            P, Q = nP, nQ
            lam_log += sl
            m >>= sl
            level += 1
        total = grp.zero()
        for sidx in range(m):   # finish: V_s = Q_s + lambda * s * P_s by double-and-add, then a tree (order-free in an abelian group)
            v = grp.zero()
            if sidx:
                for bit in reversed(range(sidx.bit_length())):
                    v = grp.dbl(v)
                    if (sidx >> bit) & 1:
                        v = grp.add(v, P[sidx])
                for _ in range(lam_log):
                    v = grp.dbl(v)
            v = grp.add(v, Q[sidx] if Q is not None else P[sidx])
            total = grp.add(total, v)
        out.append(total)
    return out


def _tree_cta(grp, load, m, node, bt_log, bits):
    """msm_bucket_tree_kernel, one CTA: fold elements [node * BT, (node + 1) * BT) (identity beyond m) into one record
    (T, S_0 .. S_{bt_log - 1}) -- or just T when `bits` is false -- with the kernel's slot layout and ping-pong steps."""
    BT = 1 << bt_log
    recs = []                                   # records of 2 buckets, straight from "global memory"
    for t in range(BT // 2):
        e = node * BT + 2 * t
        a = load(e) if e < m else grp.zero()
        b = load(e + 1) if e + 1 < m else grp.zero()
        recs.append([grp.add(a, b), b] if bits else [grp.add(a, b)])
    for l in range(1, bt_log):
        per_in = l + 1 if bits else 1
        n_out = BT >> (l + 1)
        nxt = []
        for j in range(n_out):
            L, R = recs[2 * j], recs[2 * j + 1]
            assert len(L) == per_in and len(R) == per_in
            rec = [grp.add(L[q], R[q]) for q in range(per_in)]
            if bits:
                rec.append(R[0])                # S_l = T of the upper half
            nxt.append(rec)
        recs = nxt
    assert len(recs) == 1
    return recs[0]


def bucket_tree(grp, buckets, n_sets, B, bt_log=8):
    """sum_i (i + 1) * B_i per set through msm_bucket_tree_kernel + host_bucket_combine: R = X + 2^shift * sum_b 2^b S_b."""
    BT = 1 << bt_log
    FIN = 2 * bt_log + 2
    out = []
    for w in range(n_sets):
        P = list(buckets[w * B:(w + 1) * B])
        Q = None
        m, shift = B, 0
        if m > BT * BT:                         # one running-sum level first (msm_bucket_level_kernel<false>, lambda = 1)
            shift = (m.bit_length() - 1) - 2 * bt_log
            S = 1 << shift
            nP, nQ = [], []
            for sg in range(m >> shift):
                run, acc = grp.zero(), grp.zero()
                for j in range(S - 1, 0, -1):
                    run = grp.add(run, P[sg * S + j])
                    acc = grp.add(acc, run)
                run = grp.add(run, P[sg * S])
                nP.append(run)
                nQ.append(grp.add(acc, run))
            P, Q = nP, nQ
            m >>= shift
        n_bits = m.bit_length() - 1
        n_nodes = (m + BT - 1) // BT
        fin = [None] * FIN
        if n_nodes == 1:
            rec = _tree_cta(grp, lambda e: P[e], m, 0, bt_log, True)
            fin[0] = rec[0]
            for b in range(bt_log):
                fin[1 + b] = rec[1 + b]
        else:
            nodes = [_tree_cta(grp, lambda e: P[e], m, nd, bt_log, True) for nd in range(n_nodes)]
            for x in range(bt_log + 1):         # the colsum launch: CTA x folds slot x of every node record
                rec = _tree_cta(grp, lambda e: nodes[e][x], n_nodes, 0, bt_log, x == 0)
                fin[x] = rec[0]
                if x == 0:
                    for b in range(bt_log):
                        fin[bt_log + 1 + b] = rec[1 + b]
        if Q is not None:
            nodes_q = [_tree_cta(grp, lambda e: Q[e], m, nd, bt_log, False) for nd in range(n_nodes)]
            fin[FIN - 1] = _tree_cta(grp, lambda e: nodes_q[e][0], n_nodes, 0, bt_log, False)[0]
        # host_bucket_combine
        acc = grp.zero()
        for b in reversed(range(n_bits)):
            acc = grp.add(grp.dbl(acc), fin[1 + b])
        for _ in range(shift):
            acc = grp.dbl(acc)
        out.append(grp.add(acc, fin[FIN - 1] if shift else fin[0]))
    return out


def msm(grp, scalars, bases, c, L1=8, LK=4, final_max=16, seg_log=2, tables=False, batch=1, w_lo=0, w_hi=None, cta_scan_max=0, tree_log=None, residue=None):
    """scalars: batch * n values (vector j = scalars[j n:(j+1) n], tables only).  Returns the list of `batch` results (tables), or the
    single result; with a window range [w_lo, w_hi) the partial sum over those windows (tables: already carrying 2^(c w))."""
    n = len(bases)
    assert len(scalars) == n * batch and (batch == 1 or tables)
    W_all = window_count(c)
    w_hi = W_all if w_hi is None else w_hi
    res, log_mod = residue if residue else (0, 0)   # residue shard (tables + tree only): digits with (|d| - 1) mod 2^log_mod == res, bucket (|d| - 1) >> log_mod
    assert not residue or (tables and tree_log and c >= log_mod + 2)
    B = (1 << (c - 1)) >> log_mod
    n_sets = batch if tables else (w_hi - w_lo)
    nb = n_sets * B
    if tables:   # tables[w * n + i] = 2^(c w) * P_i
        points = []
        for w in range(W_all):
            for i in range(n):
                t = bases[i]
                for _ in range(c * w):
                    t = grp.dbl(t)
                points.append(t)
    else:
        points = bases
    digs = [recode(s, c, W_all) for s in scalars]

    def key_val(gi, w, d):
        piece, i = divmod(gi, n)
        dm = abs(d) - 1
        if (dm & ((1 << log_mod) - 1)) != res:
            return None, None
        key = (piece * B if tables else (w - w_lo) * B) + (dm >> log_mod)
        val = (w * n + i if tables else i) | ((1 << 31) if d < 0 else 0)
        return key, val

    counts = [0] * nb
    for gi in range(n * batch):
        for w in range(w_lo, w_hi):
            d = digs[gi][w]
            if d and key_val(gi, w, d)[0] is not None:
                counts[key_val(gi, w, d)[0]] += 1
    offsets, run = [], 0
    for x in counts:
        offsets.append(run)
        run += x
    offsets.append(run)                       # offsets[nb] = number of valid digits
    offsets += [0, 0, 0]
    cursor = list(offsets[:nb])
    svals = [0] * (run + 4)
    for gi in range(n * batch):
        for w in range(w_lo, w_hi):
            d = digs[gi][w]
            if d:
                k, v = key_val(gi, w, d)
                if k is None:
                    continue
                svals[cursor[k]] = v
                cursor[k] += 1
    buckets = [grp.zero()] * nb
    keys, pts = reduce_first(grp, offsets, nb, svals, points, L1, buckets)
    while len(keys) > final_max:
        if len(keys) <= cta_scan_max:
            keys, pts = cta_level(grp, keys, pts, final_max, buckets)
        else:
            keys, pts = reduce_level(grp, keys, pts, LK, buckets)
    final_level(grp, keys, pts, buckets)
    wins = bucket_tree(grp, buckets, n_sets, B, tree_log) if tree_log else bucket_hierarchy(grp, buckets, n_sets, B, seg_log)
    if log_mod:   # host_residue_fixup: true weight of bucket b' is 2^log_mod b' + res + 1
        fixed = []
        for j in range(n_sets):
            total = grp.zero()
            for b in range(B):
                total = grp.add(total, buckets[j * B + b])
            r = wins[j]
            for _ in range(log_mod):
                r = grp.dbl(r)
            k = (1 << log_mod) - res - 1
            m = grp.zero()
            for _ in range(k):
                m = grp.add(m, total)
            fixed.append(grp.add(r, grp.neg(m)))
        wins = fixed
    if tables:
        return wins                           # one (partial) commitment per scalar vector
    if w_hi - w_lo != W_all:
        return wins                           # window sums of the range, folded by the caller
    acc = grp.zero()
    for w in reversed(range(W_all)):
        for _ in range(c):
            acc = grp.dbl(acc)
        acc = grp.add(acc, wins[w])
    return acc
