// Host-side tail of the MSM: Horner fold of the <= 64 per-window sums the GPU returns
// (sum_w 2^(c w) W_w) and normalisation to an affine point.  254 dependent doublings are a
// latency chain that one CPU core finishes in ~0.1 ms; a GPU thread would need ~10x longer, and
// BASELINE.json's north_star prescribes "partial G1 sums are reduced on the host".
// Own 4 x 64-bit Montgomery arithmetic (unsigned __int128); not shared with oracle/.
#include <stdint.h>
#include <string.h>

namespace sb {
namespace {

typedef unsigned __int128 u128;
typedef uint64_t u64;

const u64 QM[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
const u64 QINV = 0x87d20782e4866389ULL;  // -q^-1 mod 2^64
const u64 QONE[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};

struct fe { u64 v[4]; };

inline bool is_zero(const fe &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
inline bool eq(const fe &a, const fe &b) { return a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2] && a.v[3] == b.v[3]; }
inline bool geq_q(const u64 t[4]) {
    for (int i = 3; i >= 0; i--) {
        if (t[i] > QM[i]) return true;
        if (t[i] < QM[i]) return false;
    }
    return true;
}
inline void sub_q(u64 t[4]) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)t[i] - QM[i] - br;
        t[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
}
inline fe add(const fe &a, const fe &b) {
    fe r;
    u64 c = 0;
    for (int i = 0; i < 4; i++) {
        u128 s = (u128)a.v[i] + b.v[i] + c;
        r.v[i] = (u64)s;
        c = (u64)(s >> 64);
    }
    if (c || geq_q(r.v)) sub_q(r.v);
    return r;
}
inline fe sub(const fe &a, const fe &b) {
    fe r;
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - br;
        r.v[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
    if (br) {
        u64 c = 0;
        for (int i = 0; i < 4; i++) {
            u128 s = (u128)r.v[i] + QM[i] + c;
            r.v[i] = (u64)s;
            c = (u64)(s >> 64);
        }
    }
    return r;
}
inline fe mul(const fe &a, const fe &b) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u64 carry = 0;
        u128 acc;
        for (int j = 0; j < 4; j++) {
            acc = (u128)a.v[j] * b.v[i] + t[j] + carry;
            t[j] = (u64)acc;
            carry = (u64)(acc >> 64);
        }
        acc = (u128)t[4] + carry;
        t[4] = (u64)acc;
        t[5] = (u64)(acc >> 64);
        u64 m = t[0] * QINV;
        acc = (u128)m * QM[0] + t[0];
        carry = (u64)(acc >> 64);
        for (int j = 1; j < 4; j++) {
            acc = (u128)m * QM[j] + t[j] + carry;
            t[j - 1] = (u64)acc;
            carry = (u64)(acc >> 64);
        }
        acc = (u128)t[4] + carry;
        t[3] = (u64)acc;
        t[4] = t[5] + (u64)(acc >> 64);
    }
    fe r;
    memcpy(r.v, t, 32);
    if (t[4] || geq_q(r.v)) sub_q(r.v);
    return r;
}
inline fe sqr(const fe &a) { return mul(a, a); }
inline fe dbl(const fe &a) { return add(a, a); }
// R^3 mod q: turns the plain inverse of a Montgomery representation (a R)^-1 into the Montgomery form a^-1 R with one product
const u64 QR3[4] = {0xb1cd6dafda1530dfULL, 0x62f210e6a7283db6ULL, 0xef7f0b0c0ada0afbULL, 0x20fd6e902d592544ULL};
fe inv(const fe &a) {  // binary extended Euclid (as csrc/fp.cuh inv): ~500 rounds of shift / subtract instead of 380 products
    u64 u[4], v[4];
    memcpy(u, a.v, 32);
    memcpy(v, QM, 32);
    fe x1, x2;
    memset(&x1, 0, sizeof x1);
    memset(&x2, 0, sizeof x2);
    x1.v[0] = 1;
    auto lt = [](const u64 a4[4], const u64 b4[4]) {
        for (int i = 3; i >= 0; i--) {
            if (a4[i] < b4[i]) return true;
            if (a4[i] > b4[i]) return false;
        }
        return false;
    };
    auto sub4 = [](u64 r[4], const u64 a4[4], const u64 b4[4]) {
        u64 br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)a4[i] - b4[i] - br;
            r[i] = (u64)d;
            br = (u64)(d >> 64) & 1;
        }
    };
    while (u[0] | u[1] | u[2] | u[3]) {
        if (u[0] & 1) {
            if (lt(u, v)) {
                fe t = sub(x2, x1);
                x2 = x1;
                x1 = t;
                u64 e[4];
                sub4(e, v, u);
                memcpy(v, u, 32);
                memcpy(u, e, 32);
            } else {
                x1 = sub(x1, x2);
                sub4(u, u, v);
            }
        }
        for (int i = 0; i < 3; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 63);
        u[3] >>= 1;
        u64 h[4], c = 0;
        const u64 odd = 0 - (x1.v[0] & 1);
        for (int i = 0; i < 4; i++) {
            u128 sum = (u128)x1.v[i] + (QM[i] & odd) + c;
            h[i] = (u64)sum;
            c = (u64)(sum >> 64);
        }
        for (int i = 0; i < 3; i++) x1.v[i] = (h[i] >> 1) | (h[i + 1] << 63);
        x1.v[3] = h[3] >> 1;
    }
    fe r3;
    memcpy(r3.v, QR3, 32);
    return mul(x2, r3);
}

struct pt { fe x, y, zz, zzz; };  // XYZZ, identity zz == 0

pt pdbl(const pt &p) {
    if (is_zero(p.zz)) return p;
    pt r;
    fe u = dbl(p.y), v = sqr(u), w = mul(u, v), s = mul(p.x, v), xx = sqr(p.x);
    fe m = add(dbl(xx), xx);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = mul(v, p.zz);
    r.zzz = mul(w, p.zzz);
    return r;
}
pt padd(const pt &a, const pt &b) {
    if (is_zero(b.zz)) return a;
    if (is_zero(a.zz)) return b;
    fe u1 = mul(a.x, b.zz), u2 = mul(b.x, a.zz), s1 = mul(a.y, b.zzz), s2 = mul(b.y, a.zzz);
    fe p = sub(u2, u1), r = sub(s2, s1);
    if (is_zero(p)) {
        if (is_zero(r)) return pdbl(a);
        pt id;
        memset(&id, 0, sizeof id);
        return id;
    }
    fe pp = sqr(p), ppp = mul(p, pp), q = mul(u1, pp);
    pt o;
    o.x = sub(sub(sqr(r), ppp), dbl(q));
    o.y = sub(mul(r, sub(q, o.x)), mul(s1, ppp));
    o.zz = mul(mul(a.zz, b.zz), pp);
    o.zzz = mul(mul(a.zzz, b.zzz), ppp);
    return o;
}

}  // namespace

// win: n_windows XYZZ points (4 x 32 B each, Montgomery, GPU limb layout == u64 LE layout).
// out_affine = sum_w 2^(c w) win[w], 64 B halo2curves G1Affine (identity = zeros).
void host_fold_windows(const uint8_t *win, int n_windows, int c, uint8_t out_affine[64]) {
    pt acc;
    memset(&acc, 0, sizeof acc);
    for (int w = n_windows - 1; w >= 0; w--) {
        for (int i = 0; i < c; i++) acc = pdbl(acc);
        pt q;
        memcpy(&q, win + (size_t)w * 128, 128);
        acc = padd(acc, q);
    }
    if (is_zero(acc.zz)) {
        memset(out_affine, 0, 64);
        return;
    }
    fe zi = inv(mul(acc.zz, acc.zzz));
    fe x = mul(acc.x, mul(zi, acc.zzz));
    fe y = mul(acc.y, mul(zi, acc.zz));
    memcpy(out_affine, x.v, 32);
    memcpy(out_affine + 32, y.v, 32);
}

// Tail of the tree bucket reduction (msm.cu, msm_bucket_tree_kernel): fin = records of 128 B XYZZ points, [0] = T, [1 + b] = S_b;
// out = fin[x_slot] + 2^shift * sum_{b < n_bits} 2^b S_b  (Horner over the bits, then `shift` more doublings), as an XYZZ point.
void host_bucket_combine(const uint8_t *fin, int n_bits, int shift, int x_slot, uint8_t out_xyzz[128]) {
    pt acc;
    memset(&acc, 0, sizeof acc);
    for (int b = n_bits - 1; b >= 0; b--) {
        acc = pdbl(acc);
        pt q;
        memcpy(&q, fin + (size_t)(1 + b) * 128, 128);
        acc = padd(acc, q);
    }
    for (int i = 0; i < shift; i++) acc = pdbl(acc);
    pt x;
    memcpy(&x, fin + (size_t)x_slot * 128, 128);
    acc = padd(acc, x);
    memcpy(out_xyzz, &acc, 128);
}

// Residue shard of a table MSM (msm.cu): the rank kept the digits with (|d| - 1) mod 2^log_mod == res in bucket b' = (|d| - 1) >> log_mod, whose true
// weight is 2^log_mod b' + res + 1.  r = sum (b' + 1) V_b' comes in, total = sum V_b';  r <- 2^log_mod r - (2^log_mod - res - 1) total.
void host_residue_fixup(uint8_t r_xyzz[128], const uint8_t total_xyzz[128], int log_mod, int res) {
    pt r, t;
    memcpy(&r, r_xyzz, 128);
    memcpy(&t, total_xyzz, 128);
    for (int i = 0; i < log_mod; i++) r = pdbl(r);
    const unsigned k = (1u << log_mod) - (unsigned)res - 1u;
    pt m;
    memset(&m, 0, sizeof m);
    for (int b = 31; b >= 0; b--) {   // m = k * total
        m = pdbl(m);
        if ((k >> b) & 1) m = padd(m, t);
    }
    if (!is_zero(m.zz)) {
        fe zero;
        memset(&zero, 0, sizeof zero);
        m.y = sub(zero, m.y);
        r = padd(r, m);
    }
    memcpy(r_xyzz, &r, 128);
}

// Fq Montgomery (32 B) -> canonical little-endian bytes (transcript encodings of commitments)
void host_fq_to_canonical(const uint8_t mont[32], uint8_t canon_le[32]) {
    fe a, one_c;
    memcpy(a.v, mont, 32);
    memset(&one_c, 0, sizeof one_c);
    one_c.v[0] = 1;
    fe c = mul(a, one_c);
    for (int i = 0; i < 32; i++) canon_le[i] = (uint8_t)(c.v[i >> 3] >> (8 * (i & 7)));
}

// T[j] = 2^j * G for j < 254, affine halo2curves layout (G = (1, 2))
void host_pow2_table(uint8_t *out) {
    pt p;
    memcpy(p.x.v, QONE, 32);
    p.y = dbl(p.x);  // 2 in Montgomery form
    memcpy(p.zz.v, QONE, 32);
    memcpy(p.zzz.v, QONE, 32);
    for (int j = 0; j < 254; j++) {
        fe zi = inv(mul(p.zz, p.zzz));
        fe x = mul(p.x, mul(zi, p.zzz));
        fe y = mul(p.y, mul(zi, p.zz));
        memcpy(out + 64 * j, x.v, 32);
        memcpy(out + 64 * j + 32, y.v, 32);
        p = pdbl(p);
    }
}

// sum of n affine points (multi-GPU partial results are combined with this)
void host_sum_affine(const uint8_t *pts, int n, uint8_t out_affine[64]) {
    pt acc;
    memset(&acc, 0, sizeof acc);
    for (int i = 0; i < n; i++) {
        pt q;
        memcpy(&q.x, pts + (size_t)i * 64, 32);
        memcpy(&q.y, pts + (size_t)i * 64 + 32, 32);
        if (is_zero(q.x) && is_zero(q.y)) continue;
        memcpy(q.zz.v, QONE, 32);
        memcpy(q.zzz.v, QONE, 32);
        acc = padd(acc, q);
    }
    if (is_zero(acc.zz)) {
        memset(out_affine, 0, 64);
        return;
    }
    fe zi = inv(mul(acc.zz, acc.zzz));
    fe x = mul(acc.x, mul(zi, acc.zzz));
    fe y = mul(acc.y, mul(zi, acc.zz));
    memcpy(out_affine, x.v, 32);
    memcpy(out_affine + 32, y.v, 32);
}

}  // namespace sb
