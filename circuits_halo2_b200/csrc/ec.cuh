// BN254 G1 (y^2 = x^3 + 3) group law for the MSM kernels: extended Jacobian "XYZZ" accumulators
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2) and halo2curves-layout affine inputs.
//
// Replaces halo2curves 0.1.0 `bn256::{G1Affine, G1}` as used by halo2 `best_multiexp`
// (reference call sites: zk_prover/src/circuits/utils.rs:75-76,94-102 via ParamsKZG::commit*).
// A group element is unique, so the choice of coordinates is free; results are normalised to
// affine before they cross the C ABI.  Every exceptional case (identity operand, P + P, P + -P)
// is handled so results are exact for *any* input, including adversarial / degenerate bases.
#pragma once
#include "fp.cuh"

namespace sb {

struct affine_t {  // halo2curves G1Affine: 64 B, identity = (0, 0)
    fq_t x, y;
    SB_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

struct xyzz_t {  // identity: zz == 0
    fq_t x, y, zz, zzz;
    SB_HD static xyzz_t identity() {
        xyzz_t r;
        r.x = fq_t::zero(); r.y = fq_t::zero(); r.zz = fq_t::zero(); r.zzz = fq_t::zero();
        return r;
    }
    SB_HD bool is_identity() const { return zz.is_zero(); }
    SB_HD static xyzz_t from_affine(const affine_t &p) {
        xyzz_t r;
        if (p.is_identity()) return identity();
        r.x = p.x; r.y = p.y; r.zz = fq_t::one(); r.zzz = fq_t::one();
        return r;
    }
};

// 2 * (x, y) for an affine point (mdbl-2008-s-1)
SB_HD xyzz_t dbl_affine(const affine_t &p) {
    xyzz_t r;
    fq_t u = dbl(p.y);
    fq_t v = sqr(u);
    fq_t w = mul(u, v);
    fq_t s = mul(p.x, v);
    fq_t xx = sqr(p.x);
    fq_t m = add(dbl(xx), xx);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * P (dbl-2008-s-1, a = 0)
SB_HD xyzz_t dbl(const xyzz_t &p) {
    if (p.is_identity()) return p;
    xyzz_t r;
    fq_t u = dbl(p.y);
    fq_t v = sqr(u);
    fq_t w = mul(u, v);
    fq_t s = mul(p.x, v);
    fq_t xx = sqr(p.x);
    fq_t m = add(dbl(xx), xx);
    r.x = sub(sqr(m), dbl(s));
    r.y = sub(mul(m, sub(s, r.x)), mul(w, p.y));
    r.zz = mul(v, p.zz);
    r.zzz = mul(w, p.zzz);
    return r;
}

// acc += (q.x, neg ? -q.y : q.y)   (madd-2008-s: 8M + 2S on the common path)
SB_HD void madd(xyzz_t &acc, const affine_t &q, bool negate) {
    if (q.is_identity()) return;
    fq_t qy = negate ? neg(q.y) : q.y;
    if (acc.is_identity()) {
        acc.x = q.x; acc.y = qy; acc.zz = fq_t::one(); acc.zzz = fq_t::one();
        return;
    }
    fq_t u2 = mul(q.x, acc.zz);
    fq_t s2 = mul(qy, acc.zzz);
    fq_t p = sub(u2, acc.x);
    fq_t r = sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) {
            affine_t t; t.x = q.x; t.y = qy;
            acc = dbl_affine(t);
        } else {
            acc = xyzz_t::identity();
        }
        return;
    }
    fq_t pp = sqr(p);
    fq_t ppp = mul(p, pp);
    fq_t qq = mul(acc.x, pp);
    fq_t x3 = sub(sub(sqr(r), ppp), dbl(qq));
    fq_t y3 = sub(mul(r, sub(qq, x3)), mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = mul(acc.zz, pp);
    acc.zzz = mul(acc.zzz, ppp);
}

// acc += q   (add-2008-s: 12M + 2S)
SB_HD void add(xyzz_t &acc, const xyzz_t &q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) { acc = q; return; }
    fq_t u1 = mul(acc.x, q.zz);
    fq_t u2 = mul(q.x, acc.zz);
    fq_t s1 = mul(acc.y, q.zzz);
    fq_t s2 = mul(q.y, acc.zzz);
    fq_t p = sub(u2, u1);
    fq_t r = sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = dbl(acc);
        else acc = xyzz_t::identity();
        return;
    }
    fq_t pp = sqr(p);
    fq_t ppp = mul(p, pp);
    fq_t qq = mul(u1, pp);
    fq_t x3 = sub(sub(sqr(r), ppp), dbl(qq));
    fq_t y3 = sub(mul(r, sub(qq, x3)), mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = mul(mul(acc.zz, q.zz), pp);
    acc.zzz = mul(mul(acc.zzz, q.zzz), ppp);
}

SB_HD xyzz_t neg(const xyzz_t &p) {
    xyzz_t r = p;
    r.y = neg(p.y);
    return r;
}

// XYZZ -> affine (one field inversion; off the hot path)
SB_HD affine_t to_affine(const xyzz_t &p) {
    affine_t r;
    if (p.is_identity()) { r.x = fq_t::zero(); r.y = fq_t::zero(); return r; }
    // 1/ZZZ; x = X * ZZZ^-2 * ZZ^2 ... use: 1/ZZ = ZZ^2 / ZZZ^2 * ... simpler: invert both via one inversion
    fq_t zi = inv(mul(p.zz, p.zzz));      // 1 / (ZZ * ZZZ)
    fq_t zz_inv = mul(zi, p.zzz);         // 1 / ZZ
    fq_t zzz_inv = mul(zi, p.zz);         // 1 / ZZZ
    r.x = mul(p.x, zz_inv);
    r.y = mul(p.y, zzz_inv);
    return r;
}

}  // namespace sb
