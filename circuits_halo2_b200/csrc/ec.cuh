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

// Field products of the group law.  With SB_EC_NOINLINE_MUL the Montgomery product is ONE shared subroutine (operands and result
// by value, i.e. in registers under the device ABI) instead of ten inlined copies per addition: the fully inlined madd body is
// ~40 KB of SASS and ncu shows `no_instruction` (instruction-cache miss) as the top stall of the MSM accumulation kernel.
#if defined(SB_EC_NOINLINE_MUL) && defined(__CUDA_ARCH__)
static __device__ __noinline__ fq_t ec_mul(fq_t a, fq_t b) { return mul(a, b); }
#else
SB_HD fq_t ec_mul(const fq_t &a, const fq_t &b) { return mul(a, b); }
#endif
// Squares: the dedicated squaring (fp.cuh sqr, 91 + 9 wide multiply-adds against 120 + 8) is 14 % cheaper on the multiplier pipe in isolation, but as a
// second shared subroutine of the MSM kernels it bought nothing measurable (level 1 of a k = 20 proof: 30.7 ms with it, 30.6 without; 2^22 MSM
// 361.8 vs 361.7 Mpts/s, profiles/r02q) -- its extra carry bookkeeping lands on the same pipe.  Opt in with -DSB_EC_SQR.
#if defined(SB_EC_SQR) && defined(SB_EC_NOINLINE_MUL) && defined(__CUDA_ARCH__)
static __device__ __noinline__ fq_t ec_sqr(fq_t a) { return sqr(a); }
#elif defined(SB_EC_SQR)
SB_HD fq_t ec_sqr(const fq_t &a) { return sqr(a); }
#else
SB_HD fq_t ec_sqr(const fq_t &a) { return ec_mul(a, a); }
#endif

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
    fq_t v = ec_sqr(u);
    fq_t w = ec_mul(u, v);
    fq_t s = ec_mul(p.x, v);
    fq_t xx = ec_sqr(p.x);
    fq_t m = add(dbl(xx), xx);
    r.x = sub(ec_sqr(m), dbl(s));
    r.y = sub(ec_mul(m, sub(s, r.x)), ec_mul(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * P (dbl-2008-s-1, a = 0)
SB_HD xyzz_t dbl(const xyzz_t &p) {
    if (p.is_identity()) return p;
    xyzz_t r;
    fq_t u = dbl(p.y);
    fq_t v = ec_sqr(u);
    fq_t w = ec_mul(u, v);
    fq_t s = ec_mul(p.x, v);
    fq_t xx = ec_sqr(p.x);
    fq_t m = add(dbl(xx), xx);
    r.x = sub(ec_sqr(m), dbl(s));
    r.y = sub(ec_mul(m, sub(s, r.x)), ec_mul(w, p.y));
    r.zz = ec_mul(v, p.zz);
    r.zzz = ec_mul(w, p.zzz);
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
    fq_t u2 = ec_mul(q.x, acc.zz);
    fq_t s2 = ec_mul(qy, acc.zzz);
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
    fq_t pp = ec_sqr(p);
    fq_t ppp = ec_mul(p, pp);
    fq_t qq = ec_mul(acc.x, pp);
    fq_t x3 = sub(sub(ec_sqr(r), ppp), dbl(qq));
    fq_t y3 = sub(ec_mul(r, sub(qq, x3)), ec_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = ec_mul(acc.zz, pp);
    acc.zzz = ec_mul(acc.zzz, ppp);
}

// acc += q   (add-2008-s: 12M + 2S)
SB_HD void add(xyzz_t &acc, const xyzz_t &q) {
    if (q.is_identity()) return;
    if (acc.is_identity()) { acc = q; return; }
    fq_t u1 = ec_mul(acc.x, q.zz);
    fq_t u2 = ec_mul(q.x, acc.zz);
    fq_t s1 = ec_mul(acc.y, q.zzz);
    fq_t s2 = ec_mul(q.y, acc.zzz);
    fq_t p = sub(u2, u1);
    fq_t r = sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) acc = dbl(acc);
        else acc = xyzz_t::identity();
        return;
    }
    fq_t pp = ec_sqr(p);
    fq_t ppp = ec_mul(p, pp);
    fq_t qq = ec_mul(u1, pp);
    fq_t x3 = sub(sub(ec_sqr(r), ppp), dbl(qq));
    fq_t y3 = sub(ec_mul(r, sub(qq, x3)), ec_mul(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = ec_mul(ec_mul(acc.zz, q.zz), pp);
    acc.zzz = ec_mul(ec_mul(acc.zzz, q.zzz), ppp);
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
    fq_t zi = inv(ec_mul(p.zz, p.zzz));      // 1 / (ZZ * ZZZ)
    fq_t zz_inv = ec_mul(zi, p.zzz);         // 1 / ZZ
    fq_t zzz_inv = ec_mul(zi, p.zz);         // 1 / ZZZ
    r.x = ec_mul(p.x, zz_inv);
    r.y = ec_mul(p.y, zzz_inv);
    return r;
}

}  // namespace sb
