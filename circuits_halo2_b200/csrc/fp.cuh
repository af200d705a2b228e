// 256-bit Montgomery field arithmetic for BN254 Fr / Fq on 8 x 32-bit limbs (sm_100a).
//
// Replaces halo2curves 0.1.0 `bn256::{Fr,Fq}` (reference use: zk_prover/src/circuits/utils.rs:10-13,
// every `Fr as Fp` import).  Same memory layout: 32 B little-endian, Montgomery form, R = 2^256, so a
// Rust `&[Fr]` can be handed over untouched.
//
// Multiplication is an operand-scanning CIOS whose partial products are kept in two interleaved
// accumulators ("even"/"odd" limb alignment) so that every 32x32->64 product is one
// mad.lo.cc/madc.hi.cc pair = one IMAD.WIDE with carry-in/out on the FMA pipe and no carry ripple
// between products.  The carry primitives below have a bit-exact host emulation so the very same
// source is unit-tested on the CPU (tests/test_host_field.py) before it ever reaches a GPU.
#pragma once
#ifndef __CUDACC_RTC__  // the NVRTC-compiled evaluate_h kernels (expr_jit.cu) embed this header and bring their own typedefs
#include <stdint.h>
#endif

#ifdef __CUDACC__
#define SB_HD __host__ __device__ __forceinline__
#else
#define SB_HD inline
#endif

namespace sb {

// ------------------------------------------------------------------ carry-chain primitives
namespace ptx {
#ifdef __CUDA_ARCH__
#define SB_ASM2(name, ins)                                                                                  \
    __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b) {                                      \
        uint32_t r;                                                                                         \
        asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                                        \
        return r;                                                                                           \
    }
#define SB_ASM3(name, ins)                                                                                  \
    __device__ __forceinline__ uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                          \
        uint32_t r;                                                                                         \
        asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));                            \
        return r;                                                                                           \
    }
SB_ASM2(add_cc, "add.cc.u32")
SB_ASM2(addc_cc, "addc.cc.u32")
SB_ASM2(addc, "addc.u32")
SB_ASM2(sub_cc, "sub.cc.u32")
SB_ASM2(subc_cc, "subc.cc.u32")
SB_ASM2(subc, "subc.u32")
SB_ASM3(mad_lo_cc, "mad.lo.cc.u32")
SB_ASM3(madc_lo_cc, "madc.lo.cc.u32")
SB_ASM3(mad_hi_cc, "mad.hi.cc.u32")
SB_ASM3(madc_hi_cc, "madc.hi.cc.u32")
SB_ASM3(madc_hi, "madc.hi.u32")
SB_ASM3(madc_lo, "madc.lo.u32")
#undef SB_ASM2
#undef SB_ASM3
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
// Host emulation of the PTX condition-code register (one carry/borrow bit).
static thread_local uint32_t CF = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; CF = (uint32_t)(d >> 32) & 1; return (uint32_t)d; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - CF; CF = (uint32_t)(d >> 32) & 1; return (uint32_t)d; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c + CF; CF = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + CF; }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return mul_lo(a, b) + c + CF; }
#endif
}  // namespace ptx

// ------------------------------------------------------------------ field parameters
// Limb tables are constexpr *functions* so that, after unrolling, every modulus limb is an
// immediate operand of its IMAD (no constant-bank or register traffic).
struct FrParams {  // scalar field r (InclusionVerifier.sol:210)
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
    SB_HD static constexpr uint32_t mod(int i) {
        return i == 0 ? 0xf0000001u : i == 1 ? 0x43e1f593u : i == 2 ? 0x79b97091u : i == 3 ? 0x2833e848u
             : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    SB_HD static constexpr uint32_t one(int i) {  // R mod r
        return i == 0 ? 0x4ffffffbu : i == 1 ? 0xac96341cu : i == 2 ? 0x9f60cd29u : i == 3 ? 0x36fc7695u
             : i == 4 ? 0x7879462eu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    SB_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
        return i == 0 ? 0xae216da7u : i == 1 ? 0x1bb8e645u : i == 2 ? 0xe35c59e3u : i == 3 ? 0x53fe3ab1u
             : i == 4 ? 0x53bb8085u : i == 5 ? 0x8c49833du : i == 6 ? 0x7f4e44a5u : 0x0216d0b1u;
    }
    SB_HD static constexpr uint32_t r3(int i) {  // R^3 mod r
        return i == 0 ? 0xb4bf0040u : i == 1 ? 0x5e94d8e1u : i == 2 ? 0x1cfbb6b8u : i == 3 ? 0x2a489cbeu
             : i == 4 ? 0xa19fcfedu : i == 5 ? 0x893cc664u : i == 6 ? 0x7fcc657cu : 0x0cf8594bu;
    }
};
struct FqParams {  // base field q (InclusionVerifier.sol:209)
    static constexpr uint32_t INV = 0xe4866389u;
    SB_HD static constexpr uint32_t mod(int i) {
        return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u
             : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    SB_HD static constexpr uint32_t one(int i) {
        return i == 0 ? 0xc58f0d9du : i == 1 ? 0xd35d438du : i == 2 ? 0xf5c70b3du : i == 3 ? 0x0a78eb28u
             : i == 4 ? 0x7879462cu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    SB_HD static constexpr uint32_t r2(int i) {
        return i == 0 ? 0x538afa89u : i == 1 ? 0xf32cfc5bu : i == 2 ? 0xd44501fbu : i == 3 ? 0xb5e71911u
             : i == 4 ? 0x0a417ff6u : i == 5 ? 0x47ab1effu : i == 6 ? 0xcab8351fu : 0x06d89f71u;
    }
    SB_HD static constexpr uint32_t r3(int i) {
        return i == 0 ? 0xda1530dfu : i == 1 ? 0xb1cd6dafu : i == 2 ? 0xa7283db6u : i == 3 ? 0x62f210e6u
             : i == 4 ? 0x0ada0afbu : i == 5 ? 0xef7f0b0cu : i == 6 ? 0x2d592544u : 0x20fd6e90u;
    }
};

// ------------------------------------------------------------------ field element
template <class P>
struct Fp {
    uint32_t v[8];

    SB_HD static Fp zero() { Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = 0;
        return r; }
    SB_HD static Fp one() { Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
        return r; }
    SB_HD static Fp r2() { Fp r;
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = P::r2(i);
        return r; }
    SB_HD bool is_zero() const { return (v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) == 0; }
    SB_HD bool operator==(const Fp &o) const {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= v[i] ^ o.v[i];
        return d == 0;
    }
    SB_HD bool operator!=(const Fp &o) const { return !(*this == o); }
};

// r = (t >= p) ? t - p : t      (t < 2p)
template <class P>
SB_HD void final_sub(uint32_t r[8], const uint32_t t[8]) {
    uint32_t s[8];
    s[0] = ptx::sub_cc(t[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = ptx::subc_cc(t[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);  // 0 or 0xffffffff
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? t[i] : s[i];
}

template <class P>
SB_HD Fp<P> add(const Fp<P> &a, const Fp<P> &b) {
    uint32_t t[8];
    t[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = ptx::addc_cc(a.v[i], b.v[i]);
    t[7] = ptx::addc(a.v[7], b.v[7]);  // p < 2^254: no carry out of limb 7
    Fp<P> r;
    final_sub<P>(r.v, t);
    return r;
}

template <class P>
SB_HD Fp<P> sub(const Fp<P> &a, const Fp<P> &b) {
    uint32_t t[8];
    t[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
    uint32_t borrow = ptx::subc(0, 0);  // all-ones iff a < b
    Fp<P> r;
    r.v[0] = ptx::add_cc(t[0], P::mod(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = ptx::addc_cc(t[i], P::mod(i) & borrow);
    r.v[7] = ptx::addc(t[7], P::mod(7) & borrow);
    return r;
}

template <class P>
SB_HD Fp<P> neg(const Fp<P> &a) { return sub(Fp<P>::zero(), a); }
template <class P>
SB_HD Fp<P> dbl(const Fp<P> &a) { return add(a, a); }

// One CIOS round on the interleaved accumulators (see header comment).
//   E: limbs aligned to their index (new "even" accumulator = previous "odd")
//   O: previous "even" accumulator, E-aligned and with O[0] == 0; it is shifted down two limbs in
//      place and becomes the new "odd" accumulator (limb j sits at position j + 1).
template <class P, bool FIRST>
SB_HD void cios_round(uint32_t E[8], uint32_t O[8], const uint32_t a[8], uint32_t bi) {
    using namespace ptx;
    if (FIRST) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            E[j] = mul_lo(a[j], bi);
            E[j + 1] = mul_hi(a[j], bi);
            O[j] = mul_lo(a[j + 1], bi);
            O[j + 1] = mul_hi(a[j + 1], bi);
        }
    } else {
        E[0] = add_cc(E[0], O[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2]);
            O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 3]);
        }
        O[6] = madc_lo_cc(a[7], bi, 0);
        O[7] = madc_hi(a[7], bi, 0);
        E[0] = mad_lo_cc(a[0], bi, E[0]);
        E[1] = madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            E[j] = madc_lo_cc(a[j], bi, E[j]);
            E[j + 1] = madc_hi_cc(a[j], bi, E[j + 1]);
        }
        O[7] = addc(O[7], 0);
    }
    uint32_t m = mul_lo(E[0], P::INV);
    O[0] = mad_lo_cc(P::mod(1), m, O[0]);
    O[1] = madc_hi_cc(P::mod(1), m, O[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        O[j] = madc_lo_cc(P::mod(j + 1), m, O[j]);
        O[j + 1] = madc_hi_cc(P::mod(j + 1), m, O[j + 1]);
    }
    E[0] = mad_lo_cc(P::mod(0), m, E[0]);
    E[1] = madc_hi_cc(P::mod(0), m, E[1]);
#pragma unroll
    for (int j = 2; j < 8; j += 2) {
        E[j] = madc_lo_cc(P::mod(j), m, E[j]);
        E[j + 1] = madc_hi_cc(P::mod(j), m, E[j + 1]);
    }
    O[7] = addc(O[7], 0);
}

// Montgomery product a * b * 2^-256 mod p, fully reduced.
template <class P>
SB_HD Fp<P> mul(const Fp<P> &a, const Fp<P> &b) {
    using namespace ptx;
    uint32_t X[8], Y[8];
    cios_round<P, true>(X, Y, a.v, b.v[0]);
    cios_round<P, false>(Y, X, a.v, b.v[1]);
    cios_round<P, false>(X, Y, a.v, b.v[2]);
    cios_round<P, false>(Y, X, a.v, b.v[3]);
    cios_round<P, false>(X, Y, a.v, b.v[4]);
    cios_round<P, false>(Y, X, a.v, b.v[5]);
    cios_round<P, false>(X, Y, a.v, b.v[6]);
    cios_round<P, false>(Y, X, a.v, b.v[7]);
    // last round: "even" = Y (Y[0] == 0 after reduction), "odd" = X.  T / 2^32:
    uint32_t t[8];
    t[0] = add_cc(Y[1], X[0]);
#pragma unroll
    for (int j = 1; j < 7; j++) t[j] = addc_cc(Y[j + 1], X[j]);
    t[7] = addc(X[7], 0);
    Fp<P> r;
    final_sub<P>(r.v, t);
    return r;
}

// Montgomery square a * a * 2^-256 mod p, fully reduced: 100 wide multiply-adds instead of the product's 128 (IMAD.WIDE is 4.6 issue cycles per warp
// on sm_100a, everything else here is 2: profiles/r02q_imad_pipes.json).  Separated operand scanning: (A) the 28 products a_i a_j, i < j, row by
// row into an even-aligned and an odd-aligned 512-bit accumulator -- a row is one carry chain per alignment and its carry-out lands in a limb the
// accumulator has not reached yet; (B) S = even + odd, doubled, plus the 8 squares a_i^2 as one more chain; (C) eight reduction rounds m = T_i *
// (-p^-1), T += m p 2^(32 i), again one chain per alignment, whose carry-outs all land in limbs >= 8 -- never read by a later m -- and are
// therefore only counted and added once at the end.
template <class P>
SB_HD Fp<P> sqr(const Fp<P> &x) {
    using namespace ptx;
    const uint32_t *a = x.v;
    uint32_t te[16], to[16];
#pragma unroll
    for (int k = 0; k < 16; k++) te[k] = to[k] = 0;
    // (A) upper triangle
#pragma unroll
    for (int i = 0; i < 7; i++) {
        // odd distance j - i: limb i + j is odd-aligned to the row start
        {
            bool first = true;
            int top = 0;
#pragma unroll
            for (int j = i + 1; j < 8; j += 2) {
                const int q = i + j;
                to[q] = first ? mad_lo_cc(a[i], a[j], to[q]) : madc_lo_cc(a[i], a[j], to[q]);
                to[q + 1] = madc_hi_cc(a[i], a[j], to[q + 1]);
                first = false;
                top = q + 2;
            }
            if (top < 16) to[top] = addc(to[top], 0);
        }
        if (i + 2 < 8) {
            bool first = true;
            int top = 0;
#pragma unroll
            for (int j = i + 2; j < 8; j += 2) {
                const int q = i + j;
                te[q] = first ? mad_lo_cc(a[i], a[j], te[q]) : madc_lo_cc(a[i], a[j], te[q]);
                te[q + 1] = madc_hi_cc(a[i], a[j], te[q + 1]);
                first = false;
                top = q + 2;
            }
            if (top < 16) te[top] = addc(te[top], 0);
        }
    }
    // (B) T = 2 (te + to) + sum_i a_i^2 2^(64 i)
    uint32_t t[16];
    t[0] = 0;
    t[1] = to[1];
    t[2] = add_cc(te[2], to[2]);
#pragma unroll
    for (int k = 3; k < 15; k++) t[k] = addc_cc(te[k], to[k]);
    t[15] = addc(te[15], to[15]);
#pragma unroll
    for (int k = 15; k >= 1; k--) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
    t[0] = mad_lo_cc(a[0], a[0], 0);
    t[1] = madc_hi_cc(a[0], a[0], t[1]);
#pragma unroll
    for (int i = 1; i < 7; i++) {
        t[2 * i] = madc_lo_cc(a[i], a[i], t[2 * i]);
        t[2 * i + 1] = madc_hi_cc(a[i], a[i], t[2 * i + 1]);
    }
    t[14] = madc_lo_cc(a[7], a[7], t[14]);
    t[15] = madc_hi(a[7], a[7], t[15]);
    // (C) Montgomery reduction
    uint32_t cl[16];
#pragma unroll
    for (int k = 8; k < 16; k++) cl[k] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t m = mul_lo(t[i], P::INV);
        t[i] = mad_lo_cc(m, P::mod(0), t[i]);
        t[i + 1] = madc_hi_cc(m, P::mod(0), t[i + 1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            t[i + j] = madc_lo_cc(m, P::mod(j), t[i + j]);
            t[i + j + 1] = madc_hi_cc(m, P::mod(j), t[i + j + 1]);
        }
        cl[i + 8] = addc(cl[i + 8], 0);
        t[i + 1] = mad_lo_cc(m, P::mod(1), t[i + 1]);
        t[i + 2] = madc_hi_cc(m, P::mod(1), t[i + 2]);
#pragma unroll
        for (int j = 3; j < 8; j += 2) {
            t[i + j] = madc_lo_cc(m, P::mod(j), t[i + j]);
            t[i + j + 1] = madc_hi_cc(m, P::mod(j), t[i + j + 1]);
        }
        if (i + 9 < 16) cl[i + 9] = addc(cl[i + 9], 0);  // i = 7: a carry out of limb 15 cannot happen (a^2 + m p < 2^512)
    }
    uint32_t r[8];
    r[0] = add_cc(t[8], cl[8]);
#pragma unroll
    for (int k = 1; k < 7; k++) r[k] = addc_cc(t[8 + k], cl[8 + k]);
    r[7] = addc(t[15], cl[15]);
    Fp<P> out;
    final_sub<P>(out.v, r);
    return out;
}

template <class P>
SB_HD Fp<P> to_mont(const Fp<P> &a) { return mul(a, Fp<P>::r2()); }
template <class P>
SB_HD Fp<P> from_mont(const Fp<P> &a) {
    Fp<P> o = Fp<P>::zero();
    o.v[0] = 1;
    return mul(a, o);
}

// a^(p-2) (Fermat): 254 dependent squarings, ~0.26 ms of latency for one thread on a B200 (profiles/r02n).  Kept as the cross-check of inv().
template <class P>
SB_HD Fp<P> inv_fermat(const Fp<P> &a) {
    Fp<P> acc = Fp<P>::one();
    for (int i = 255; i >= 0; i--) {
        acc = sqr(acc);
        // exponent p - 2: only limb 0 differs from p (p is odd and p[0] >= 2)
        uint32_t w = P::mod(i >> 5);
        if ((i >> 5) == 0) w -= 2;
        if ((w >> (i & 31)) & 1) acc = mul(acc, a);
    }
    return acc;
}

// a^-1 (0 for a = 0) by the binary extended Euclid: at most ~2 * 254 rounds of a 256-bit compare / subtract / shift and no product, against the
// 380 dependent products of Fermat's exponentiation -- the batch inversions and normalisations that call this are single latency chains.
// Invariants for the integer A = a.v (the Montgomery representation): x1 * A = u and x2 * A = v (mod p), v odd; the loop ends with u = 0,
// v = gcd = 1 and x2 = A^-1 = a^-1 R^-1, which one product with R^3 turns into the Montgomery form a^-1 R.
template <class P>
SB_HD Fp<P> inv(const Fp<P> &a) {
    using namespace ptx;
    uint32_t u[8], v[8];
    Fp<P> x1 = Fp<P>::zero(), x2 = Fp<P>::zero();
    x1.v[0] = 1;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u[i] = a.v[i];
        v[i] = P::mod(i);
    }
    while ((u[0] | u[1] | u[2] | u[3] | u[4] | u[5] | u[6] | u[7]) != 0) {
        if (u[0] & 1) {
            uint32_t d[8];
            d[0] = sub_cc(u[0], v[0]);
#pragma unroll
            for (int i = 1; i < 8; i++) d[i] = subc_cc(u[i], v[i]);
            const uint32_t lt = subc(0, 0);  // all-ones iff u < v
            if (lt) {                        // (u, v, x1, x2) <- (v - u, u, x2 - x1, x1)
                const Fp<P> t = sub(x2, x1);
                x2 = x1;
                x1 = t;
                uint32_t e[8];
                e[0] = sub_cc(v[0], u[0]);
#pragma unroll
                for (int i = 1; i < 8; i++) e[i] = subc_cc(v[i], u[i]);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    v[i] = u[i];
                    u[i] = e[i];
                }
            } else {
                x1 = sub(x1, x2);
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = d[i];
            }
        }
        // u is even: halve it, and x1 with it (x1 / 2 mod p = (x1 + p) / 2 for odd x1; x1 + p < 2^255)
#pragma unroll
        for (int i = 0; i < 7; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
        u[7] >>= 1;
        const uint32_t odd = 0u - (x1.v[0] & 1u);
        uint32_t h[8];
        h[0] = add_cc(x1.v[0], P::mod(0) & odd);
#pragma unroll
        for (int i = 1; i < 7; i++) h[i] = addc_cc(x1.v[i], P::mod(i) & odd);
        h[7] = addc(x1.v[7], P::mod(7) & odd);
#pragma unroll
        for (int i = 0; i < 7; i++) x1.v[i] = (h[i] >> 1) | (h[i + 1] << 31);
        x1.v[7] = h[7] >> 1;
    }
    Fp<P> r3;
#pragma unroll
    for (int i = 0; i < 8; i++) r3.v[i] = P::r3(i);
    return mul(x2, r3);
}

typedef Fp<FrParams> fr_t;
typedef Fp<FqParams> fq_t;

// ------------------------------------------------------------------ 128-bit vector global access
#ifdef __CUDACC__
template <class P>
__device__ __forceinline__ Fp<P> load_fp(const void *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    Fp<P> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <class P>
__device__ __forceinline__ Fp<P> ldg_fp(const void *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    Fp<P> r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
template <class P>
__device__ __forceinline__ void store_fp(void *p, const Fp<P> &a) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}
#endif

}  // namespace sb
