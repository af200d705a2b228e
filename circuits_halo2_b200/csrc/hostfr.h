// Host-side BN254 scalar-field arithmetic (4 x 64-bit Montgomery limbs, unsigned __int128) for the
// O(1)-per-proof scalar work of create_proof: challenges, rotations of x, SHPLONK interpolation,
// vanishing-polynomial evaluations.  Same memory layout as halo2curves `Fr` and as the device `fr_t`.
// Own code (not shared with oracle/).
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace sb {
namespace hfr {

typedef unsigned __int128 u128;
typedef uint64_t u64;

struct Fr {
    u64 v[4];
    bool operator==(const Fr &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator!=(const Fr &o) const { return !(*this == o); }
};

static const u64 MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 INV = 0xc2e1f593efffffffULL;  // -r^-1 mod 2^64
static const Fr ONE = {{0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL}};
static const Fr R2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
// R^3 mod r, for from_u512
static const Fr ZERO = {{0, 0, 0, 0}};

inline bool geq_mod(const u64 t[4]) {
    for (int i = 3; i >= 0; i--) {
        if (t[i] > MOD[i]) return true;
        if (t[i] < MOD[i]) return false;
    }
    return true;
}
inline void sub_mod(u64 t[4]) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)t[i] - MOD[i] - br;
        t[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
}
inline Fr add(const Fr &a, const Fr &b) {
    Fr r;
    u64 c = 0;
    for (int i = 0; i < 4; i++) {
        u128 s = (u128)a.v[i] + b.v[i] + c;
        r.v[i] = (u64)s;
        c = (u64)(s >> 64);
    }
    if (c || geq_mod(r.v)) sub_mod(r.v);
    return r;
}
inline Fr sub(const Fr &a, const Fr &b) {
    Fr r;
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.v[i] - b.v[i] - br;
        r.v[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
    if (br) {
        u64 c = 0;
        for (int i = 0; i < 4; i++) {
            u128 s = (u128)r.v[i] + MOD[i] + c;
            r.v[i] = (u64)s;
            c = (u64)(s >> 64);
        }
    }
    return r;
}
inline Fr neg(const Fr &a) { return sub(ZERO, a); }
inline Fr mul(const Fr &a, const Fr &b) {
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
        u64 m = t[0] * INV;
        acc = (u128)m * MOD[0] + t[0];
        carry = (u64)(acc >> 64);
        for (int j = 1; j < 4; j++) {
            acc = (u128)m * MOD[j] + t[j] + carry;
            t[j - 1] = (u64)acc;
            carry = (u64)(acc >> 64);
        }
        acc = (u128)t[4] + carry;
        t[3] = (u64)acc;
        t[4] = t[5] + (u64)(acc >> 64);
    }
    Fr r;
    memcpy(r.v, t, 32);
    if (t[4] || geq_mod(r.v)) sub_mod(r.v);
    return r;
}
inline Fr sqr(const Fr &a) { return mul(a, a); }
inline Fr pow(const Fr &a, const u64 e[4]) {
    Fr acc = ONE;
    for (int i = 255; i >= 0; i--) {
        acc = sqr(acc);
        if ((e[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
    }
    return acc;
}
inline Fr pow_u64(const Fr &a, u64 e) {
    u64 ee[4] = {e, 0, 0, 0};
    return pow(a, ee);
}
inline Fr inv(const Fr &a) {
    u64 e[4];
    memcpy(e, MOD, 32);
    e[0] -= 2;
    return pow(a, e);
}
inline bool is_zero(const Fr &a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
// canonical little-endian limbs -> Montgomery (value must be < r)
inline Fr from_canonical(const u64 c[4]) {
    Fr t;
    memcpy(t.v, c, 32);
    return mul(t, R2);
}
inline void to_canonical(const Fr &a, u64 out[4]) {
    Fr one_c = {{1, 0, 0, 0}};
    Fr t = mul(a, one_c);
    memcpy(out, t.v, 32);
}
inline Fr from_u64(u64 x) {
    u64 c[4] = {x, 0, 0, 0};
    return from_canonical(c);
}
// 512-bit little-endian integer (8 limbs) mod r -> Montgomery  (halo2curves `Fr::from_u512`)
inline Fr from_u512(const u64 l[8]) {
    Fr lo, hi;
    memcpy(lo.v, l, 32);
    memcpy(hi.v, l + 4, 32);
    // lo, hi are arbitrary 256-bit values (possibly >= r): mul() tolerates inputs < 2^256 because the CIOS
    // bound only needs one operand < r; R2 / R3 are.  lo*R2*R^-1 = lo*R ; hi*R3*R^-1 = hi*R^2 = (hi*2^256)*R
    static const Fr R3 = mul(R2, R2);
    return add(mul(lo, R2), mul(hi, R3));
}
// 256-bit big/little-endian byte string mod r -> Montgomery
inline Fr from_bytes_le_wide(const uint8_t *bytes, int len) {  // len <= 64
    u64 l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < len; i++) l[i >> 3] |= (u64)bytes[i] << (8 * (i & 7));
    return from_u512(l);
}
inline void to_bytes_le(const Fr &a, uint8_t out[32]) {
    u64 c[4];
    to_canonical(a, c);
    for (int i = 0; i < 32; i++) out[i] = (uint8_t)(c[i >> 3] >> (8 * (i & 7)));
}
inline void to_bytes_be(const Fr &a, uint8_t out[32]) {
    uint8_t le[32];
    to_bytes_le(a, le);
    for (int i = 0; i < 32; i++) out[i] = le[31 - i];
}
// numeric comparison of canonical values (halo2curves `impl Ord for Fr`)
inline int cmp(const Fr &a, const Fr &b) {
    u64 ca[4], cb[4];
    to_canonical(a, ca);
    to_canonical(b, cb);
    for (int i = 3; i >= 0; i--) {
        if (ca[i] < cb[i]) return -1;
        if (ca[i] > cb[i]) return 1;
    }
    return 0;
}

}  // namespace hfr
}  // namespace sb
