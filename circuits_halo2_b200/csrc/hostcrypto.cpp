#include "hostcrypto.h"

#include <string.h>

namespace sb {

// ------------------------------------------------------------------ Keccak-256 (original padding 0x01)
static const uint64_t KRC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
                                 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
                                 0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
                                 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
                                 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KPIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
static inline uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

static void keccak_f(uint64_t st[25]) {
    for (int round = 0; round < 24; round++) {
        uint64_t bc[5];
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) {
            uint64_t t = bc[(i + 4) % 5] ^ rol64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; i++) {
            int j = KPIL[i];
            uint64_t b = st[j];
            st[j] = rol64(t, KROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; i++) bc[i] = st[j + i];
            for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KRC[round];
    }
}

void keccak256(const uint8_t *data, size_t len, uint8_t out[32]) {
    const size_t rate = 136;
    uint64_t st[25];
    memset(st, 0, sizeof st);
    auto absorb = [&](const uint8_t *blk) {
        for (size_t i = 0; i < rate / 8; i++) {
            uint64_t w = 0;
            for (int b = 0; b < 8; b++) w |= (uint64_t)blk[8 * i + b] << (8 * b);
            st[i] ^= w;
        }
        keccak_f(st);
    };
    while (len >= rate) {
        absorb(data);
        data += rate;
        len -= rate;
    }
    uint8_t last[136];
    memset(last, 0, sizeof last);
    memcpy(last, data, len);
    last[len] = 0x01;
    last[rate - 1] |= 0x80;
    absorb(last);
    for (int i = 0; i < 4; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(st[i] >> (8 * b));
}

// ------------------------------------------------------------------ Blake2b
static const uint64_t B2IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                                 0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
static const uint8_t B2SIGMA[12][16] = {{0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
                                        {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
                                        {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
                                        {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
                                        {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
                                        {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

static void b2_compress(uint64_t h[8], const uint8_t block[128], const uint64_t t[2], bool last) {
    uint64_t m[16], v[16];
    for (int i = 0; i < 16; i++) {
        m[i] = 0;
        for (int b = 0; b < 8; b++) m[i] |= (uint64_t)block[8 * i + b] << (8 * b);
    }
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2IV[i]; }
    v[12] ^= t[0];
    v[13] ^= t[1];
    if (last) v[14] = ~v[14];
#define B2G(a, b, c, d, x, y) \
    v[a] = v[a] + v[b] + (x); v[d] = ror64(v[d] ^ v[a], 32); v[c] = v[c] + v[d]; v[b] = ror64(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + (y); v[d] = ror64(v[d] ^ v[a], 16); v[c] = v[c] + v[d]; v[b] = ror64(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2SIGMA[r];
        B2G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2G(1, 5, 9, 13, m[s[2]], m[s[3]]) B2G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2G(3, 7, 11, 15, m[s[6]], m[s[7]])
        B2G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2G(1, 6, 11, 12, m[s[10]], m[s[11]]) B2G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef B2G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}

void Blake2b::init(size_t outlen, const uint8_t personal[16]) {
    for (int i = 0; i < 8; i++) h[i] = B2IV[i];
    h[0] ^= 0x01010000ULL ^ (uint64_t)outlen;  // digest length, key length 0, fanout 1, depth 1
    if (personal) {
        uint64_t p0 = 0, p1 = 0;
        for (int b = 0; b < 8; b++) { p0 |= (uint64_t)personal[b] << (8 * b); p1 |= (uint64_t)personal[8 + b] << (8 * b); }
        h[6] ^= p0;
        h[7] ^= p1;
    }
    t[0] = t[1] = 0;
    buflen = 0;
    memset(buf, 0, sizeof buf);
}
void Blake2b::update(const uint8_t *in, size_t len) {
    while (len > 0) {
        if (buflen == 128) {  // buffer full and more input follows: compress it (never the last block here)
            t[0] += 128;
            if (t[0] < 128) t[1]++;
            b2_compress(h, buf, t, false);
            buflen = 0;
        }
        size_t take = 128 - buflen;
        if (take > len) take = len;
        memcpy(buf + buflen, in, take);
        buflen += take;
        in += take;
        len -= take;
    }
}
void Blake2b::final(uint8_t *out) const {
    uint64_t hh[8], tt[2] = {t[0], t[1]};
    memcpy(hh, h, sizeof hh);
    uint8_t last[128];
    memset(last, 0, sizeof last);
    memcpy(last, buf, buflen);
    tt[0] += buflen;
    if (tt[0] < buflen) tt[1]++;
    b2_compress(hh, last, tt, true);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 8; b++) out[8 * i + b] = (uint8_t)(hh[i] >> (8 * b));
}

// ------------------------------------------------------------------ ChaCha20Rng (rand_chacha 0.3.1)
static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define CQR(a, b, c, d) \
    a += b; d = rotl32(d ^ a, 16); c += d; b = rotl32(b ^ c, 12); a += b; d = rotl32(d ^ a, 8); c += d; b = rotl32(b ^ c, 7);
static void chacha_block(const uint32_t key[8], uint64_t counter, uint32_t out[16]) {
    uint32_t init[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                         (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
    uint32_t s[16];
    memcpy(s, init, sizeof s);
    for (int i = 0; i < 10; i++) {
        CQR(s[0], s[4], s[8], s[12]) CQR(s[1], s[5], s[9], s[13]) CQR(s[2], s[6], s[10], s[14]) CQR(s[3], s[7], s[11], s[15])
        CQR(s[0], s[5], s[10], s[15]) CQR(s[1], s[6], s[11], s[12]) CQR(s[2], s[7], s[8], s[13]) CQR(s[3], s[4], s[9], s[14])
    }
    for (int i = 0; i < 16; i++) out[i] = s[i] + init[i];
}
void ChaCha20Rng::seed(const uint8_t seed32[32]) {
    for (int i = 0; i < 8; i++) key[i] = (uint32_t)seed32[4 * i] | ((uint32_t)seed32[4 * i + 1] << 8) | ((uint32_t)seed32[4 * i + 2] << 16) | ((uint32_t)seed32[4 * i + 3] << 24);
    counter = 0;
    index = 16;
}
void ChaCha20Rng::seed_from_u64(uint64_t state) {
    const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
    uint8_t s[32];
    for (int i = 0; i < 8; i++) {
        state = state * MUL + INC;
        uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
        uint32_t rot = (uint32_t)(state >> 59);
        uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
        s[4 * i] = (uint8_t)x; s[4 * i + 1] = (uint8_t)(x >> 8); s[4 * i + 2] = (uint8_t)(x >> 16); s[4 * i + 3] = (uint8_t)(x >> 24);
    }
    seed(s);
}
uint32_t ChaCha20Rng::next_u32() {
    if (index >= 16) {
        chacha_block(key, counter++, block);
        index = 0;
    }
    return block[index++];
}
uint64_t ChaCha20Rng::next_u64() {
    uint64_t lo = next_u32();
    uint64_t hi = next_u32();
    return lo | (hi << 32);
}
void ChaCha20Rng::fill_bytes(uint8_t *out, size_t n) {
    size_t i = 0;
    while (i < n) {
        uint32_t w = next_u32();
        for (int b = 0; b < 4 && i < n; b++, i++) out[i] = (uint8_t)(w >> (8 * b));
    }
}
hfr::Fr ChaCha20Rng::next_fr() {
    uint64_t l[8];
    for (int i = 0; i < 8; i++) l[i] = next_u64();
    return hfr::from_u512(l);
}

}  // namespace sb

// ------------------------------------------------------------------ test hooks (exercised by the CPU test-suite)
extern "C" {
int32_t sb_test_keccak256(const uint8_t *data, size_t len, uint8_t out[32]) { sb::keccak256(data, len, out); return 0; }
int32_t sb_test_blake2b512(const uint8_t *data, size_t len, const uint8_t personal[16], uint8_t out[64]) {
    sb::Blake2b b;
    b.init(64, personal);
    // feed in two pieces to exercise buffering
    size_t half = len / 2;
    b.update(data, half);
    b.update(data + half, len - half);
    b.final(out);
    return 0;
}
int32_t sb_test_chacha_fr(uint64_t seed_u64, uint32_t skip_bytes, uint32_t count, uint8_t *out /* count x 32 B Montgomery */) {
    sb::ChaCha20Rng r;
    r.seed_from_u64(seed_u64);
    std::vector<uint8_t> tmp(skip_bytes);
    if (skip_bytes) r.fill_bytes(tmp.data(), skip_bytes);
    for (uint32_t i = 0; i < count; i++) {
        sb::hfr::Fr x = r.next_fr();
        memcpy(out + 32 * i, x.v, 32);
    }
    return 0;
}
int32_t sb_test_host_fr(int32_t op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
    sb::hfr::Fr x, y, z;
    memcpy(x.v, a, 32);
    memcpy(y.v, b, 32);
    z = op == 0 ? sb::hfr::mul(x, y) : op == 1 ? sb::hfr::add(x, y) : op == 2 ? sb::hfr::sub(x, y) : sb::hfr::inv(x);
    memcpy(out, z.v, 32);
    return 0;
}
}
