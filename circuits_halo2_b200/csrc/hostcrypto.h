// Host-side hashing / RNG of create_proof (Fiat-Shamir stays on the host):
//   Keccak-256 (halo2_solidity_verifier::Keccak256Transcript, utils.rs:170),
//   Blake2b-512 with personalisation (halo2 Blake2bWrite, utils.rs:93),
//   rand_chacha 0.3.1 ChaCha20Rng (the seeded blinding RNG of the byte-parity runs).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "hostfr.h"

namespace sb {

void keccak256(const uint8_t *data, size_t len, uint8_t out[32]);

struct Blake2b {
    uint64_t h[8], t[2];
    uint8_t buf[128];
    size_t buflen;
    void init(size_t outlen, const uint8_t personal[16]);
    void update(const uint8_t *in, size_t len);
    void final(uint8_t *out /* 64 */) const;  // does not disturb the running state (halo2 clones the hasher)
};

struct ChaCha20Rng {
    uint32_t key[8];
    uint64_t counter;
    uint32_t block[16];
    int index;  // next unread word of `block` (16 = empty)
    void seed(const uint8_t seed32[32]);
    void seed_from_u64(uint64_t state);  // rand_core SeedableRng::seed_from_u64 (PCG32 expansion)
    uint32_t next_u32();
    uint64_t next_u64();
    void fill_bytes(uint8_t *out, size_t n);
    hfr::Fr next_fr();  // halo2curves Fr::random: 8 x next_u64 -> from_u512
};

}  // namespace sb
