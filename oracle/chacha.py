"""CPU ORACLE (test infrastructure) -- rand_chacha 0.3.1 `ChaCha20Rng` restated (Cargo.lock pins
rand_chacha 0.3.1 as a dependency of halo2_proofs): 256-bit seed = key, 64-bit block counter in
words 12-13, stream id 0 in words 14-15, 20 rounds, output consumed as little-endian u32 words.
`next_u64` = two consecutive words (low first); `fill_bytes(n)` consumes ceil(n/4) words.
`Fr::random(rng)` (halo2curves 0.1.0) = the 512-bit little-endian integer built from eight
`next_u64()` outputs, reduced mod r (`from_u512`).  SURVEY A.1 / A.5."""
import struct

from . import bn254 as B

_M = 0xFFFFFFFF


def _rotl(x, n):
    return ((x << n) & _M) | (x >> (32 - n))


def _qr(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & _M; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _M; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _M; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _M; s[b] = _rotl(s[b] ^ s[c], 7)


def chacha20_block(key_words, counter, stream=0):
    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + \
           [counter & _M, (counter >> 32) & _M, stream & _M, (stream >> 32) & _M]
    s = list(init)
    for _ in range(10):
        _qr(s, 0, 4, 8, 12); _qr(s, 1, 5, 9, 13); _qr(s, 2, 6, 10, 14); _qr(s, 3, 7, 11, 15)
        _qr(s, 0, 5, 10, 15); _qr(s, 1, 6, 11, 12); _qr(s, 2, 7, 8, 13); _qr(s, 3, 4, 9, 14)
    return [(x + y) & _M for x, y in zip(s, init)]


class ChaCha20Rng:
    def __init__(self, seed: bytes):
        assert len(seed) == 32
        self.key = struct.unpack("<8I", seed)
        self.counter = 0
        self.buf = []

    @classmethod
    def seed_from_u64(cls, state: int) -> "ChaCha20Rng":
        """rand_core `SeedableRng::seed_from_u64` (PCG32 expansion of the u64 into the 32-byte seed)."""
        mul, inc = 6364136223846793005, 11634580027462260723
        out = b""
        for _ in range(8):
            state = (state * mul + inc) & 0xFFFFFFFFFFFFFFFF
            xorshifted = (((state >> 18) ^ state) >> 27) & _M
            rot = state >> 59
            x = ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & _M
            out += struct.pack("<I", x)
        return cls(out)

    def next_u32(self) -> int:
        if not self.buf:
            self.buf = chacha20_block(self.key, self.counter)
            self.counter += 1
        return self.buf.pop(0)

    def next_u64(self) -> int:
        lo = self.next_u32()
        hi = self.next_u32()
        return lo | (hi << 32)

    def fill_bytes(self, n: int) -> bytes:
        words = [self.next_u32() for _ in range((n + 3) // 4)]
        return struct.pack(f"<{len(words)}I", *words)[:n]

    def next_fr(self) -> int:
        """halo2curves `Fr::random`."""
        v = 0
        for i in range(8):
            v |= self.next_u64() << (64 * i)
        return v % B.R
