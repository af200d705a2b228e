"""CPU ORACLE (test infrastructure) -- the two Fiat-Shamir transcripts the reference uses.

* `KeccakTranscript`  = halo2_solidity_verifier::Keccak256Transcript (utils.rs:170; the on-chain
  verifier's view is InclusionVerifier.sol:72-112,279-367): buffer of 32-byte big-endian words,
  challenge = keccak256(buffer) mod r, buffer <- hash, and a lone 0x01 is appended when a second
  challenge is squeezed with nothing absorbed in between.  Points are written as x || y (64 B).
* `Blake2bTranscript` = halo2_proofs::transcript::Blake2bWrite<_, _, Challenge255<_>> (utils.rs:93):
  Blake2b-512, personal "Halo2-Transcript", prefix bytes 0 challenge / 1 point / 2 scalar,
  little-endian reprs, points written compressed (32 B: x LE, sign of y in bit 6 of the last byte).
"""
import hashlib

from . import bn254 as B
from .keccak import keccak256


class KeccakTranscript:
    def __init__(self):
        self.buf = bytearray()
        self.proof = bytearray()

    def common_scalar(self, s: int):
        self.buf += int(s).to_bytes(32, "big")

    def common_point(self, p):
        if p is None:
            raise ValueError("Cannot write points at infinity to the transcript")
        self.buf += p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")

    def write_point(self, p):
        self.common_point(p)
        self.proof += p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")

    def write_scalar(self, s: int):
        self.common_scalar(s)
        self.proof += int(s).to_bytes(32, "big")

    def squeeze_challenge(self) -> int:
        data = bytes(self.buf) + (b"\x01" if len(self.buf) == 0x20 else b"")
        h = keccak256(data)
        self.buf = bytearray(h)
        return int.from_bytes(h, "big") % B.R

    def finalize(self) -> bytes:
        return bytes(self.proof)


class Blake2bTranscript:
    def __init__(self):
        self.state = hashlib.blake2b(digest_size=64, person=b"Halo2-Transcript")
        self.proof = bytearray()

    def common_scalar(self, s: int):
        self.state.update(b"\x02" + int(s).to_bytes(32, "little"))

    def common_point(self, p):
        if p is None:
            raise ValueError("cannot write points at infinity to the transcript")
        self.state.update(b"\x01" + p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little"))

    @staticmethod
    def compress(p) -> bytes:
        if p is None:
            b = bytearray(32)
            b[31] |= 0x80
            return bytes(b)
        b = bytearray(p[0].to_bytes(32, "little"))
        b[31] |= (p[1] & 1) << 6
        return bytes(b)

    def write_point(self, p):
        self.common_point(p)
        self.proof += self.compress(p)

    def write_scalar(self, s: int):
        self.common_scalar(s)
        self.proof += int(s).to_bytes(32, "little")

    def squeeze_challenge(self) -> int:
        self.state.update(b"\x00")
        h = self.state.copy().digest()
        return int.from_bytes(h, "little") % B.R

    def finalize(self) -> bytes:
        return bytes(self.proof)
