"""Judge a GPU-made proof with the REFERENCE's verifier contract (build container only).

    python tools/verify_with_reference.py gpurun_out/proof_k17.npz [tau]

The proof dump (bench.py --dump-proof) holds the proof bytes, the key's fixed / permutation commitments, the
instances and k.  The verifier program is contracts/src/InclusionVerifier.sol (fixture tests/golden/InclusionVerifier.sol), interpreted by oracle/yul.py, with
its embedded verifying-key constants replaced by this key's (k, domain constants, commitments, transcript
representative, and -tau.G2 of the unsafe synthetic SRS)."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from oracle import bn254 as B  # noqa: E402

from oracle.reference_verifier import verifier_for_key  # noqa: E402


def main():
    d = np.load(sys.argv[1])
    k = int(d["k"][0])
    tau = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x5A110000 + k
    pts = lambda a: [B.g1_from_mont_bytes(c.tobytes()) for c in a]
    v = verifier_for_key(k, tau, pts(d["fixed_comms"]), pts(d["sigma_comms"]), int(d["transcript_repr"][0]))
    proof = d["proof"].tobytes()
    inst = [B.fr_from_mont_bytes(x.tobytes()) for x in d["instances"]]
    ok = v.verify(proof, inst)
    bad = bytearray(proof)
    bad[0x400] ^= 1
    print(f"k={k}: reference verifier accepts the GPU proof: {ok}; tampered proof accepted: {v.verify(bytes(bad), inst)}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
