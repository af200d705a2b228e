"""Judge a GPU-made proof with the REFERENCE's verifier contract (build container only).

    python tools/verify_with_reference.py gpurun_out/proof_k17.npz [tau]

The proof dump (bench.py --dump-proof) holds the proof bytes, the key's fixed / permutation commitments, the
instances and k.  The verifier program is contracts/src/InclusionVerifier.sol, interpreted by oracle/yul.py, with
its embedded verifying-key constants replaced by this key's (k, domain constants, commitments, transcript
representative, and -tau.G2 of the unsafe synthetic SRS)."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from oracle import bn254 as B  # noqa: E402
from oracle import pairing as P  # noqa: E402
from oracle.yul import SolidityVerifier  # noqa: E402

SOL = "/root/reference/contracts/src/InclusionVerifier.sol"
Q = B.Q


def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, -a[1] * d % Q)


def f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if y1 != y2:
            return None
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_mul((2, 0), y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_mul(p, k):
    acc = None
    while k:
        if k & 1:
            acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def main():
    d = np.load(sys.argv[1])
    k = int(d["k"][0])
    tau = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x5A110000 + k
    dom = B.EvaluationDomain(6, k)
    rep = {"k": k, "n_inv": dom.ifft_divisor, "omega": dom.omega, "omega_inv": dom.omega_inv, "omega_inv_to_l": pow(dom.omega_inv, 6, B.R),
           "vk_digest": int(d["transcript_repr"][0])}
    for i, c in enumerate(d["fixed_comms"]):
        x, y = B.g1_from_mont_bytes(c.tobytes())
        rep[f"fixed_comms[{i}].x"], rep[f"fixed_comms[{i}].y"] = x, y
    for i, c in enumerate(d["sigma_comms"]):
        x, y = B.g1_from_mont_bytes(c.tobytes())
        rep[f"permutation_comms[{i}].x"], rep[f"permutation_comms[{i}].y"] = x, y
    sg2 = g2_mul(P.G2_GEN, tau)
    neg = (sg2[0], ((-sg2[1][0]) % Q, (-sg2[1][1]) % Q))
    rep.update({"neg_s_g2_x_1": neg[0][1], "neg_s_g2_x_2": neg[0][0], "neg_s_g2_y_1": neg[1][1], "neg_s_g2_y_2": neg[1][0]})
    v = SolidityVerifier.from_file(SOL).patched(rep)
    proof = d["proof"].tobytes()
    inst = [B.fr_from_mont_bytes(x.tobytes()) for x in d["instances"]]
    ok = v.verify(proof, inst)
    bad = bytearray(proof)
    bad[0x400] ^= 1
    print(f"k={k}: reference verifier accepts the GPU proof: {ok}; tampered proof accepted: {v.verify(bytes(bad), inst)}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
