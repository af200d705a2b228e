"""CPU ORACLE (test infrastructure) -- judge a proof with the REFERENCE's own verifier contract.

The program is contracts/src/InclusionVerifier.sol (committed unchanged as the fixture tests/golden/InclusionVerifier.sol; the test
tests/test_oracle_golden.py::test_verifier_fixture_is_the_reference_file compares it with /root/reference where that tree is mounted),
interpreted by oracle/yul.py.  For keys other than the one the contract embeds (k = 11 on ptau/hermez-raw-11) the verifying-key
CONSTANTS -- k, n^-1, omega, omega^-1, omega^-6, vk digest, the 11 fixed + 6 permutation commitments, -s.G2 of the SRS -- are replaced by
the key's own; the verifier algorithm (transcript, gate / permutation / lookup algebra, SHPLONK, pairing: .sol:72-1400) runs unmodified.
The cost does not depend on k."""
from __future__ import annotations

import os
from typing import Sequence

from . import bn254 as B
from . import pairing as P
from .yul import SolidityVerifier

SOL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "InclusionVerifier.sol")
Q = B.Q


def _f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)


def _f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, -a[1] * d % Q)


def _f2_sub(a, b):
    return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)


def _g2_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if y1 != y2:
            return None
        lam = _f2_mul(_f2_mul((3, 0), _f2_mul(x1, x1)), _f2_inv(_f2_mul((2, 0), y1)))
    else:
        lam = _f2_mul(_f2_sub(y2, y1), _f2_inv(_f2_sub(x2, x1)))
    x3 = _f2_sub(_f2_sub(_f2_mul(lam, lam), x1), x2)
    return (x3, _f2_sub(_f2_mul(lam, _f2_sub(x1, x3)), y1))


def g2_mul(p, k: int):
    acc = None
    while k:
        if k & 1:
            acc = _g2_add(acc, p)
        p = _g2_add(p, p)
        k >>= 1
    return acc


def verifier_for_key(k: int, tau: int, fixed_comms: Sequence, sigma_comms: Sequence, transcript_repr: int) -> SolidityVerifier:
    """The reference contract with this key's constants.  fixed_comms / sigma_comms: affine points as (x, y) ints."""
    dom = B.EvaluationDomain(6, k)
    rep = {"k": k, "n_inv": dom.ifft_divisor, "omega": dom.omega, "omega_inv": dom.omega_inv, "omega_inv_to_l": pow(dom.omega_inv, 6, B.R), "vk_digest": transcript_repr}
    for i, (x, y) in enumerate(fixed_comms):
        rep[f"fixed_comms[{i}].x"], rep[f"fixed_comms[{i}].y"] = x, y
    for i, (x, y) in enumerate(sigma_comms):
        rep[f"permutation_comms[{i}].x"], rep[f"permutation_comms[{i}].y"] = x, y
    sg2 = g2_mul(P.G2_GEN, tau)
    neg = (sg2[0], ((-sg2[1][0]) % Q, (-sg2[1][1]) % Q))
    rep.update({"neg_s_g2_x_1": neg[0][1], "neg_s_g2_x_2": neg[0][0], "neg_s_g2_y_1": neg[1][1], "neg_s_g2_y_2": neg[1][0]})
    return SolidityVerifier.from_file(SOL).patched(rep)


def verify_mont(k: int, tau: int, fixed_comms_mont, sigma_comms_mont, transcript_repr: int, proof: bytes, instances_mont) -> bool:
    """Same, with the key's commitments and the instances as Montgomery uint64 arrays (the C ABI's layout)."""
    import numpy as np
    pts = lambda a: [B.g1_from_mont_bytes(np.ascontiguousarray(c).tobytes()) for c in a]
    v = verifier_for_key(k, tau, pts(fixed_comms_mont), pts(sigma_comms_mont), transcript_repr)
    return v.verify(bytes(proof), [B.fr_from_mont_bytes(np.ascontiguousarray(x).tobytes()) for x in instances_mont])


def expected_key_commitments(k: int, tau: int, n_fixed: int, n_perm: int, fixed_cells, fixed_values_mont, perm_cells):
    """keygen_vk's commitments in CLOSED FORM for an SRS whose secret is known (the unsafe test SRS): commit_lagrange(v) = [sum_i v_i L_i(tau)] G
    with L_i(tau) = omega^i (tau^n - 1) / (n (tau - omega^i)).  A fixed column is the sum over its assigned cells; a permutation column is the
    identity column delta^c omega^i -- whose commitment is [delta^c tau] G because sum_i omega^i L_i(X) = X -- plus the cells copy constraints
    moved.  Independent of every MSM / NTT / SRS array, at any k, in milliseconds: pins the k = 17 / 20 / 23 keys.
    Returns (fixed, sigma) lists of affine points (x, y)."""
    import numpy as np
    from . import cpu
    R = B.R
    n = 1 << k
    dom = B.EvaluationDomain(6, k)
    c0 = (pow(tau, n, R) - 1) * dom.ifft_divisor % R
    lag = {}

    def L(i):
        if i not in lag:
            w = pow(dom.omega, i, R)
            lag[i] = w * c0 % R * pow((tau - w) % R, -1, R) % R
        return lag[i]
    g = np.frombuffer(B.g1_to_mont_bytes((1, 2)), dtype=np.uint64)
    mont = lambda x: np.frombuffer(B.fr_to_mont_bytes(x % R), dtype=np.uint64)
    point = lambda s: B.g1_from_mont_bytes(cpu.g1_mul(g, mont(s)).tobytes())
    fsum = [0] * n_fixed
    for (c, row), v in zip(np.asarray(fixed_cells), np.asarray(fixed_values_mont)):
        fsum[int(c)] = (fsum[int(c)] + B.fr_from_mont_bytes(np.ascontiguousarray(v).tobytes()) * L(int(row))) % R
    ssum = [pow(B.DELTA, c, R) * tau % R for c in range(n_perm)]
    for c, row, tc, trow in np.asarray(perm_cells):
        c, row, tc, trow = int(c), int(row), int(tc), int(trow)
        new = pow(B.DELTA, tc, R) * pow(dom.omega, trow, R) % R
        old = pow(B.DELTA, c, R) * pow(dom.omega, row, R) % R
        ssum[c] = (ssum[c] + (new - old) * L(row)) % R
    return [point(s) for s in fsum], [point(s) for s in ssum]
