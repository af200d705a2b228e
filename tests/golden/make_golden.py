"""Regenerates tests/golden/* from the reference tree (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Outputs (all *data*, no reference source code):
  verifier_constants.json  vk / domain / SRS constants embedded in contracts/src/InclusionVerifier.sol:217-271
  hermez-raw-11            copy of backend/ptau/hermez-raw-11 (KZG SRS, k = 11; 262 404 B)
  inclusion_proof_solidity_calldata.json, commitment_solidity_calldata.json   zk_prover/examples/*.json
  entry_16.csv             csv/entry_16.csv
  mst_hashes.json          known-answer hashes quoted in the Rust tests
  poseidon_params.json     Poseidon round constants / MDS of the circuit (chips/poseidon/poseidon_params.rs)
  mst_inclusion_cs.json    the circuit's constraint system (gates, lookup, permutation, queries) recovered from the verifier
"""
import json
import os
import re
import shutil
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))

sol = open(os.path.join(ref, "contracts/src/InclusionVerifier.sol")).read()
consts = {}
for m in re.finditer(r"mstore\(0x[0-9a-f]+, (0x[0-9a-f]{64})\) // ([a-z_0-9\[\]\.]+)", sol):
    consts[m.group(2)] = m.group(1)
consts["_source"] = "contracts/src/InclusionVerifier.sol:217-271"
m = re.search(r"let delta := (\d+)", sol)
if m:
    consts["delta"] = hex(int(m.group(1)))
json.dump(consts, open(os.path.join(here, "verifier_constants.json"), "w"), indent=1, sort_keys=True)

shutil.copy(os.path.join(ref, "backend/ptau/hermez-raw-11"), os.path.join(here, "hermez-raw-11"))
for f in ("inclusion_proof_solidity_calldata.json", "commitment_solidity_calldata.json"):
    shutil.copy(os.path.join(ref, "zk_prover/examples", f), os.path.join(here, f))
shutil.copy(os.path.join(ref, "csv/entry_16.csv"), os.path.join(here, "entry_16.csv"))

tests_rs = open(os.path.join(ref, "zk_prover/src/circuits/tests.rs")).read()
backend_rs = open(os.path.join(ref, "backend/src/tests.rs")).read()
hashes = sorted(set(re.findall(r"0x[0-9a-f]{62,64}", tests_rs)))
root = sorted(set(re.findall(r"0x[0-9a-f]{62,64}", backend_rs)))
json.dump({"_source": "zk_prover/src/circuits/tests.rs (leaf hashes), backend/src/tests.rs (root hash)",
           "circuit_tests_hex": hashes, "backend_tests_hex": root},
          open(os.path.join(here, "mst_hashes.json"), "w"), indent=1)
# Poseidon constants (data): zk_prover/src/chips/poseidon/poseidon_params.rs:18-987
pp = open(os.path.join(ref, "zk_prover/src/chips/poseidon/poseidon_params.rs")).read()
def _consts(block):
    vals = []
    for m in re.finditer(r"Fp::from_raw\(\[\s*(0x[0-9a-f_]+),\s*(0x[0-9a-f_]+),\s*(0x[0-9a-f_]+),\s*(0x[0-9a-f_]+),?\s*\]\)", block):
        limbs = [int(g.replace("_", ""), 16) for g in m.groups()]
        vals.append(hex(sum(l << (64 * i) for i, l in enumerate(limbs))))
    return vals
i_rc, i_mds, i_inv = pp.index("ROUND_CONSTANTS"), pp.index("pub(crate) const MDS:"), pp.index("pub(crate) const MDS_INV")
rc, mds, mds_inv = _consts(pp[i_rc:i_mds]), _consts(pp[i_mds:i_inv]), _consts(pp[i_inv:])
assert len(rc) == 128 and len(mds) == 4 and len(mds_inv) == 4
json.dump({"_source": "zk_prover/src/chips/poseidon/poseidon_params.rs:18-987 (WIDTH 2, RATE 1, R_F 8, R_P 56)",
           "round_constants": [rc[2 * i:2 * i + 2] for i in range(64)], "mds": [mds[0:2], mds[2:4]], "mds_inv": [mds_inv[0:2], mds_inv[2:4]]},
          open(os.path.join(here, "poseidon_params.json"), "w"))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle.sol_cs import constraint_system_from_sol  # noqa: E402
cs = constraint_system_from_sol(sol)
json.dump(cs, open(os.path.join(here, "mst_inclusion_cs.json"), "w"), indent=None, separators=(",", ":"))
print("wrote", sorted(os.listdir(here)))
