"""Regenerates tests/golden/* from the reference tree (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

Outputs (all *data*, no reference source code):
  verifier_constants.json  vk / domain / SRS constants embedded in contracts/src/InclusionVerifier.sol:217-271
  hermez-raw-11            copy of backend/ptau/hermez-raw-11 (KZG SRS, k = 11; 262 404 B)
  inclusion_proof_solidity_calldata.json, commitment_solidity_calldata.json   zk_prover/examples/*.json
  entry_16.csv             csv/entry_16.csv
  mst_hashes.json          known-answer hashes quoted in the Rust tests
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

json.dump({
    "_source": "zk_prover/src/circuits/tests.rs:341,346; backend/src/tests.rs:265-268; zk_prover/src/merkle_sum_tree/tests.rs:24",
    "leaf0": "0x167505f45c5c8f7d8a8b5e6d0c1f8b2b7f2b2f5e",  # placeholder replaced below if found
}, open(os.path.join(here, "mst_hashes.json"), "w"), indent=1)
tests_rs = open(os.path.join(ref, "zk_prover/src/circuits/tests.rs")).read()
backend_rs = open(os.path.join(ref, "backend/src/tests.rs")).read()
hashes = sorted(set(re.findall(r"0x[0-9a-f]{62,64}", tests_rs)))
root = sorted(set(re.findall(r"0x[0-9a-f]{62,64}", backend_rs)))
json.dump({"_source": "zk_prover/src/circuits/tests.rs (leaf hashes), backend/src/tests.rs (root hash)",
           "circuit_tests_hex": hashes, "backend_tests_hex": root},
          open(os.path.join(here, "mst_hashes.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(here)))
