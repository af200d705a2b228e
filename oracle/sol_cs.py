"""CPU ORACLE tooling -- recover the PLONKish constraint system of the reference circuit from its
generated verifier (contracts/src/InclusionVerifier.sol:495-1002), which states every gate polynomial,
the lookup argument's input/table expressions, the permutation columns and the query order explicitly.

The result (tests/golden/mst_inclusion_cs.json, written by tests/golden/make_golden.py) is the *shape*
halo2's `ConstraintSystem` has for `MstInclusionCircuit<LEVELS, 2, 8>` after selector compression
(SURVEY.md 8 / A.5 / A.8): it is what `keygen_vk` would hand to `create_proof` as `pk.vk.cs`, and what
the C ABI's `sb_create_proof` consumes as its constraint-system description.

Expression JSON: ["const", "0x.."] | ["advice"|"fixed"|"instance", column, rotation] |
                 ["neg", e] | ["add", e, e] | ["mul", e, e]
"""
from __future__ import annotations

import re
from typing import Dict, List

PROOF_CPTR = 0x64


def _split_blocks(section: str) -> List[str]:
    """top-level { ... } blocks of a Yul statement list"""
    out, depth, start = [], 0, None
    for i, ch in enumerate(section):
        if ch == "{":
            if depth == 0:
                start = i
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                out.append(section[start + 1:i])
    return out


def _name_to_query(name: str):
    m = re.fullmatch(r"([afi])_(\d+)(?:_(next|prev)_(\d+))?", name)
    if not m:
        return None
    kind = {"a": "advice", "f": "fixed", "i": "instance"}[m.group(1)]
    rot = 0
    if m.group(3):
        rot = int(m.group(4)) * (1 if m.group(3) == "next" else -1)
    return [kind, int(m.group(2)), rot]


def _parse_ssa(block: str, offsets: Dict[int, list]):
    """evaluate a straight-line `let x := op(..)` block symbolically; returns env name -> expression"""
    env: Dict[str, list] = {}

    def operand(tok: str):
        tok = tok.strip()
        if tok in env:
            return env[tok]
        if re.fullmatch(r"0x[0-9a-fA-F]+|\d+", tok):
            return ["const", hex(int(tok, 0))]
        raise KeyError(tok)

    for m in re.finditer(r"let (\w+) := ([^\n]+)", block):
        name, rhs = m.group(1), m.group(2).strip()
        mm = re.fullmatch(r"calldataload\((0x[0-9a-f]+)\)", rhs)
        if mm:
            q = _name_to_query(name)
            if q is None:
                raise ValueError(f"unnamed evaluation {name}")
            off = int(mm.group(1), 16)
            if off in offsets and offsets[off] != q:
                raise ValueError(f"offset {off:#x} maps to two queries")
            offsets[off] = q
            env[name] = q
            continue
        mm = re.fullmatch(r"(addmod|mulmod)\(([^,]+), ([^,]+), r\)", rhs)
        if mm:
            env[name] = ["add" if mm.group(1) == "addmod" else "mul", operand(mm.group(2)), operand(mm.group(3))]
            continue
        mm = re.fullmatch(r"sub\(r, ([^)]+)\)", rhs)
        if mm:
            env[name] = ["neg", operand(mm.group(1))]
            continue
        if re.fullmatch(r"0x[0-9a-fA-F]+|\d+", rhs):
            env[name] = ["const", hex(int(rhs, 0))]
            continue
        if name in ("theta", "beta", "gamma", "input", "table", "lhs", "rhs", "eval", "l_0", "l_last", "perm_z_last", "left_sub_right"):
            continue
        raise ValueError(f"unparsed statement: let {name} := {rhs}")
    return env


def constraint_system_from_sol(sol: str) -> dict:
    start = sol.index("// Compute quotient evavluation")
    end = sol.index("pop(y)", start)
    body = sol[start:end]
    body = body[body.index("let y := mload(Y_MPTR)"):]
    blocks = _split_blocks(body)
    offsets: Dict[int, list] = {}
    gates, lookups, perm_cols = [], [], []
    for blk in blocks:
        if "L_0_MPTR" in blk or "L_LAST_MPTR" in blk or "L_BLIND_MPTR" in blk:
            if "let input" in blk:  # lookup product term: nested input / table blocks
                inner = _split_blocks(blk)
                assert len(inner) == 2
                env_i = _parse_ssa(inner[0], offsets)
                env_t = _parse_ssa(inner[1], offsets)
                mi = re.search(r"input := (\w+)", inner[0])
                mt = re.search(r"table := (\w+)", inner[1])
                lookups.append({"input": [env_i[mi.group(1)]], "table": [env_t[mt.group(1)]]})
            elif "let gamma" in blk:  # permutation product term: (value, sigma) pairs on the lhs
                for m in re.finditer(r"lhs := mulmod\(lhs, addmod\(addmod\((calldataload\((0x[0-9a-f]+)\)|mload\(INSTANCE_EVAL_MPTR\)), mulmod\(beta, calldataload\((0x[0-9a-f]+)\)", blk):
                    perm_cols.append({"value_offset": int(m.group(2), 16) if m.group(2) else None, "sigma_offset": int(m.group(3), 16)})
            continue
        env = _parse_ssa(blk, offsets)
        m = re.search(r"quotient_eval_numer := (?:addmod\(mulmod\(quotient_eval_numer, y, r\), )?(\w+)", blk)
        gates.append(env[m.group(1)])

    # evaluation order = calldata order
    def proof_off(o):
        return o - PROOF_CPTR
    advice_q = [offsets[o] for o in sorted(offsets) if offsets[o][0] == "advice"]
    fixed_q = [offsets[o] for o in sorted(offsets) if offsets[o][0] == "fixed"]
    inv = {tuple(v): k for k, v in offsets.items()}
    permutation = []
    for pc in perm_cols:
        if pc["value_offset"] is None:
            permutation.append(["instance", 0])
        else:
            kind, col, rot = offsets[pc["value_offset"]]
            assert rot == 0
            permutation.append([kind, col])
    n_adv = 1 + max(q[1] for q in advice_q)
    n_fix = 1 + max(q[1] for q in fixed_q)
    m = re.search(r"mstore\(0x[0-9a-f]+, (0x[0-9a-f]{64})\) // num_instances", sol)
    evals_start = proof_off(min(offsets))
    cs = {
        "_source": "derived from contracts/src/InclusionVerifier.sol:495-1002 by oracle/sol_cs.py",
        "num_advice_columns": n_adv, "num_fixed_columns": n_fix, "num_instance_columns": 1,
        "num_instances": int(m.group(1), 16),
        "advice_queries": [[q[1], q[2]] for q in advice_q],
        "fixed_queries": [[q[1], q[2]] for q in fixed_q],
        "instance_queries": [[0, 0]],
        "gates": gates,
        "lookups": lookups,
        "permutation_columns": permutation,
        "degree": 6,               # max gate degree: quotient = 5 pieces (.sol:11-12), extended_k = k + 3
        "blinding_factors": 5,     # omega_inv_to_l = omega^-6 (.sol:222)
        "evals_proof_offset": evals_start,
        "sigma_eval_offsets": [proof_off(pc["sigma_offset"]) for pc in perm_cols],
    }
    return cs
