"""CPU ORACLE (test infrastructure, NOT the product) -- a small Yul interpreter + EVM precompiles.

Purpose: run the reference's *own* verifier, `contracts/src/InclusionVerifier.sol` (the assembly
block of `verifyProof`, lines 71-1409), on (proof, instances) without a Solidity tool-chain, so the
oracle prover's and the GPU prover's proofs are judged by the reference's executable specification
rather than by a transliteration.  The contract text is read from the reference tree at test time
(CPU container only) or from a caller-supplied string; it is never copied into this repository.
`patch_vk()` rewrites the embedded verifying-key constants so the same program checks other k / vk.

Supported subset = what that file uses: let / := / multi-assign, function definitions, for, if,
blocks, and the builtins add sub mul mod addmod mulmod lt eq and iszero shl pop mload mstore mstore8
calldataload keccak256 gas staticcall (precompiles 5 modexp, 6 ecAdd, 7 ecMul, 8 ecPairing)
revert return.
"""
from __future__ import annotations

import re
from typing import Dict, List, Tuple

from . import bn254 as B
from . import pairing
from .keccak import keccak256

M256 = (1 << 256) - 1

_TOKEN = re.compile(r"\s*(?:(//[^\n]*)|(0x[0-9a-fA-F]+|\d+)|([A-Za-z_][A-Za-z_0-9]*)|(:=|->|[{}(),]))")


def tokenize(src: str) -> List[str]:
    out, pos = [], 0
    src = src.rstrip()
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"yul: bad token at {pos}: {src[pos:pos + 30]!r}")
        pos = m.end()
        if m.group(1):
            continue
        out.append(m.group(2) or m.group(3) or m.group(4))
    return out


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else None

    def eat(self, tok=None):
        v = self.t[self.i]
        if tok is not None and v != tok:
            raise SyntaxError(f"yul: expected {tok!r}, got {v!r} at token {self.i}")
        self.i += 1
        return v

    def block(self):
        self.eat("{")
        stmts = []
        while self.peek() != "}":
            stmts.append(self.stmt())
        self.eat("}")
        return ("block", stmts)

    def stmt(self):
        p = self.peek()
        if p == "{":
            return self.block()
        if p == "let":
            self.eat()
            names = [self.eat()]
            while self.peek() == ",":
                self.eat()
                names.append(self.eat())
            val = None
            if self.peek() == ":=":
                self.eat()
                val = self.expr()
            return ("let", names, val)
        if p == "function":
            self.eat()
            name = self.eat()
            self.eat("(")
            params = []
            while self.peek() != ")":
                params.append(self.eat())
                if self.peek() == ",":
                    self.eat()
            self.eat(")")
            rets = []
            if self.peek() == "->":
                self.eat()
                rets.append(self.eat())
                while self.peek() == ",":
                    self.eat()
                    rets.append(self.eat())
            return ("function", name, params, rets, self.block())
        if p == "for":
            self.eat()
            init = self.block()
            cond = self.expr()
            post = self.block()
            body = self.block()
            return ("for", init, cond, post, body)
        if p == "if":
            self.eat()
            cond = self.expr()
            return ("if", cond, self.block())
        # assignment or expression statement
        if self.peek(1) in (":=", ","):
            names = [self.eat()]
            while self.peek() == ",":
                self.eat()
                names.append(self.eat())
            self.eat(":=")
            return ("assign", names, self.expr())
        return ("expr", self.expr())

    def expr(self):
        tok = self.eat()
        if tok[0].isdigit():
            return ("num", int(tok, 0))
        if self.peek() == "(":
            self.eat("(")
            args = []
            while self.peek() != ")":
                args.append(self.expr())
                if self.peek() == ",":
                    self.eat()
            self.eat(")")
            return ("call", tok, args)
        return ("var", tok)


class Revert(Exception):
    pass


class Return(Exception):
    def __init__(self, data):
        self.data = data


class EVM:
    def __init__(self, calldata: bytes, constants: Dict[str, int]):
        self.calldata = calldata
        self.mem = bytearray(0x4000)
        self.constants = constants
        self.funcs: Dict[str, tuple] = {}

    # --- memory ---
    def _grow(self, end):
        if end > len(self.mem):
            self.mem.extend(b"\x00" * (end - len(self.mem) + 0x1000))

    def mload(self, p):
        self._grow(p + 32)
        return int.from_bytes(self.mem[p:p + 32], "big")

    def mstore(self, p, v):
        self._grow(p + 32)
        self.mem[p:p + 32] = (v & M256).to_bytes(32, "big")

    # --- precompiles ---
    def staticcall(self, addr, inp, insz, outp, outsz):
        self._grow(max(inp + insz, outp + outsz))
        data = bytes(self.mem[inp:inp + insz])
        word = lambda i: int.from_bytes(data[32 * i:32 * i + 32], "big")
        try:
            if addr == 5:  # modexp with 32-byte base/exp/mod
                bl, el, ml = word(0), word(1), word(2)
                assert (bl, el, ml) == (32, 32, 32)
                res = pow(word(3), word(4), word(5)).to_bytes(32, "big")
            elif addr == 6:
                p1, p2 = self._g1(word(0), word(1)), self._g1(word(2), word(3))
                res = self._enc(B.g1_add(p1, p2))
            elif addr == 7:
                res = self._enc(B.g1_mul(self._g1(word(0), word(1)), word(2)))
            elif addr == 8:
                pairs = []
                for k in range(insz // 192):
                    g1 = self._g1(word(6 * k), word(6 * k + 1))
                    xi, xr, yi, yr = word(6 * k + 2), word(6 * k + 3), word(6 * k + 4), word(6 * k + 5)
                    g2 = None if (xi | xr | yi | yr) == 0 else ((xr, xi), (yr, yi))
                    pairs.append((g1, g2))
                res = (1 if pairing.pairing_check(pairs) else 0).to_bytes(32, "big")
            else:
                return 0
        except (AssertionError, ValueError):
            return 0
        self.mem[outp:outp + outsz] = res[:outsz]
        return 1

    @staticmethod
    def _g1(x, y):
        if x == 0 and y == 0:
            return None
        assert x < B.Q and y < B.Q and B.g1_is_on_curve((x, y))
        return (x, y)

    @staticmethod
    def _enc(p):
        if p is None:
            return b"\x00" * 64
        return p[0].to_bytes(32, "big") + p[1].to_bytes(32, "big")

    # --- evaluation ---
    def call(self, name, args, scopes):
        a = [self.eval(x, scopes) for x in args]
        if name == "add": return (a[0] + a[1]) & M256
        if name == "sub": return (a[0] - a[1]) & M256
        if name == "mul": return (a[0] * a[1]) & M256
        if name == "mod": return a[0] % a[1] if a[1] else 0
        if name == "addmod": return (a[0] + a[1]) % a[2] if a[2] else 0
        if name == "mulmod": return (a[0] * a[1]) % a[2] if a[2] else 0
        if name == "lt": return int(a[0] < a[1])
        if name == "eq": return int(a[0] == a[1])
        if name == "and": return a[0] & a[1]
        if name == "iszero": return int(a[0] == 0)
        if name == "shl": return (a[1] << a[0]) & M256
        if name == "mload": return self.mload(a[0])
        if name == "mstore": self.mstore(a[0], a[1]); return None
        if name == "mstore8": self._grow(a[0] + 1); self.mem[a[0]] = a[1] & 0xFF; return None
        if name == "calldataload":
            chunk = self.calldata[a[0]:a[0] + 32]
            return int.from_bytes(chunk + b"\x00" * (32 - len(chunk)), "big")
        if name == "keccak256":
            self._grow(a[0] + a[1])
            return int.from_bytes(keccak256(bytes(self.mem[a[0]:a[0] + a[1]])), "big")
        if name == "gas": return 1 << 60
        if name == "pop": return None
        if name == "staticcall": return self.staticcall(a[1], a[2], a[3], a[4], a[5])
        if name == "revert": raise Revert()
        if name == "return":
            self._grow(a[0] + a[1])
            raise Return(bytes(self.mem[a[0]:a[0] + a[1]]))
        fn = self.funcs.get(name)
        if fn is None:
            raise NameError(f"yul: unknown function {name}")
        _, params, rets, body = fn
        scope = dict(zip(params, a))
        for r in rets:
            scope[r] = 0
        self.exec_block(body, [scope])
        vals = [scope[r] for r in rets]
        return vals[0] if len(vals) == 1 else (tuple(vals) if vals else None)

    def eval(self, e, scopes):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "var":
            for s in reversed(scopes):
                if e[1] in s:
                    return s[e[1]]
            if e[1] in self.constants:
                return self.constants[e[1]]
            if e[1] == "true":
                return 1
            if e[1] == "false":
                return 0
            raise NameError(f"yul: unknown identifier {e[1]}")
        return self.call(e[1], e[2], scopes)

    def assign(self, names, val, scopes, declare):
        vals = list(val) if isinstance(val, tuple) else [val]
        if len(vals) != len(names):
            raise ValueError("yul: arity mismatch in assignment")
        for n, v in zip(names, vals):
            if declare:
                scopes[-1][n] = v
            else:
                for s in reversed(scopes):
                    if n in s:
                        s[n] = v
                        break
                else:
                    raise NameError(f"yul: assignment to unknown {n}")

    def exec_block(self, blk, scopes, new_scope=True):
        if new_scope:
            scopes = scopes + [{}]
        for st in blk[1]:
            if st[0] == "function":
                self.funcs[st[1]] = (st[1], st[2], st[3], st[4])
        for st in blk[1]:
            self.exec(st, scopes)

    def exec(self, st, scopes):
        k = st[0]
        if k == "block":
            self.exec_block(st, scopes)
        elif k == "let":
            if st[2] is None:
                for n in st[1]:
                    scopes[-1][n] = 0
            else:
                self.assign(st[1], self.eval(st[2], scopes), scopes, True)
        elif k == "assign":
            self.assign(st[1], self.eval(st[2], scopes), scopes, False)
        elif k == "expr":
            self.eval(st[1], scopes)
        elif k == "if":
            if self.eval(st[1], scopes):
                self.exec_block(st[2], scopes)
        elif k == "for":
            loop_scopes = scopes + [{}]
            self.exec_block(st[1], loop_scopes, new_scope=False)
            while self.eval(st[2], loop_scopes):
                self.exec_block(st[4], loop_scopes)
                self.exec_block(st[3], loop_scopes)
        elif k == "function":
            pass
        else:
            raise SyntaxError(f"yul: unknown statement {k}")


class SolidityVerifier:
    """`verifyProof(bytes proof, uint256[] instances)` of a halo2-solidity-verifier contract."""

    def __init__(self, sol_text: str):
        self.sol_text = sol_text
        self.constants = {m.group(1): int(m.group(2), 0)
                          for m in re.finditer(r"uint256 internal constant\s+(\w+)\s*=\s*(0x[0-9a-fA-F]+|\d+);", sol_text)}
        start = sol_text.index("assembly {") + len("assembly ")
        depth, i = 0, start
        while True:
            ch = sol_text[i]
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
                if depth == 0:
                    break
            i += 1
        self.ast = Parser(tokenize(sol_text[start:i + 1])).block()

    @classmethod
    def from_file(cls, path: str) -> "SolidityVerifier":
        with open(path) as f:
            return cls(f.read())

    def patched(self, replacements: Dict[str, int]) -> "SolidityVerifier":
        """New verifier with `// name` vk constants replaced (e.g. {'k': 13, 'omega': ..., 'fixed_comms[0].x': ...})."""
        text = self.sol_text
        for name, val in replacements.items():
            pat = re.compile(r"(mstore\(0x[0-9a-f]+, )0x[0-9a-f]{64}(\) // " + re.escape(name) + r")\n")
            text, n = pat.subn(lambda m: f"{m.group(1)}0x{val:064x}{m.group(2)}\n", text)
            if n != 1:
                raise KeyError(f"vk constant {name!r} not found exactly once")
        return SolidityVerifier(text)

    @staticmethod
    def encode_calldata(proof: bytes, instances: List[int]) -> bytes:
        """ABI encoding of verifyProof(bytes,uint256[]) (selector bytes are irrelevant to the assembly)."""
        head = b"\x00" * 4
        proof_off = 0x40
        padded = proof + b"\x00" * (-len(proof) % 32)
        inst_off = proof_off + 32 + len(padded)
        body = proof_off.to_bytes(32, "big") + inst_off.to_bytes(32, "big")
        body += len(proof).to_bytes(32, "big") + padded
        body += len(instances).to_bytes(32, "big") + b"".join(int(x).to_bytes(32, "big") for x in instances)
        return head + body

    def verify(self, proof: bytes, instances: List[int]) -> bool:
        evm = EVM(self.encode_calldata(proof, instances), self.constants)
        try:
            evm.exec_block(self.ast, [{}])
        except Revert:
            return False
        except Return as r:
            return int.from_bytes(r.data, "big") == 1
        return False
