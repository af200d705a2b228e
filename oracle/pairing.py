"""CPU ORACLE (test infrastructure, NOT the product) -- BN254 optimal-ate pairing check in python ints.

Used only to execute the EVM precompile 0x08 (ecPairing) inside oracle/yul.py when the reference's
verifier contract (contracts/src/InclusionVerifier.sol:187-206,1393-1400) is interpreted.
Textbook construction (EIP-197 semantics): Fq12 = Fq[w]/(w^12 - 18 w^6 + 82), G2 twisted into Fq12,
Miller loop over the ate count 29793968203157093288, final exponentiation (q^12 - 1)/r.
Pinned by tests/test_verifier_golden.py: e(P, Q)^(ab) bilinearity and the reference's golden proof.
"""
from __future__ import annotations

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE_LOOP_COUNT = 63
FQ12_MOD = [82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0]  # w^12 = 18 w^6 - 82


class FQP:
    """Element of Fq[w]/(modulus), coefficients little-endian."""
    __slots__ = ("c",)
    deg = 12
    mod = FQ12_MOD

    def __init__(self, c):
        self.c = [x % Q for x in c]

    @classmethod
    def one(cls):
        return cls([1] + [0] * (cls.deg - 1))

    @classmethod
    def zero(cls):
        return cls([0] * cls.deg)

    def __add__(self, o):
        return type(self)([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return type(self)([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return type(self)([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k):
        return type(self)([a * k for a in self.c])

    def __mul__(self, o):
        if isinstance(o, int):
            return self.scale(o)
        d = self.deg
        b = [0] * (2 * d - 1)
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        for exp in range(2 * d - 2, d - 1, -1):
            top = b[exp] % Q
            if top:
                b[exp] = 0
                for i, m in enumerate(self.mod):
                    if m:
                        b[exp - d + i] -= top * m
        return type(self)(b[:d])

    def inv(self):
        # extended Euclid over Fq[w]
        d = self.deg
        lm, hm = [1] + [0] * d, [0] * (d + 1)
        low, high = self.c + [0], [m % Q for m in self.mod] + [1]

        def deg(p):
            k = len(p) - 1
            while k and p[k] == 0:
                k -= 1
            return k

        def poly_rounded_div(a, b):
            dega, degb = deg(a), deg(b)
            temp = list(a)
            o = [0] * len(a)
            for i in range(dega - degb, -1, -1):
                qv = temp[degb + i] * pow(b[degb], -1, Q) % Q
                o[i] = (o[i] + qv) % Q
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - qv * b[c]) % Q
            return o[: deg(o) + 1]

        while deg(low):
            r = poly_rounded_div(high, low)
            r += [0] * (d + 1 - len(r))
            nm, new = list(hm), list(high)
            for i in range(d + 1):
                for j in range(d + 1 - i):
                    nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
                    new[i + j] = (new[i + j] - low[i] * r[j]) % Q
            lm, low, hm, high = nm, new, lm, low
        k = pow(low[0], -1, Q)
        return type(self)([x * k for x in lm[:d]])

    def __pow__(self, e):
        out, base = type(self).one(), self
        while e:
            if e & 1:
                out = out * base
            base = base * base
            e >>= 1
        return out


def fq12(c):
    return FQP(list(c) + [0] * (12 - len(c)))


# --- generic affine arithmetic over FQP (None = identity) -----------------------------------
def _double(p):
    x, y = p
    lam = (x * x).scale(3) * (y.scale(2)).inv()
    nx = lam * lam - x.scale(2)
    ny = lam * (x - nx) - y
    return (nx, ny)


def _add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if y1 == y2:
            return _double(p1)
        return None
    lam = (y2 - y1) * (x2 - x1).inv()
    nx = lam * lam - x1 - x2
    ny = lam * (x1 - nx) - y1
    return (nx, ny)


def _linefunc(p1, p2, t):
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not x1 == x2:
        m = (y2 - y1) * (x2 - x1).inv()
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = (x1 * x1).scale(3) * (y1.scale(2)).inv()
        return m * (xt - x1) - (yt - y1)
    return xt - x1


W = fq12([0, 1])
W2, W3 = W * W, W * W * W


def twist(pt):
    """G2 point ((x_re, x_im), (y_re, y_im)) over Fq2 = Fq[i]/(i^2+1)  ->  curve over Fq12."""
    (xr, xi), (yr, yi) = pt
    xc = [xr - xi * 9, xi]
    yc = [yr - yi * 9, yi]
    nx = fq12([xc[0]] + [0] * 5 + [xc[1]])
    ny = fq12([yc[0]] + [0] * 5 + [yc[1]])
    return (nx * W2, ny * W3)


def cast_g1(pt):
    return (fq12([pt[0]]), fq12([pt[1]]))


def miller_loop(q2, p1):
    if q2 is None or p1 is None:
        return FQP.one()
    r_pt = q2
    f = FQP.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(r_pt, r_pt, p1)
        r_pt = _double(r_pt)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(r_pt, q2, p1)
            r_pt = _add(r_pt, q2)
    q1 = (q2[0] ** Q, q2[1] ** Q)
    nq2 = (q1[0] ** Q, -(q1[1] ** Q))
    f = f * _linefunc(r_pt, q1, p1)
    r_pt = _add(r_pt, q1)
    f = f * _linefunc(r_pt, nq2, p1)
    return f


def final_exponentiate(f):
    return f ** ((Q ** 12 - 1) // R)


def pairing_check(pairs):
    """EIP-197: prod e(g1_i, g2_i) == 1.  g1 = (x, y) or None; g2 = ((x_re, x_im), (y_re, y_im)) or None."""
    acc = FQP.one()
    for g1, g2 in pairs:
        if g1 is None or g2 is None:
            continue
        acc = acc * miller_loop(twist(g2), cast_g1(g1))
    return final_exponentiate(acc) == FQP.one()


G2_GEN = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
           11559732032986387107991004021392285783925812861821192530917403151452391805634),
          (8495653923123431417604973247489272438418190587263600148770280649306958101930,
           4082367875863433681332203403145435568316851327593401208105741076214120093531))
