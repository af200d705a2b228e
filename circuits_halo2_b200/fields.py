"""BN254 scalar-field constants the host side needs (halo2curves bn256::Fr).

Pure-python integers for *constants only* (roots of unity, Montgomery encoding of a handful of
values); all vector arithmetic runs on the GPU through the C ABI."""
import numpy as np

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_S = 28
ROOT_OF_UNITY = 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C  # 7^((r-1)/2^28)
MONT_R = 1 << 256


def omega(k: int) -> int:
    """Generator of the 2^k-th roots of unity: ROOT_OF_UNITY^(2^(28-k)) (EvaluationDomain::new)."""
    if not 0 <= k <= FR_S:
        raise ValueError("k out of range for the 2-adicity of Fr")
    w = ROOT_OF_UNITY
    for _ in range(FR_S - k):
        w = w * w % FR_MODULUS
    return w


def fr_to_mont(x: int) -> np.ndarray:
    """(4,) uint64 Montgomery limbs of x, the layout `&Fr` has in Rust."""
    return np.frombuffer(((x % FR_MODULUS) * MONT_R % FR_MODULUS).to_bytes(32, "little"), dtype=np.uint64).copy()


def fr_from_mont(limbs) -> int:
    v = int.from_bytes(np.ascontiguousarray(limbs).tobytes(), "little")
    return v * pow(MONT_R, -1, FR_MODULUS) % FR_MODULUS
