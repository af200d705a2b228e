"""circuits_halo2_b200 -- B200-native core of Summa's halo2 (KZG / BN254) prover.

Host-side mirror of the reference's call surface for the hot path, over the C ABI of
libsumma_b200.so (hand-written sm_100a kernels):

    arithmetic.best_multiexp / best_fft          halo2_proofs::arithmetic
    domain.EvaluationDomain                      halo2_proofs::poly::EvaluationDomain
    params.ParamsKZG                             halo2_proofs::poly::kzg::commitment::ParamsKZG
    plonk.ProvingKey / create_proof              halo2_proofs::plonk::{ProvingKey, create_proof}
    merkle_sum_tree.MerkleSumTree                zk_prover::merkle_sum_tree::MerkleSumTree

Arrays are numpy views of halo2curves' memory layout (Montgomery, little-endian u64 limbs):
Fr vectors have shape (n, 4) uint64, G1Affine vectors (n, 8) uint64.
"""
from .context import Context, default_context  # noqa: F401
from .arithmetic import best_fft, best_fft_dist, best_multiexp  # noqa: F401
from .domain import EvaluationDomain  # noqa: F401
from .params import ParamsKZG  # noqa: F401
from .merkle_sum_tree import Entry, MerkleProof, MerkleSumTree, Node  # noqa: F401
from .plonk import mst_inclusion_witness, BatchProver, LocalComm, ShardComm, ShmComm, shard_range, ProvingKey, create_proof, create_proof_dev, create_proof_sparse, seed_from_u64, TRANSCRIPT_BLAKE2B, TRANSCRIPT_KECCAK  # noqa: F401

__all__ = ["Context", "default_context", "best_fft", "best_multiexp", "EvaluationDomain", "ParamsKZG", "ProvingKey", "create_proof", "create_proof_sparse", "create_proof_dev", "seed_from_u64", "MerkleSumTree", "Entry", "Node", "MerkleProof"]
