/* libsumma_b200 -- C ABI of the B200-native halo2 (KZG / BN254) prover core.
 *
 * Drop-in boundary for the hot path of summa-dev/circuits-halo2's `zk_prover`.  The reference has
 * no FFI seam of its own: `zk_prover/src/circuits/utils.rs:14-26` calls generic Rust functions of
 * the (un-vendored) crate halo2_proofs 0.2.0 @ summa-dev/halo2#8386d6e.  The seam is therefore a
 * Cargo `[patch]` of halo2_proofs whose function *bodies* call the entry points below
 * (INTEGRATION.md shows the Rust `extern "C"` block and the patched bodies).  Each entry point
 * names the Rust item it replaces.
 *
 * Conventions
 *  - Every function returns an int32 status (SB_OK == 0); nothing throws or unwinds across the ABI.
 *  - Field elements and points use halo2curves' in-memory layout, so Rust passes `slice.as_ptr()`:
 *      Fr / Fq   : 32 B, 4 x u64 little-endian limbs, Montgomery form (R = 2^256)
 *      G1Affine  : x || y (64 B), identity = (0, 0)
 *      G1        : x || y || z Jacobian (96 B), identity z = 0
 *  - `*_dev` variants take DEVICE pointers (cudaMalloc / torch storage) and a `cudaStream_t` passed
 *    as `void*` (NULL = the context's own stream).  Host variants copy in/out on the context stream.
 *  - Caller owns every buffer it passes; the library owns what hides behind its opaque handles.
 *  - All entry points are re-entrant for distinct contexts; calls on one context are serialised.
 */
#ifndef SUMMA_B200_H
#define SUMMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_CUDA 1       /* a CUDA runtime call failed (sb_last_error() has the text) */
#define SB_ERR_ARG 2        /* invalid argument (null pointer, size mismatch, k out of range) */
#define SB_ERR_NO_DEVICE 3  /* no CUDA device visible: there is NO CPU fallback */
#define SB_ERR_ALLOC 4      /* device or host allocation failed */

#define SB_BASIS_MONOMIAL 0 /* ParamsKZG::g          -> ParamsKZG::commit          */
#define SB_BASIS_LAGRANGE 1 /* ParamsKZG::g_lagrange -> ParamsKZG::commit_lagrange */

typedef struct sb_ctx sb_ctx;       /* one GPU + its stream, scratch arena and plan cache */
typedef struct sb_srs sb_srs;       /* device-resident KZG bases (ParamsKZG) */
typedef struct sb_domain sb_domain; /* device-resident EvaluationDomain constants + NTT plans */

/* ---- library / context ------------------------------------------------------------------ */
int32_t sb_version(void);
const char *sb_last_error(void); /* thread-local text of the last non-OK status */
int32_t sb_device_count(int32_t *out_count);
int32_t sb_ctx_create(int32_t device, sb_ctx **out_ctx);
int32_t sb_ctx_destroy(sb_ctx *ctx);
int32_t sb_ctx_synchronize(sb_ctx *ctx);
/* throughput mode: host waits of this context sleep on a blocking event instead of spinning (many worker contexts per host: BatchProver) */
int32_t sb_ctx_set_blocking_sync(sb_ctx *ctx, int32_t on);
/* the context's own CUDA stream (a cudaStream_t): callers that time with CUDA events record them here */
int32_t sb_ctx_stream(const sb_ctx *ctx, void **out_stream);
/* device memory helpers for callers without a CUDA runtime of their own (Rust FFI crate) */
int32_t sb_dev_alloc(sb_ctx *ctx, size_t bytes, void **out_dptr);
int32_t sb_dev_free(sb_ctx *ctx, void *dptr);
int32_t sb_dev_upload(sb_ctx *ctx, void *dst_dptr, const void *src_host, size_t bytes);
int32_t sb_dev_download(sb_ctx *ctx, void *dst_host, const void *src_dptr, size_t bytes);

/* ---- field vectors (halo2curves bn256::Fr; used by EvaluationDomain and the parity tests) -- */
/* op: 0 mul, 1 add, 2 sub (element-wise, n elements, host buffers) */
int32_t sb_fr_vec_op(sb_ctx *ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);
int32_t sb_fq_vec_op(sb_ctx *ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n);

/* ---- MSM: halo2_proofs::arithmetic::best_multiexp ------------------------------------------
 * Rust: `pub fn best_multiexp<C: CurveAffine>(coeffs: &[C::Scalar], bases: &[C]) -> C::Curve`
 * (reached from utils.rs:75-76,94-102,171-178 through ParamsKZG::{commit, commit_lagrange}). */
/* generic form, host buffers; result as Jacobian with z = 1 (or z = 0 for the identity) */
int32_t sb_best_multiexp(sb_ctx *ctx, const uint8_t *coeffs, const uint8_t *bases, size_t n, uint8_t out_jacobian[96]);
/* device-resident operands (bases: n x 64 B, scalars: n x 32 B); result affine on the host */
int32_t sb_msm_g1_dev(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, uint8_t out_affine[64], void *stream);

/* ---- SRS: halo2_proofs::poly::kzg::commitment::ParamsKZG (utils.rs:55,64,70) ---------------- */
/* g / g_lagrange: 2^k affine points each, exactly the arrays ParamsKZG holds */
int32_t sb_srs_upload(sb_ctx *ctx, uint32_t k, const uint8_t *g, const uint8_t *g_lagrange, sb_srs **out_srs);
/* `ParamsKZG::setup(k, rng)` (utils.rs:70) with an explicit secret tau (Montgomery Fr): UNSAFE test / benchmark SRS,
 * generated entirely on the device (powers of tau, Lagrange weights, 2 x 2^k fixed-base products) */
int32_t sb_srs_setup_unsafe(sb_ctx *ctx, uint32_t k, const uint8_t tau[32], sb_srs **out_srs);
int32_t sb_srs_download(sb_ctx *ctx, const sb_srs *srs, uint8_t *g_out, uint8_t *g_lagrange_out);
/* `ParamsKZG::downsize(new_k)` (utils.rs:62-66): the first 2^new_k monomial bases and their Lagrange bases, recomputed on the device by an
 * inverse FFT over group elements (halo2 `g_to_lagrange`).  Returns a new handle; new_k <= k. */
int32_t sb_srs_downsize(sb_ctx *ctx, const sb_srs *srs, uint32_t new_k, sb_srs **out_srs);
/* same handle over bases that already live on the device (borrowed: the caller keeps ownership) */
int32_t sb_srs_wrap_dev(sb_ctx *ctx, uint32_t k, const void *d_g, const void *d_g_lagrange, sb_srs **out_srs);
int32_t sb_srs_destroy(sb_srs *srs);
/* Fixed-base precomputation for ParamsKZG::{commit, commit_lagrange}: for every base P_i of the chosen bases
 * (basis_mask = bit set of 1 << SB_BASIS_*) store 2^(c w) * P_i for all windows w, so that every signed-digit window of a scalar
 * falls into ONE shared bucket set and the window can be 20-22 bits wide (12-13 mixed additions per point instead of 16).
 * window_bits 0 = choose from k.  Costs ceil(255 / c) x the memory of the bases (k = 20: 832 MiB per basis); results are
 * bit-identical with and without it.  sb_pk_create does this for its SRS unless the environment sets SB_NO_TABLES. */
int32_t sb_srs_precompute(sb_ctx *ctx, sb_srs *srs, int32_t basis_mask, uint32_t window_bits);
/* ParamsKZG::commit (basis 0) / commit_lagrange (basis 1): scalars host, n <= 2^k */
int32_t sb_msm_g1(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const uint8_t *scalars, size_t n, uint8_t out_affine[64]);
int32_t sb_msm_g1_srs_dev(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const void *d_scalars, size_t n, uint8_t out_affine[64], void *stream);

/* data-parallel core of ParamsKZG::setup (utils.rs:70: g[i] = [tau^i] G): out[i] = scalars[i] * G,
 * affine, device pointers (n x 32 B in, n x 64 B out).  Also synthesises bench / test bases. */
int32_t sb_g1_fixed_base_mul_dev(sb_ctx *ctx, const void *d_scalars, size_t n, void *d_out_affine, void *stream);

/* ---- NTT: halo2_proofs::arithmetic::best_fft ------------------------------------------------
 * Rust: `pub fn best_fft<Scalar: Field, G: FftGroup<Scalar>>(a: &mut [G], omega: Scalar, log_n: u32)`
 * natural order in, natural order out, in place: a[k] <- sum_j a[j] omega^(jk). */
int32_t sb_best_fft(sb_ctx *ctx, uint8_t *a, const uint8_t omega[32], uint32_t log_n);
int32_t sb_ntt_dev(sb_ctx *ctx, void *d_a, const uint8_t omega[32], uint32_t log_n, void *stream);

/* ---- halo2_proofs::poly::EvaluationDomain<Fr> -----------------------------------------------
 * `EvaluationDomain::new(j, k)`: j = cs.degree(); extended_k = k + ceil(log2(j - 1)). */
int32_t sb_domain_create(sb_ctx *ctx, uint32_t j, uint32_t k, sb_domain **out_domain);
int32_t sb_domain_destroy(sb_domain *domain);
int32_t sb_domain_extended_k(const sb_domain *domain, uint32_t *out);
/* host-buffer forms (in place unless two pointers are given) */
int32_t sb_lagrange_to_coeff(sb_ctx *ctx, const sb_domain *d, uint8_t *a /* 2^k */);
int32_t sb_coeff_to_lagrange(sb_ctx *ctx, const sb_domain *d, uint8_t *a /* 2^k */);
int32_t sb_coeff_to_extended(sb_ctx *ctx, const sb_domain *d, const uint8_t *coeff /* 2^k */, uint8_t *ext /* 2^ext_k */);
int32_t sb_extended_to_coeff(sb_ctx *ctx, const sb_domain *d, const uint8_t *ext /* 2^ext_k */, uint8_t *coeff /* (j-1)*2^k */);
int32_t sb_divide_by_vanishing_poly(sb_ctx *ctx, const sb_domain *d, uint8_t *ext /* 2^ext_k */);
/* device-pointer forms */
int32_t sb_lagrange_to_coeff_dev(sb_ctx *ctx, const sb_domain *d, void *d_a, void *stream);
int32_t sb_coeff_to_lagrange_dev(sb_ctx *ctx, const sb_domain *d, void *d_a, void *stream);
int32_t sb_coeff_to_extended_dev(sb_ctx *ctx, const sb_domain *d, const void *d_coeff, void *d_ext, void *stream);
int32_t sb_extended_to_coeff_dev(sb_ctx *ctx, const sb_domain *d, void *d_ext /* clobbered */, void *d_coeff, void *stream);
int32_t sb_divide_by_vanishing_poly_dev(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *stream);

/* ---- halo2_proofs::plonk::{keygen_pk, create_proof} --------------------------------------------
 * Rust: `create_proof<KZGCommitmentScheme<Bn256>, ProverSHPLONK<'_, Bn256>, _, R: RngCore, T, C>(params, &pk,
 *        &[circuit], &[&[&[F]]], rng, &mut transcript)`  (utils.rs:94-102 Blake2bWrite, :171-178 Keccak256Transcript).
 * The constraint system travels as JSON (schema: tests/golden/mst_inclusion_cs.json -- gates / lookups as
 * expression trees over (advice|fixed|instance, column, rotation), permutation columns, query lists, degree,
 * blinding_factors); fixed_values / sigma_values are `pk.fixed_values` and `pk.permutation.permutations`
 * (Lagrange form, n x 32 B per column, column-major).  The handle keeps every per-circuit polynomial in HBM. */
typedef struct sb_pk sb_pk;
int32_t sb_pk_create(sb_ctx *ctx, const sb_srs *srs, const char *cs_json, uint32_t k, const uint8_t *fixed_values, const uint8_t *sigma_values,
                     const uint8_t transcript_repr[32] /* vk.transcript_repr, Montgomery */, sb_pk **out_pk);
/* same, from keygen's sparse output: the assigned fixed cells (col,row pairs + 32 B values; all other cells 0) and the
 * cells (col,row,to_col,to_row) that copy constraints moved away from the identity permutation.  Dense columns are built
 * on the device, so a k = 20..23 key costs no host memory. */
int32_t sb_pk_create_sparse(sb_ctx *ctx, const sb_srs *srs, const char *cs_json, uint32_t k, const uint32_t *fixed_cells, const uint8_t *fixed_cell_values,
                            size_t n_fixed, const uint32_t *perm_cells, size_t n_perm, const uint8_t transcript_repr[32], sb_pk **out_pk);
int32_t sb_pk_destroy(sb_pk *pk);
/* keygen_vk's commitments of the fixed and permutation columns (affine, 64 B each) */
int32_t sb_pk_commitments(const sb_pk *pk, uint8_t *fixed_comms, uint8_t *sigma_comms);
/* instances: n_instances x 32 B (Montgomery); advice: num_advice x n x 32 B assigned cells (rows >= n - blinding - 1
 * are overwritten by blinding); rng_seed: ChaCha20Rng::from_seed; transcript_kind 0 = Blake2b (full_prover),
 * 1 = Keccak256 / EVM (gen_proof_solidity_calldata).  Writes the proof bytes and their length. */
int32_t sb_create_proof(sb_ctx *ctx, const sb_pk *pk, const uint8_t *instances, size_t n_instances, const uint8_t *advice, const uint8_t rng_seed[32],
                        int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len);
/* The same proof from other forms of the witness (identical bytes for identical cells, seed and key):
 *  _dev:    advice already in device memory (A x n x 32 B, column-major; left untouched) -- a device witness generator, or bench.py's
 *           HBM-resident measurement;
 *  _sparse: only the ASSIGNED advice cells, n_cells x (col, row) u32 pairs + n_cells x 32 B values, all other cells zero.
 *           `MstInclusionCircuit` assigns a few thousand cells whatever k is, so a k = 23 proof uploads ~1 MB instead of 768 MiB. */
int32_t sb_create_proof_dev(sb_ctx *ctx, const sb_pk *pk, const uint8_t *instances, size_t n_instances, const void *d_advice, const uint8_t rng_seed[32],
                            int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len);
int32_t sb_create_proof_sparse(sb_ctx *ctx, const sb_pk *pk, const uint8_t *instances, size_t n_instances, const uint32_t *advice_cells, const uint8_t *advice_cell_values,
                               size_t n_cells, const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len);
/* ---- one proof sharded over the GPUs of a box (BASELINE configs[3]) --------------------------------
 * One process (or thread) per GPU, each with its own sb_ctx / sb_srs / sb_pk built from the same inputs.  Every rank calls
 * sb_create_proof_sharded with identical arguments; the ranks run the transcript in lock step and return the same proof bytes.
 * Sharded work: every commitment MSM by signed-digit WINDOW (rank r accumulates windows [r W / world, (r + 1) W / world) over all bases; the 128-byte
 * XYZZ partials meet in allgather_host and are added on the host -- a base-range split remains selectable with SB_SHARD_MSM_BY_RANGE),
 * the coset NTTs, evaluate_h and the division by t(X) by cosets of the extended domain (the quotient's values meet in
 * allgather_dev).  `world` must divide 2^(extended_k - k).  The callbacks are the host's collective library (NCCL, MPI, gloo):
 * both gather `bytes_per_rank` bytes from every rank in rank order and return 0 on success.  allgather_dev works in place on
 * device memory: rank r's part already sits at d_buf + r * bytes_per_rank; it must be complete (or stream-ordered on `stream`)
 * when it returns. */
typedef struct sb_comm {
    int32_t rank, world;
    void *user;
    int32_t (*allgather_host)(void *user, const void *send, void *recv, size_t bytes_per_rank);
    int32_t (*allgather_dev)(void *user, void *d_buf, size_t bytes_per_rank, void *stream);
    /* all-to-all of device memory: block q of d_send (bytes_per_pair each) goes to rank q, block q of d_recv comes from rank q; complete (or
     * stream-ordered on `stream`) on return.  Needed by sb_ntt_dist (and by sharded proofs at k >= SB_DIST_NTT_MIN_K); may be NULL otherwise. */
    int32_t (*alltoall_dev)(void *user, const void *d_send, void *d_recv, size_t bytes_per_pair, void *stream);
} sb_comm;
/* Distributed four-step NTT (north_star: "an NCCL all-to-all over NVLink ... for a distributed four-step NTT at the largest K"): best_fft of a
 * vector that every rank holds in full at d_a, in place; each rank computes 1 / world of both passes, an all-to-all transposes between them and an
 * all-gather replicates the result.  Sizes 2^16 .. 2^24 (two-pass plans); smaller sizes are transformed locally by every rank.  `scale` (or NULL)
 * multiplies the result (n^-1 of an inverse transform).  Every rank must call it with the same arguments. */
int32_t sb_ntt_dist(sb_ctx *ctx, const sb_comm *comm, void *d_a, const uint8_t omega[32], uint32_t log_n, const uint8_t *scale /* 32 B or NULL */, void *stream);
int32_t sb_create_proof_sharded(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const uint8_t *advice,
                                const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len);
int32_t sb_create_proof_sharded_dev(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const void *d_advice,
                                    const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len);
int32_t sb_create_proof_sharded_sparse(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const uint32_t *advice_cells,
                                       const uint8_t *advice_cell_values, size_t n_cells, const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out,
                                       size_t proof_cap, size_t *proof_len);
/* A ready-made sb_comm for the ranks of ONE box (one process per GPU) that needs no collective library: small host records meet in a POSIX
 * shared-memory mailbox (microseconds per exchange), device buffers move by direct peer copies over NVLink through CUDA IPC mappings.  `name` must be
 * the same on every rank and unique per job (a stale segment of the same name must not exist); `ctx` may be NULL for host-only use.  The call
 * returns when all `world` ranks have attached.  The filled sb_comm stays valid until sb_comm_shm_destroy. */
typedef struct sb_shm_comm sb_shm_comm;
int32_t sb_comm_shm_create(sb_ctx *ctx, const char *name, int32_t rank, int32_t world, sb_comm *out_comm, sb_shm_comm **out_handle);
int32_t sb_comm_shm_destroy(sb_shm_comm *handle);
/* device time (ms) of the fused evaluate_h kernel of the last create_proof on this context and its program shape:
 * instructions, field products, additions/subtractions, live value slots */
int32_t sb_last_h_profile(const sb_ctx *ctx, float *out_ms, uint32_t out_program[4]);
/* rows the fused program ran on in that proof on this rank: (owned cosets of the quotient argument) x 2^k */
int32_t sb_last_h_rows(const sb_ctx *ctx, uint64_t *out_rows);
/* 1 when that evaluate_h ran the key's program as NVRTC-compiled straight-line code, 0 when the interpreter ran it (NVRTC absent, or SB_NO_JIT) */
int32_t sb_last_h_jit(const sb_ctx *ctx, int32_t *out_used);
/* host wall-clock (ms) of the stages of the last create_proof: [0] advice upload + commitments, [1] lookup permute + commitments,
 * [2] permutation products, [3] lookup product, [4] random polynomial, [5] coset NTTs, [6] evaluate_h, [7] quotient + commitments,
 * [8] evaluations, [9] SHPLONK */
int32_t sb_last_proof_stages(const sb_ctx *ctx, float out_ms[12]);
/* the commitments (MSM launch sets) of the last create_proof on this context: summed device times of the phases listed at sb_msm_phase_times,
 * the number of signed digits accumulated (= level-1 mixed additions: sum over launch sets of windows x points x vectors) and of launch sets */
int32_t sb_last_proof_msm(const sb_ctx *ctx, float out_ms[5], uint64_t *out_digits, uint32_t *out_launch_sets);
/* device -> host bytes of the last proof's commitments (per commitment: the bucket reduction's 18 XYZZ records of 128 B, folded on the host) */
int32_t sb_last_proof_d2h(const sb_ctx *ctx, uint64_t *out_bytes);
/* ---- halo2_proofs::plonk::evaluation::Evaluator::evaluate_h as a standalone entry (the finer-grained Cargo patch: the body of evaluate_h) ----
 * The quotient NUMERATOR sum_i y^(T-1-i) term_i of the constraint system `cs_json` (gate polynomials, then the permutation argument's terms, then
 * every lookup's terms: SURVEY A.8) over caller-supplied columns, 2^log_rows values each on one common domain, in this order:
 *   advice (A) | fixed (F) | instance (1) | sigma (P) | permutation Z (ceil(P / (degree - 2))) | l_0, l_last, l_active, X | per lookup: Z, A', S'
 * A row rotation r reads (row + (r << rot_scale_log)) mod 2^log_rows: halo2's extended domain in natural order has rot_scale_log = extended_k - k;
 * one coset of the size-n subgroup has 0.  X is the column of evaluation points.  Division by t(X) is NOT included (sb_divide_by_vanishing_poly). */
int32_t sb_evaluate_h(sb_ctx *ctx, const char *cs_json, const uint8_t *const *columns, size_t n_columns, uint32_t log_rows, uint32_t rot_scale_log, const uint8_t theta[32],
                      const uint8_t beta[32], const uint8_t gamma[32], const uint8_t y[32], uint8_t *out);
int32_t sb_evaluate_h_dev(sb_ctx *ctx, const char *cs_json, const void *const *d_columns, size_t n_columns, uint32_t log_rows, uint32_t rot_scale_log, const uint8_t theta[32],
                          const uint8_t beta[32], const uint8_t gamma[32], const uint8_t y[32], void *d_out, void *stream);
/* ParamsKZG::commit / commit_lagrange for m polynomials of n coefficients at once (scalars: m x n x 32 B, contiguous; out: m x 64 B): one launch set
 * per 8 vectors over the fixed-base tables (create_proof commits its advice / lookup / product / quotient columns this way) */
int32_t sb_msm_g1_batch(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const uint8_t *scalars, size_t n, size_t m, uint8_t *out_affine);
int32_t sb_msm_g1_batch_dev(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const void *d_scalars, size_t n, size_t m, uint8_t *out_affine, void *stream);
/* the grand-product column of the permutation / lookup arguments (halo2 permutation::Argument::commit, lookup commit_product: SURVEY A.6 / A.7):
 * z[0] = init, z[i + 1] = z[i] * numerators[i] / denominators[i]; n_z <= n + 1 values written */
int32_t sb_grand_product(sb_ctx *ctx, const uint8_t *numerators, const uint8_t *denominators, size_t n, const uint8_t init[32], uint8_t *z, size_t n_z);
/* building blocks of create_proof with host buffers (halo2 arithmetic::{eval_polynomial, kate_division},
 * poly::batch_invert, the grand-product scan of permutation / lookup Z, lookup `permute_expression_pair`) */
int32_t sb_fr_batch_invert(sb_ctx *ctx, uint8_t *a, size_t n);                                   /* zeros stay zero */
int32_t sb_fr_running_product(sb_ctx *ctx, const uint8_t *a, size_t n_a, const uint8_t init[32], uint8_t *z, size_t n_z); /* z[0]=init, z[i]=z[i-1]*a[i-1] */
int32_t sb_fr_eval_polynomial(sb_ctx *ctx, const uint8_t *coeffs, size_t n, const uint8_t *points, size_t n_points, uint8_t *out);
int32_t sb_fr_sort(sb_ctx *ctx, uint8_t *a, size_t n);                                            /* ascending canonical value (Fr `Ord`) */
int32_t sb_lookup_permute(sb_ctx *ctx, const uint8_t *input, const uint8_t *table, size_t n, size_t usable, uint8_t *permuted_input, uint8_t *permuted_table);
int32_t sb_kate_division(sb_ctx *ctx, const uint8_t *a, uint32_t log_n, const uint8_t b[32], uint8_t *q /* 2^log_n slots, top one zero */);
/* host-side primitives of the transcript / RNG, exported for the CPU test-suite */
int32_t sb_test_keccak256(const uint8_t *data, size_t len, uint8_t out[32]);
int32_t sb_test_blake2b512(const uint8_t *data, size_t len, const uint8_t personal[16], uint8_t out[64]);
int32_t sb_test_chacha_fr(uint64_t seed_u64, uint32_t skip_bytes, uint32_t count, uint8_t *out);
/* evaluate_h compiler on the CPU: quotient-numerator program of `cs_json` (challenges from `seed`) run by the host interpreter on one
 * pseudo-random row vs a direct walk of the expression trees; out_shape = instructions, field products, add/sub, live value slots */
int32_t sb_test_h_program(const char *cs_json, uint64_t seed, uint8_t out_program_value[32], uint8_t out_direct_value[32], uint32_t out_shape[4]);
/* the CUDA source the NVRTC path compiles for `cs_json`'s quotient-numerator program (cap 0: size query) */
int32_t sb_test_h_jit_source(const char *cs_json, char *out, size_t cap, size_t *out_len);
/* Host tail of a table MSM, callable without a GPU (test hook; csrc/host_g1.cpp).  `fin`: 18 XYZZ records of 128 B as the bucket-tree kernels leave them
 * ([0] = sum of all buckets T, [1 + b] = S_b, [17] = the plain sum of the Q vector when a running-sum level ran first); out = fin[x_slot] + 2^shift *
 * sum_{b < n_bits} 2^b S_b, then -- for a bucket-residue shard (log_mod > 0) -- 2^log_mod * out - (2^log_mod - res - 1) * T; normalised to 64 B affine. */
int32_t sb_test_msm_host_tail(const uint8_t *fin, int32_t n_bits, int32_t shift, int32_t x_slot, int32_t log_mod, int32_t res, uint8_t out_affine[64]);
int32_t sb_test_host_fr(int32_t op, const uint8_t a[32], const uint8_t b[32], uint8_t out[32]);

/* ---- zk_prover::merkle_sum_tree: MerkleSumTree::from_entries / Tree::generate_proof ------------------
 * The tree is built and kept in HBM: Keccak-256 of the usernames (entry.rs:15-27), Poseidon leaf hashes
 * H(username, balances...) (node.rs:16-27,57-69), middle nodes H(sum balances..., hash_l, hash_r) (node.rs:32-45,73-84),
 * one launch per level (utils/build_tree.rs:5-78).  Entries are padded with zero entries to 2^depth, depth = ceil(log2 n)
 * (mst.rs:106-114).  All field elements cross the ABI as 32 B Montgomery Fr. */
typedef struct sb_mst sb_mst;
/* usernames: concatenated UTF-8 bytes, entry i = usernames[offsets[i] .. offsets[i+1]); balances: n_entries x n_currencies u64 (N_BYTES <= 8) */
int32_t sb_mst_build(sb_ctx *ctx, const uint8_t *usernames, const uint32_t *offsets, const uint64_t *balances, size_t n_entries, uint32_t n_currencies,
                     sb_mst **out_mst);
/* the same with BigUint balances (entry.rs:10: `balances: [BigUint; N_CURRENCIES]`; merkle_sum_tree/tests.rs:130, csv/entry_16_bigints.csv hold values
 * >= 2^64): n_entries x n_currencies x 32 B little-endian integers, reduced mod r like `big_uint_to_fp` (operation_helpers.rs:10-12) */
int32_t sb_mst_build_wide(sb_ctx *ctx, const uint8_t *usernames, const uint32_t *offsets, const uint8_t *balances_le32, size_t n_entries, uint32_t n_currencies,
                          sb_mst **out_mst);
/* build_merkle_tree_from_leaves over Node::leaf_node_from_preimage: n_leaves (a power of two) x (n_currencies + 1) x 32 B: [username, balances...] */
int32_t sb_mst_build_from_preimages(sb_ctx *ctx, const uint8_t *leaf_preimages, size_t n_leaves, uint32_t n_currencies, sb_mst **out_mst);
int32_t sb_mst_destroy(sb_mst *mst);
/* depth, N_CURRENCIES and the device time (ms) the build took (H2D of the entries + all kernels) */
int32_t sb_mst_shape(const sb_mst *mst, uint32_t *out_depth, uint32_t *out_n_currencies, float *out_build_ms);
int32_t sb_mst_root(const sb_mst *mst, uint8_t out_hash[32], uint8_t *out_balances /* n_currencies x 32 B */);
int32_t sb_mst_node(const sb_mst *mst, uint32_t level, size_t index, uint8_t out_hash[32], uint8_t *out_balances); /* Tree::nodes()[level][index] */
int32_t sb_mst_level_hashes(const sb_mst *mst, uint32_t level, uint8_t *out_hashes /* 2^(depth-level) x 32 B */);
/* Tree::generate_proof (tree.rs:85-137) for n_proofs user indices at once.  Per proof, out_preimages holds
 * entry preimage (n_cur+1) | sibling leaf preimage (n_cur+1) | (depth-1) x sibling middle-node preimage (n_cur+2)  field elements,
 * out_path_indices holds depth bytes (0 = the node is a left child). */
int32_t sb_mst_proofs(const sb_mst *mst, const uint64_t *indices, size_t n_proofs, uint8_t *out_preimages, uint8_t *out_path_indices);
/* MerkleSumTree::update_leaf (mst.rs:158-197): new balances (n_currencies u64) for the entry at `index`, the path to the root is rehashed;
 * optionally returns the new root */
int32_t sb_mst_update_leaf(sb_mst *mst, size_t index, const uint64_t *new_balances, uint8_t out_root_hash[32], uint8_t *out_root_balances);
int32_t sb_mst_update_leaf_wide(sb_mst *mst, size_t index, const uint8_t *new_balances_le32 /* n_currencies x 32 B LE */, uint8_t out_root_hash[32], uint8_t *out_root_balances);
/* Tree::verify_proof (tree.rs:139-186) for n_proofs proofs in the layout sb_mst_proofs writes, against one root; out_ok[j] = 1 iff proof j
 * hashes up to root_hash with balances equal to root_balances */
int32_t sb_mst_verify_proofs(sb_ctx *ctx, uint32_t n_currencies, uint32_t depth, const uint8_t *preimages, const uint8_t *path_indices, const uint8_t root_hash[32],
                             const uint8_t *root_balances, size_t n_proofs, uint8_t *out_ok);

/* ---- zk_prover::circuits::merkle_sum_tree::MstInclusionCircuit: witness generation (host) ------------------------------------------------------
 * The advice cells `Circuit::synthesize` assigns (circuits/merkle_sum_tree.rs:228-520) for one Merkle proof in the layout sb_mst_proofs writes, placed
 * like halo2's SimpleFloorPlanner places the circuit's regions; plus the circuit's instances [leaf hash, root hash, root balances...]
 * (merkle_sum_tree.rs:54-58).  Output is sparse: (column, row) u32 pairs + 32 B values of the NON-ZERO cells, ready for sb_create_proof_sparse.
 * cap_cells = 0 only reports the number of cells.  Errors: the circuit does not fit 2^k rows; a balance exceeds N_BYTES bytes. */
int32_t sb_mst_inclusion_witness(uint32_t levels, uint32_t n_currencies, uint32_t n_bytes, uint32_t k, const uint8_t *preimages, const uint8_t *path_indices,
                                 uint32_t *out_cells, uint8_t *out_values, size_t cap_cells, size_t *out_n_cells, uint8_t *out_instances);

/* ---- instrumentation ------------------------------------------------------------------------ */
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int32_t sb_launch_count(const sb_ctx *ctx, uint64_t *out);
/* device time (ms, CUDA events on the launching stream) of the phases of the LAST MSM on this
 * context: [0] recode + counting sort, [1] reduce level 1 (dominant kernel), [2] reduce levels >= 2,
 * [3] bucket reduction, [4] whole device part.  out_shape (optional): c, windows, L1, seg_log. */
int32_t sb_msm_phase_times(const sb_ctx *ctx, float out_ms[5], uint32_t out_shape[4]);
/* throughput micro-kernel used to MEASURE the integer roof: each of n threads runs `iters`
 * dependent-free Montgomery products; returns the elapsed milliseconds (CUDA events). */
int32_t sb_bench_field_mul(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, int32_t field /*0 Fr,1 Fq*/, float *out_ms);
int32_t sb_bench_imad(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, float *out_ms);      /* 8 IMAD chains / thread */
int32_t sb_bench_imad_wide(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, float *out_ms); /* 8 IMAD.WIDE chains / thread */
int32_t sb_bench_imad_hi(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, float *out_ms);   /* 8 IMAD.HI chains / thread */

/* ---- host helper: sum of n affine points (folding the per-GPU partial MSM results; the
 * north_star's "partial G1 sums are reduced on the host") ---------------------------------- */
int32_t sb_g1_sum_affine(const uint8_t *pts, size_t n, uint8_t out_affine[64]);

#ifdef __cplusplus
}
#endif
#endif /* SUMMA_B200_H */
