// Shared host-side plumbing of libsumma_b200: status codes, the context object, scratch arena.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/summa_b200.h"
#include "fp.cuh"

namespace sb {

void set_last_error(const char *fmt, ...);

#define SB_CUDA_TRY(expr)                                                                            \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            ::sb::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return SB_ERR_CUDA;                                                                      \
        }                                                                                            \
    } while (0)

#define SB_TRY(expr)                                                                                 \
    do {                                                                                             \
        int32_t s__ = (expr);                                                                        \
        if (s__ != SB_OK) return s__;                                                                \
    } while (0)

#define SB_REQUIRE(cond, msg)                                                                        \
    do {                                                                                             \
        if (!(cond)) {                                                                               \
            ::sb::set_last_error("%s:%d: %s", __FILE__, __LINE__, msg);                              \
            return SB_ERR_ARG;                                                                       \
        }                                                                                            \
    } while (0)

// Grow-only device scratch buffers, one per named slot.  cudaMalloc/cudaFree are off the hot path:
// a slot is re-allocated only when a larger request arrives.
struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};

struct NttPlan;  // ntt.cu

// Developer knobs (SB_* environment variables), read ONCE by sb_ctx_create: nothing on a per-call path touches the environment.
struct Tuning {
    int tab_c = 0;            // SB_TAB_C: window bits of the fixed-base tables (0 = choose from k)
    int msm_lk = 8;           // SB_MSM_LK: chunk length of MSM reduce levels >= 2
    int msm_c = 0;            // SB_MSM_C: window bits of the table-free MSM (0 = choose from n)
    int msm_l1 = 0;           // SB_MSM_L1: chunk length of reduce level 1 (0 = choose)
    int msm_seg = -1;         // SB_MSM_SEG: log2 segment of the first bucket level (-1 = choose)
    int msm_seg1 = 2;         // SB_MSM_SEG1: log2 segment of later bucket levels
    int msm_cta_scan_max = 65536;  // SB_MSM_CTA_SCAN_MAX: reduce levels with at most this many slots use one CTA-wide segmented scan per 256 slots
    int msm_finish_at = 16384;  // SB_MSM_FINISH_AT: bucket count below which the hierarchy finishes in one step
    int ntt_tile = 0;         // SB_NTT_TILE: 11 | 12 = log2 elements per NTT tile (0 = choose from the size)
    int ntt_passes = 0;       // SB_NTT_PASSES: minimum number of NTT passes (0 = as few as the tile allows)
    int dist_ntt_min_k = 22;  // SB_DIST_NTT_MIN_K: sharded proofs run their replicated size-n transforms as distributed four-step NTTs from this k on
    int ntt_eb = 3;           // SB_NTT_EB: 3 = eight elements per thread (radix-8 stages), 2 = four (radix-4 stages, twice the warps)
    int ntt_tw_mb = 1024;     // SB_NTT_TW_MB: budget (MiB) of a plan's full inter-pass twiddle tables; above it the two-level tables are used
    bool msm_no_cta_scan = false;    // SB_MSM_NO_CTA_SCAN
    bool shard_msm_by_range = false; // SB_SHARD_MSM_BY_RANGE
    bool no_side_stream = false;     // SB_NO_SIDE_STREAM
    bool no_early_random = false;    // SB_NO_EARLY_RANDOM
    bool no_hprog_cache = false;     // SB_NO_HPROG_CACHE
    bool msm_no_bucket_tree = false; // SB_MSM_NO_BUCKET_TREE: bucket reduction by the running-sum hierarchy instead of the tree kernels
    bool no_shplonk_lagrange = false; // SB_NO_SHPLONK_LAGRANGE: SHPLONK through the division coset (8 transforms) instead of the evaluation domain (1)
    bool no_grand_shard = false;     // SB_NO_GRAND_SHARD / SB_GRAND_SHARD: never / always shard the grand products of a sharded proof by rows (default: by measured rule)
    bool grand_shard = false;
    bool no_shplonk_shard = false;   // SB_NO_SHPLONK_SHARD: sharded proofs compute SHPLONK's evaluation-domain vectors in full on every rank
    bool shard_msm_by_window = false; // SB_SHARD_MSM_BY_WINDOW / SB_SHARD_MSM_BY_RESIDUE: force the split of sharded table commitments (default: by measured rule)
    bool shard_msm_by_residue = false;
    bool no_inst_direct = false;     // SB_NO_INST_DIRECT: instance column to the cosets through transforms even when it holds few values
    bool no_binv2 = false;           // SB_NO_BINV2: batch inversion of long vectors through the single-level kernel
    bool no_jit = false;             // SB_NO_JIT: evaluate_h through the interpreter instead of the NVRTC-specialised kernel
    bool no_tables = false;          // SB_NO_TABLES
    bool no_smallkey_sort = false;   // SB_NO_SMALLKEY_SORT
};

}  // namespace sb

struct sb_ctx {
    // Lifetime: handles that keep a pointer to their context (sb_mst, sb_pk) hold a reference; sb_ctx_destroy drops the caller's
    // reference and the context (stream, scratch arena, pinned buffers) is torn down when the LAST holder lets go, so the
    // destruction order of a context and its children is free (a Rust `Drop` order, Python GC order).
    std::atomic<int> refs{1};
    sb::Tuning tune;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;  // create_proof: coset NTTs that no challenge is waiting for run here, under the MSM tails
    cudaEvent_t side_ev[2] = {nullptr, nullptr};
    // throughput mode (sb_ctx_set_blocking_sync): host waits sleep on a blocking event instead of spinning in cudaStreamSynchronize, so that many
    // worker contexts (BatchProver: several per GPU, one process per GPU) do not burn the host cores their own transcript / witness work needs
    bool blocking_sync = false;
    cudaEvent_t block_ev = nullptr;
    std::mutex mu;
    uint64_t launches = 0;
    std::map<std::string, sb::Scratch> scratch;
    std::map<std::string, sb::NttPlan *> ntt_plans;  // key: log_n || omega bytes
    void *pinned = nullptr;                          // small pinned staging buffer (results)
    size_t pinned_bytes = 0;
    // pinned upload ring: small host -> device copies (blinding rows, compiled programs) go through it so the caller's pageable
    // buffer may die at once and no stream synchronisation is needed; a slot is reused only after its copy's event has fired
    static const int STAGE_SLOTS = 16;
    static const size_t STAGE_SLOT_BYTES = 64 << 10;
    uint8_t *stage = nullptr;
    cudaEvent_t stage_ev[STAGE_SLOTS] = {};
    uint32_t stage_next = 0;
    // per-phase device times of the last MSM (CUDA events on the launching stream):
    // [0] recode+sort  [1] reduce level 1 (the dominant kernel)  [2] reduce levels >= 2
    // [3] bucket reduction  [4] whole device part
    cudaEvent_t msm_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float msm_phase_ms[5] = {0, 0, 0, 0, 0};
    uint32_t msm_last_shape[4] = {0, 0, 0, 0};  // c, W, L1, seg_log of the last MSM
    // running totals over MSM launch sets since the last sb_perf_reset (create_proof resets them at entry): the phase times above, the
    // signed digits (= level-1 mixed additions) sorted and accumulated, and the number of launch sets
    float acc_msm_ms[5] = {0, 0, 0, 0, 0};
    uint64_t acc_msm_digits = 0;
    uint32_t acc_msm_sets = 0;
    uint64_t acc_msm_d2h = 0;                      // bytes the commitments of the last proof copied back (bucket-tree records, digit counters)
    // device time of every NTT pass kernel since the last reset is NOT kept (it would need an event pair per pass); bench.py's ncu launch list has it
    // evaluate_h of the last create_proof: device time and program shape (instructions, products, add/sub, live slots)
    cudaEvent_t h_ev[2] = {nullptr, nullptr};
    float last_h_ms = 0;
    bool last_h_jit = false;   // the last evaluate_h ran the NVRTC-specialised kernel (false: the interpreter)
    uint64_t last_h_rows = 0;  // rows the fused program was evaluated on (owned cosets x n)
    uint32_t last_h_program[4] = {0, 0, 0, 0};
    float last_proof_stage_ms[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // prover.cu `mark()` stages
};

namespace sb {

int32_t scratch_get(sb_ctx *ctx, const char *slot, size_t bytes, void **out);
void ctx_read_env(sb_ctx *ctx);
cudaError_t sync_stream(sb_ctx *ctx, cudaStream_t st);  // cudaStreamSynchronize, or a sleeping wait in throughput mode
void ctx_retain(sb_ctx *ctx);
void ctx_release(sb_ctx *ctx);  // frees the context when the last reference goes

// asynchronous upload of a small host buffer through the context's pinned ring (falls back to a synchronous copy when it does not fit)
int32_t h2d_staged(sb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);

inline cudaStream_t pick_stream(sb_ctx *ctx, void *stream) { return stream ? (cudaStream_t)stream : ctx->stream; }

// launch accounting (gpu_launches in bench.py)
#define SB_LAUNCH(ctx, kernel, grid, block, smem, stream, ...)                                       \
    do {                                                                                             \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                  \
        (ctx)->launches++;                                                                           \
        SB_CUDA_TRY(cudaGetLastError());                                                             \
    } while (0)

// ---- entry points of the individual translation units (all take device pointers) ----
int32_t ntt_run(sb_ctx *ctx, void *d_a, const uint8_t omega[32], uint32_t log_n, cudaStream_t st);
// The same transform with the scalings EvaluationDomain wraps around best_fft folded in (no extra pass over the data), out of place if wanted:
//   in[i] is read as zero for i >= n_in, multiplied by pre_vec[i] (device vector) and by pre_pat[i % pre_m];
//   out[i] is multiplied by `scale`, by post_pat[i % post_m] and by post_vec[i] (device vector), and only i < n_out is stored.
struct NttFuse {
    const void *pre_vec = nullptr, *post_vec = nullptr;
    uint32_t pre_m = 0, post_m = 0;
    fr_t pre_pat[8], post_pat[8];
    bool has_scale = false;
    fr_t scale;
    uint64_t n_in = 0, n_out = 0;  // 0 = the full size
    // `batch` transforms in ONE launch set (grid.y): transform b reads d_in + b * src_stride (0 = all read the same input), writes
    // d_out + b * dst_stride (0 = densely packed), uses pre_vec + b * pre_stride and post_vec + b * post_stride (strides in elements)
    uint32_t batch = 0;
    uint64_t src_stride = 0, dst_stride = 0, pre_stride = 0, post_stride = 0;
};
int32_t ntt_run_fused(sb_ctx *ctx, const void *d_in, void *d_out, const uint8_t omega[32], uint32_t log_n, const NttFuse *fuse, cudaStream_t st);
int32_t ntt_run_dist(sb_ctx *ctx, const sb_comm *comm, void *d_a, const uint8_t omega[32], uint32_t log_n, const fr_t *scale, cudaStream_t st);
void ntt_make_plan(uint32_t log_n, uint32_t tile_log, uint32_t max_passes_hint, int *npass, uint32_t *radix);
void ntt_plans_free(sb_ctx *ctx);

int32_t msm_run(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, uint8_t out_affine[64], cudaStream_t st);

// fixed-base window tables of an SRS basis (msm.cu): tables[w * stride + i] = 2^(c w) * base_i, affine
struct MsmTables {
    void *d_tables = nullptr;
    uint32_t c = 0, W = 0;
    uint64_t stride = 0;
};
int32_t msm_tables_build(sb_ctx *ctx, const void *d_bases, size_t n, uint32_t c, MsmTables *out, cudaStream_t st, void *d_dst = nullptr);
int32_t msm_run_tables_batch_mixed(sb_ctx *ctx, const MsmTables *t0, const MsmTables *t1, const void *d_scalars, size_t n, uint32_t batch, const uint8_t *basis_of,
                                   uint8_t *out_affine, cudaStream_t st);
int32_t msm_run_tables(sb_ctx *ctx, const MsmTables *tabs, const void *d_scalars, size_t n, int32_t w_lo, int32_t w_hi, uint8_t *out, cudaStream_t st);
int32_t msm_run_tables_batch(sb_ctx *ctx, const MsmTables *tabs, const void *d_scalars, size_t n, uint32_t batch, uint8_t *out_affine, cudaStream_t st);
// piece_off (optional, <= 8 vectors): table-entry offset per vector inside one slab (mixed-basis batch)
int32_t msm_run_tables_batch_windows(sb_ctx *ctx, const MsmTables *tabs, const void *d_scalars, size_t n, uint32_t batch, int32_t w_lo, int32_t w_hi, uint8_t *out,
                                     cudaStream_t st, const uint32_t *piece_off = nullptr, uint32_t res = 0, uint32_t log_mod = 0);
// window-sharded MSM (multi-GPU): window bits / count for n points, the XYZZ sums of windows [w_lo, w_hi), and the host Horner fold
void msm_window_shape(sb_ctx *ctx, size_t n, uint32_t *c, uint32_t *W);
int32_t msm_run_windows(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, uint32_t w_lo, uint32_t w_hi, uint8_t *win_out, cudaStream_t st);
void msm_fold_windows(const uint8_t *win, uint32_t W, uint32_t c, uint8_t out_affine[64]);

int32_t fr_gen_powers(sb_ctx *ctx, void *d_out, const fr_t &base, size_t count, cudaStream_t st);
// group-element inverse FFT (ParamsKZG::downsize / g_to_lagrange)
int32_t g1_to_lagrange(sb_ctx *ctx, const void *d_g, uint32_t log_n, const fr_t &omega_inv, const fr_t &n_inv, void *d_lagrange_out, cudaStream_t st);
int32_t g1_fixed_base_mul(sb_ctx *ctx, const void *d_scalars, size_t n, void *d_out, cudaStream_t st);

int32_t fr_scale(sb_ctx *ctx, void *d_a, size_t n, const fr_t &s, cudaStream_t st);
// a[i] *= pat[i % m]  (m <= 8 constants passed by value)
int32_t fr_scale_pattern(sb_ctx *ctx, void *d_a, size_t n, const fr_t *pat, uint32_t m, cudaStream_t st);
// dst[i] = i < n_src ? src[i] * pat[i % m] : 0   for i < n_dst
int32_t fr_scale_pattern_pad(sb_ctx *ctx, const void *d_src, size_t n_src, void *d_dst, size_t n_dst, const fr_t *pat, uint32_t m, cudaStream_t st);
// d_out[(r << shift) + j] = d_in[j * 2^log_n + r] * pat[j]
int32_t fr_coset_interleave_scale(sb_ctx *ctx, const void *d_in, void *d_out, uint32_t log_n, uint32_t shift, const fr_t *pat, cudaStream_t st);
int32_t fp_vec_op(sb_ctx *ctx, int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n, cudaStream_t st);

// host-side field helpers (exact, slow; plan constants only)
fr_t fr_pow_host(const fr_t &base, uint64_t e);
fr_t fr_from_u64_host(uint64_t x);  // to Montgomery

}  // namespace sb
