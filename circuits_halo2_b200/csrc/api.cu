// C ABI of libsumma_b200 (include/summa_b200.h): contexts, SRS / domain handles, host-buffer
// wrappers around the device entry points, and the integer-roof micro-benchmarks.
#include <stdarg.h>

#include "common.cuh"

namespace sb {

void host_sum_affine(const uint8_t *pts, int n, uint8_t out_affine[64]);
void host_fold_windows(const uint8_t *win, int n_windows, int c, uint8_t out_affine[64]);
void host_bucket_combine(const uint8_t *fin, int n_bits, int shift, int x_slot, uint8_t out_xyzz[128]);
void host_residue_fixup(uint8_t r_xyzz[128], const uint8_t total_xyzz[128], int log_mod, int res);

static thread_local char g_err[512] = "";
void set_last_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int32_t h2d_staged(sb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return SB_OK;
    if (bytes > sb_ctx::STAGE_SLOT_BYTES || !ctx->stage) {
        SB_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
        return SB_OK;
    }
    const uint32_t slot = ctx->stage_next++ % sb_ctx::STAGE_SLOTS;
    SB_CUDA_TRY(cudaEventSynchronize(ctx->stage_ev[slot]));  // the copy that last used this slot (a never-recorded event is complete)
    uint8_t *h = ctx->stage + (size_t)slot * sb_ctx::STAGE_SLOT_BYTES;
    memcpy(h, h_src, bytes);
    SB_CUDA_TRY(cudaMemcpyAsync(d_dst, h, bytes, cudaMemcpyHostToDevice, st));
    SB_CUDA_TRY(cudaEventRecord(ctx->stage_ev[slot], st));
    return SB_OK;
}

int32_t scratch_get(sb_ctx *ctx, const char *slot, size_t bytes, void **out) {
    Scratch &s = ctx->scratch[slot];
    if (s.bytes < bytes) {
        if (s.ptr) {
            // in-flight kernels on the context stream may still use the old buffer
            cudaError_t e0 = cudaStreamSynchronize(ctx->stream);
            if (e0 != cudaSuccess) { set_last_error("scratch sync: %s", cudaGetErrorString(e0)); return SB_ERR_CUDA; }
            cudaDeviceSynchronize();
            cudaFree(s.ptr);
            s.ptr = nullptr;
            s.bytes = 0;
        }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&s.ptr, want);
        if (e != cudaSuccess) {
            set_last_error("scratch '%s': cudaMalloc(%zu) failed: %s", slot, want, cudaGetErrorString(e));
            s.ptr = nullptr;
            return SB_ERR_ALLOC;
        }
        s.bytes = want;
    }
    *out = s.ptr;
    return SB_OK;
}

void ctx_read_env(sb_ctx *ctx) {
    Tuning &t = ctx->tune;
    auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    auto getb = [](const char *name) { return getenv(name) != nullptr; };
    t.tab_c = geti("SB_TAB_C", 0);
    int lk = geti("SB_MSM_LK", 8);
    t.msm_lk = (lk >= 4 && lk <= 64 && lk % 4 == 0) ? lk : 8;
    t.msm_c = geti("SB_MSM_C", 0);
    t.msm_l1 = geti("SB_MSM_L1", 0);
    t.msm_seg = geti("SB_MSM_SEG", -1);
    t.msm_seg1 = geti("SB_MSM_SEG1", 2);
    t.msm_finish_at = geti("SB_MSM_FINISH_AT", 16384);
    t.msm_cta_scan_max = geti("SB_MSM_CTA_SCAN_MAX", 65536);
    ctx->blocking_sync = getb("SB_BLOCKING_SYNC");
    t.ntt_tile = geti("SB_NTT_TILE", 0);
    t.ntt_passes = geti("SB_NTT_PASSES", 0);
    t.ntt_tw_mb = geti("SB_NTT_TW_MB", 1024);
    t.ntt_eb = geti("SB_NTT_EB", 3);
    t.dist_ntt_min_k = geti("SB_DIST_NTT_MIN_K", 22);
    t.msm_no_cta_scan = getb("SB_MSM_NO_CTA_SCAN");
    t.shard_msm_by_range = getb("SB_SHARD_MSM_BY_RANGE");
    t.no_side_stream = getb("SB_NO_SIDE_STREAM");
    t.no_early_random = getb("SB_NO_EARLY_RANDOM");
    t.no_hprog_cache = getb("SB_NO_HPROG_CACHE");
    t.no_jit = getb("SB_NO_JIT");
    t.no_binv2 = getb("SB_NO_BINV2");
    t.no_grand_shard = getb("SB_NO_GRAND_SHARD");
    t.grand_shard = getb("SB_GRAND_SHARD");
    t.shard_msm_by_window = getb("SB_SHARD_MSM_BY_WINDOW");
    t.shard_msm_by_residue = getb("SB_SHARD_MSM_BY_RESIDUE");
    t.no_shplonk_shard = getb("SB_NO_SHPLONK_SHARD");
    t.no_shplonk_lagrange = getb("SB_NO_SHPLONK_LAGRANGE");
    t.msm_no_bucket_tree = getb("SB_MSM_NO_BUCKET_TREE");
    t.no_inst_direct = getb("SB_NO_INST_DIRECT");
    t.no_tables = getb("SB_NO_TABLES");
    t.no_smallkey_sort = getb("SB_NO_SMALLKEY_SORT");
}
cudaError_t sync_stream(sb_ctx *ctx, cudaStream_t st) {
    if (!ctx->blocking_sync) return cudaStreamSynchronize(st);
    if (!ctx->block_ev) {
        cudaError_t e = cudaEventCreateWithFlags(&ctx->block_ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaEventRecord(ctx->block_ev, st);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->block_ev);
}
void ctx_retain(sb_ctx *ctx) { ctx->refs.fetch_add(1, std::memory_order_relaxed); }
void ctx_release(sb_ctx *ctx) {
    if (ctx->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
    ntt_plans_free(ctx);
    for (auto &kv : ctx->scratch)
        if (kv.second.ptr) cudaFree(kv.second.ptr);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    for (int i = 0; i < sb_ctx::STAGE_SLOTS; i++)
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    for (int e = 0; e < 5; e++)
        if (ctx->msm_ev[e]) cudaEventDestroy(ctx->msm_ev[e]);
    if (ctx->block_ev) cudaEventDestroy(ctx->block_ev);
    for (int i = 0; i < 2; i++) {
        if (ctx->side_ev[i]) cudaEventDestroy(ctx->side_ev[i]);
        if (ctx->h_ev[i]) cudaEventDestroy(ctx->h_ev[i]);
    }
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
    delete ctx;
}

// ---- integer-roof probes --------------------------------------------------------------------
template <class P>
__global__ void bench_field_mul_kernel(uint4 *out, uint32_t iters) {
    // 4 independent product chains per thread (ILP) so the FMA pipe, not latency, is the limit
    Fp<P> a[4], b = Fp<P>::r2();
#pragma unroll
    for (int q = 0; q < 4; q++) {
        a[q] = Fp<P>::one();
        a[q].v[0] += threadIdx.x + q;
    }
    b.v[1] ^= blockIdx.x;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) a[q] = mul(a[q], b);
    }
    Fp<P> r = add(add(a[0], a[1]), add(a[2], a[3]));
    if (r.v[0] == 0x12345678u) store_fp(out + 2 * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x), r);
}

__global__ void bench_imad_kernel(uint32_t *out, uint32_t iters) {
    // 8 independent IMAD chains per thread (32-bit multiply-add, the unit MEASURED_PEAKS lacks)
    uint32_t x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = threadIdx.x * 2654435761u + q;
    const uint32_t m = blockIdx.x | 1u, c = threadIdx.x;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) x[q] = x[q] * m + c;
    }
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) r ^= x[q];
    if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void bench_imad_wide_kernel(uint64_t *out, uint32_t iters) {
    // 8 independent IMAD.WIDE chains (32x32+64 -> 64), the instruction the field product is made of
    uint64_t x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = threadIdx.x * 2654435761ull + q;
    const uint32_t m = blockIdx.x | 1u;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) x[q] = (uint64_t)(uint32_t)x[q] * m + x[q];
    }
    uint64_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) r ^= x[q];
    if (r == 0x12345678ull) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__global__ void bench_imad_hi_kernel(uint32_t *out, uint32_t iters) {
    // 8 independent IMAD.HI chains (upper half of 32x32, + 32-bit addend): is the high half as cheap as the low half, or as dear as IMAD.WIDE?
    uint32_t x[8];
#pragma unroll
    for (int q = 0; q < 8; q++) x[q] = threadIdx.x * 2654435761u + q;
    const uint32_t m = 0xfffffff1u - blockIdx.x, c = threadIdx.x | 0x80000000u;
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) x[q] = __umulhi(x[q], m) + c;
    }
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) r ^= x[q];
    if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace sb

using namespace sb;

#include "handles.h"

namespace {
typedef sb::CtxGuard Guard;

fr_t fr_inv_host(const fr_t &a) { return inv(a); }
fr_t fr_from_hex_limbs(const uint32_t canon[8]) {
    fr_t a;
    for (int i = 0; i < 8; i++) a.v[i] = canon[i];
    return to_mont(a);
}
}  // namespace

extern "C" {

int32_t sb_version(void) { return 100; }
const char *sb_last_error(void) { return sb::g_err; }

int32_t sb_device_count(int32_t *out_count) {
    if (!out_count) return SB_ERR_ARG;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        *out_count = 0;
        set_last_error("no CUDA device visible (%s); libsumma_b200 has no CPU fallback", cudaGetErrorString(e));
        return SB_ERR_NO_DEVICE;
    }
    *out_count = n;
    return SB_OK;
}

static int32_t ctx_init(sb_ctx *c, int32_t device) {
    c->device = device;
    cudaDeviceProp prop;
    SB_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    SB_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    SB_CUDA_TRY(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) SB_CUDA_TRY(cudaEventCreateWithFlags(&c->side_ev[i], cudaEventDisableTiming));
    c->pinned_bytes = 1 << 20;
    SB_CUDA_TRY(cudaHostAlloc(&c->pinned, c->pinned_bytes, cudaHostAllocDefault));
    SB_CUDA_TRY(cudaHostAlloc((void **)&c->stage, sb_ctx::STAGE_SLOTS * sb_ctx::STAGE_SLOT_BYTES, cudaHostAllocDefault));
    for (int i = 0; i < sb_ctx::STAGE_SLOTS; i++) SB_CUDA_TRY(cudaEventCreateWithFlags(&c->stage_ev[i], cudaEventDisableTiming));
    return SB_OK;
}

int32_t sb_ctx_create(int32_t device, sb_ctx **out_ctx) {
    if (!out_ctx) return SB_ERR_ARG;
    *out_ctx = nullptr;
    int32_t n = 0;
    SB_TRY(sb_device_count(&n));
    SB_REQUIRE(device >= 0 && device < n, "sb_ctx_create: device index out of range");
    int prev = -1;
    cudaGetDevice(&prev);
    SB_CUDA_TRY(cudaSetDevice(device));
    sb_ctx *c = new sb_ctx();
    sb::ctx_read_env(c);
    const int32_t rc = ctx_init(c, device);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);  // the caller's current device is not ours to change
    if (rc != SB_OK) {
        c->device = device;
        sb::ctx_release(c);  // frees whatever was created before the failing call
        return rc;
    }
    *out_ctx = c;
    return SB_OK;
}

int32_t sb_ctx_destroy(sb_ctx *ctx) {
    if (!ctx) return SB_OK;
    sb::ctx_release(ctx);  // the caller's reference; children created on this context (sb_mst, sb_pk) may still hold theirs
    return SB_OK;
}

int32_t sb_ctx_synchronize(sb_ctx *ctx) {
    if (!ctx) return SB_ERR_ARG;
    Guard g(ctx);
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int32_t sb_ctx_set_blocking_sync(sb_ctx *ctx, int32_t on) {
    if (!ctx) return SB_ERR_ARG;
    Guard g(ctx);
    ctx->blocking_sync = on != 0;
    return SB_OK;
}
int32_t sb_ctx_stream(const sb_ctx *ctx, void **out_stream) {
    if (!ctx || !out_stream) return SB_ERR_ARG;
    *out_stream = (void *)ctx->stream;
    return SB_OK;
}

int32_t sb_dev_alloc(sb_ctx *ctx, size_t bytes, void **out_dptr) {
    if (!ctx || !out_dptr) return SB_ERR_ARG;
    Guard g(ctx);
    cudaError_t e = cudaMalloc(out_dptr, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        set_last_error("sb_dev_alloc(%zu): %s", bytes, cudaGetErrorString(e));
        return SB_ERR_ALLOC;
    }
    return SB_OK;
}
int32_t sb_dev_free(sb_ctx *ctx, void *dptr) {
    if (!ctx) return SB_ERR_ARG;
    Guard g(ctx);
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    SB_CUDA_TRY(cudaFree(dptr));
    return SB_OK;
}
int32_t sb_dev_upload(sb_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return SB_ERR_ARG;
    Guard g(ctx);
    SB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_dev_download(sb_ctx *ctx, void *dst, const void *src, size_t bytes) {
    if (!ctx || (!dst && bytes) || (!src && bytes)) return SB_ERR_ARG;
    Guard g(ctx);
    SB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

static int32_t vec_op_host(sb_ctx *ctx, int field, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) {
    if (!ctx || !a || !b || !out) return SB_ERR_ARG;
    Guard g(ctx);
    void *da, *db;
    SB_TRY(scratch_get(ctx, "vec_a", n * 32, &da));
    SB_TRY(scratch_get(ctx, "vec_b", n * 32, &db));
    SB_CUDA_TRY(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(fp_vec_op(ctx, field, op, da, db, da, n, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(out, da, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_fr_vec_op(sb_ctx *ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) { return vec_op_host(ctx, 0, op, a, b, out, n); }
int32_t sb_fq_vec_op(sb_ctx *ctx, int32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out, size_t n) { return vec_op_host(ctx, 1, op, a, b, out, n); }

// ---- MSM ------------------------------------------------------------------------------------
int32_t sb_msm_g1_dev(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, uint8_t out_affine[64], void *stream) {
    if (!ctx || !out_affine || (n && (!d_bases || !d_scalars))) return SB_ERR_ARG;
    Guard g(ctx);
    return msm_run(ctx, d_bases, d_scalars, n, out_affine, pick_stream(ctx, stream));
}

int32_t sb_best_multiexp(sb_ctx *ctx, const uint8_t *coeffs, const uint8_t *bases, size_t n, uint8_t out_jacobian[96]) {
    if (!ctx || !out_jacobian || (n && (!coeffs || !bases))) return SB_ERR_ARG;
    Guard g(ctx);
    void *db, *ds;
    SB_TRY(scratch_get(ctx, "mx_bases", n * 64, &db));
    SB_TRY(scratch_get(ctx, "mx_scalars", n * 32, &ds));
    SB_CUDA_TRY(cudaMemcpyAsync(db, bases, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(ds, coeffs, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t aff[64];
    SB_TRY(msm_run(ctx, db, ds, n, aff, ctx->stream));
    bool id = true;
    for (int i = 0; i < 64; i++) id = id && aff[i] == 0;
    memset(out_jacobian, 0, 96);
    fq_t one = fq_t::one();
    if (id) {
        memcpy(out_jacobian + 32, one.v, 32);  // halo2curves identity: (0, 1, 0)
    } else {
        memcpy(out_jacobian, aff, 64);
        memcpy(out_jacobian + 64, one.v, 32);
    }
    return SB_OK;
}

int32_t sb_g1_fixed_base_mul_dev(sb_ctx *ctx, const void *d_scalars, size_t n, void *d_out_affine, void *stream) {
    if (!ctx || (n && (!d_scalars || !d_out_affine))) return SB_ERR_ARG;
    Guard g(ctx);
    return g1_fixed_base_mul(ctx, d_scalars, n, d_out_affine, pick_stream(ctx, stream));
}

int32_t sb_srs_upload(sb_ctx *ctx, uint32_t k, const uint8_t *g_pts, const uint8_t *g_lagrange, sb_srs **out_srs) {
    if (!ctx || !g_pts || !g_lagrange || !out_srs) return SB_ERR_ARG;
    SB_REQUIRE(k <= 28, "sb_srs_upload: k > 28");
    Guard g(ctx);
    sb_srs *s = new sb_srs();
    s->k = k;
    const size_t bytes = (size_t)64 << k;
    if (cudaMalloc(&s->d_g, bytes) != cudaSuccess || cudaMalloc(&s->d_g_lagrange, bytes) != cudaSuccess) {
        set_last_error("sb_srs_upload: cudaMalloc(2 x %zu) failed", bytes);
        if (s->d_g) cudaFree(s->d_g);
        delete s;
        return SB_ERR_ALLOC;
    }
    SB_CUDA_TRY(cudaMemcpyAsync(s->d_g, g_pts, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(s->d_g_lagrange, g_lagrange, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *out_srs = s;
    return SB_OK;
}
int32_t sb_srs_wrap_dev(sb_ctx *ctx, uint32_t k, const void *d_g, const void *d_g_lagrange, sb_srs **out_srs) {
    if (!ctx || !d_g || !d_g_lagrange || !out_srs) return SB_ERR_ARG;
    SB_REQUIRE(k <= 28, "sb_srs_wrap_dev: k > 28");
    sb_srs *s = new sb_srs();
    s->k = k;
    s->d_g = const_cast<void *>(d_g);
    s->d_g_lagrange = const_cast<void *>(d_g_lagrange);
    s->borrowed = true;
    *out_srs = s;
    return SB_OK;
}
int32_t sb_srs_destroy(sb_srs *srs) {
    if (!srs) return SB_OK;
    if (srs->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return SB_OK;
    if (!srs->borrowed) {
        cudaFree(srs->d_g);
        cudaFree(srs->d_g_lagrange);
    }
    if (srs->tab_slab) {
        cudaFree(srs->tab_slab);
    } else {
        for (int b = 0; b < 2; b++)
            if (srs->tab[b].d_tables && (b == 0 || srs->tab[1].d_tables != srs->tab[0].d_tables)) cudaFree(srs->tab[b].d_tables);
    }
    delete srs;
    return SB_OK;
}
int32_t sb_msm_g1_srs_dev(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const void *d_scalars, size_t n, uint8_t out_affine[64], void *stream) {
    if (!ctx || !srs || !out_affine || (n && !d_scalars)) return SB_ERR_ARG;
    SB_REQUIRE(basis == SB_BASIS_MONOMIAL || basis == SB_BASIS_LAGRANGE, "basis must be 0 or 1");
    SB_REQUIRE(n <= ((size_t)1 << srs->k), "msm: more scalars than SRS bases");
    Guard g(ctx);
    return srs_msm(ctx, srs, basis, d_scalars, n, out_affine, pick_stream(ctx, stream));
}
int32_t sb_srs_precompute(sb_ctx *ctx, sb_srs *srs, int32_t basis_mask, uint32_t window_bits) {
    if (!ctx || !srs) return SB_ERR_ARG;
    Guard g(ctx);
    return srs_precompute_impl(ctx, srs, basis_mask, window_bits);
}
}  // extern "C"
namespace sb {
int32_t srs_precompute_impl(sb_ctx *ctx, sb_srs *srs, int32_t basis_mask, uint32_t window_bits) {
    SB_REQUIRE(basis_mask >= 1 && basis_mask <= 3, "sb_srs_precompute: basis_mask is a bit set of (1 << SB_BASIS_*)");
    if (srs->k < 11) return SB_OK;  // tiny SRS: plain Pippenger is launch-bound either way
    uint32_t c = window_bits;
    if (c == 0) {
        // window bits: fewer windows (level-1 additions) against more buckets (2 general additions each in the bucket reduction); measured on the k = 20
        // proof (profiles/r02e_window_sweep.txt): c = 17 -> 70.0 ms, 18 -> 71.4, 19 -> 71.2, 20 -> 72.1
        c = srs->k <= 19 ? srs->k : (srs->k == 20 ? 17 : (srs->k >= 23 ? 22 : 20));
        if (ctx->tune.tab_c > 0) c = (uint32_t)ctx->tune.tab_c;
    }
    std::lock_guard<std::mutex> tab_lock(srs->tab_mu);
    if (basis_mask == 3 && !srs->tab[0].d_tables && !srs->tab[1].d_tables && srs->d_g_lagrange != srs->d_g) {
        // both bases at once: one slab, monomial tables first
        const uint32_t W = (255 + c - 1) / c;
        const size_t one = ((size_t)W << srs->k) * 64;
        void *slab = nullptr;
        if (c >= 11 && c <= 24 && cudaMalloc(&slab, 2 * one) == cudaSuccess) {
            int32_t rc = msm_tables_build(ctx, srs->d_g, (size_t)1 << srs->k, c, &srs->tab[0], ctx->stream, slab);
            if (rc == SB_OK) rc = msm_tables_build(ctx, srs->d_g_lagrange, (size_t)1 << srs->k, c, &srs->tab[1], ctx->stream, (uint8_t *)slab + one);
            if (rc != SB_OK) {
                cudaFree(slab);
                srs->tab[0] = MsmTables();
                srs->tab[1] = MsmTables();
                return rc;
            }
            srs->tab_slab = slab;
            return SB_OK;
        }
        cudaGetLastError();  // fall through to separate allocations (or to the argument check inside the builder)
    }
    for (int b = 0; b < 2; b++) {
        if (!((basis_mask >> b) & 1) || srs->tab[b].d_tables) continue;
        if (b == 1 && srs->d_g_lagrange == srs->d_g && srs->tab[0].d_tables) { srs->tab[1] = srs->tab[0]; continue; }
        SB_TRY(msm_tables_build(ctx, b == 0 ? srs->d_g : srs->d_g_lagrange, (size_t)1 << srs->k, c, &srs->tab[b], ctx->stream));
    }
    return SB_OK;
}
}  // namespace sb
extern "C" {
int32_t sb_msm_g1(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const uint8_t *scalars, size_t n, uint8_t out_affine[64]) {
    if (!ctx || !srs || !out_affine || (n && !scalars)) return SB_ERR_ARG;
    SB_REQUIRE(basis == SB_BASIS_MONOMIAL || basis == SB_BASIS_LAGRANGE, "basis must be 0 or 1");
    SB_REQUIRE(n <= ((size_t)1 << srs->k), "msm: more scalars than SRS bases");
    Guard g(ctx);
    void *ds;
    SB_TRY(scratch_get(ctx, "mx_scalars", n * 32, &ds));
    SB_CUDA_TRY(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return srs_msm(ctx, srs, basis, ds, n, out_affine, ctx->stream);
}

// ---- NTT ------------------------------------------------------------------------------------
int32_t sb_ntt_dev(sb_ctx *ctx, void *d_a, const uint8_t omega[32], uint32_t log_n, void *stream) {
    if (!ctx || !d_a || !omega) return SB_ERR_ARG;
    Guard g(ctx);
    return ntt_run(ctx, d_a, omega, log_n, pick_stream(ctx, stream));
}
int32_t sb_ntt_dist(sb_ctx *ctx, const sb_comm *comm, void *d_a, const uint8_t omega[32], uint32_t log_n, const uint8_t *scale, void *stream) {
    if (!ctx || !comm || !d_a || !omega) return SB_ERR_ARG;
    SB_REQUIRE(comm->world >= 1 && comm->rank >= 0 && comm->rank < comm->world, "sb_comm: rank / world out of range");
    Guard g(ctx);
    fr_t s;
    if (scale) memcpy(s.v, scale, 32);
    return ntt_run_dist(ctx, comm, d_a, omega, log_n, scale ? &s : nullptr, pick_stream(ctx, stream));
}
int32_t sb_best_fft(sb_ctx *ctx, uint8_t *a, const uint8_t omega[32], uint32_t log_n) {
    if (!ctx || !a || !omega) return SB_ERR_ARG;
    SB_REQUIRE(log_n <= 28, "best_fft: log_n > 28");
    Guard g(ctx);
    const size_t bytes = (size_t)32 << log_n;
    void *d;
    SB_TRY(scratch_get(ctx, "fft_host", bytes, &d));
    SB_CUDA_TRY(cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(ntt_run(ctx, d, omega, log_n, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// ---- EvaluationDomain -------------------------------------------------------------------------
int32_t sb_domain_create(sb_ctx *ctx, uint32_t j, uint32_t k, sb_domain **out_domain) {
    if (!ctx || !out_domain) return SB_ERR_ARG;
    SB_REQUIRE(j >= 2, "EvaluationDomain::new: j (constraint degree) must be >= 2");
    sb_domain *d = new sb_domain();
    d->j = j;
    d->k = k;
    d->quotient_degree = j - 1;
    uint32_t ext = k;
    while ((1ull << ext) < ((uint64_t)d->quotient_degree << k)) ext++;
    d->ext_k = ext;
    if (ext > 28) {
        delete d;
        set_last_error("EvaluationDomain::new: extended_k %u > 28", ext);
        return SB_ERR_ARG;
    }
    // halo2curves bn256::Fr constants, canonical little-endian 32-bit limbs
    static const uint32_t ROOT_OF_UNITY[8] = {0x60c37c9cu, 0xd34f1ed9u, 0xd39329c8u, 0x3215cf6du, 0x3dd31f74u, 0x98865ea9u, 0x166d18b7u, 0x03ddb9f5u};
    static const uint32_t ZETA[8] = {0x36636f23u, 0xb8ca0b2du, 0xec2bc5e9u, 0xcc37a73fu, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u};
    fr_t root = fr_from_hex_limbs(ROOT_OF_UNITY);
    fr_t w = root;
    for (uint32_t i = ext; i < 28; i++) w = sqr(w);
    d->ext_omega = w;
    for (uint32_t i = k; i < ext; i++) w = sqr(w);
    d->omega = w;
    d->omega_inv = fr_inv_host(d->omega);
    d->ext_omega_inv = fr_inv_host(d->ext_omega);
    d->ifft_divisor = fr_inv_host(fr_from_u64_host(1ull << k));
    d->ext_ifft_divisor = fr_inv_host(fr_from_u64_host(1ull << ext));
    fr_t zeta = fr_from_hex_limbs(ZETA), zeta2 = sqr(zeta);
    d->coset[0] = fr_t::one(); d->coset[1] = zeta; d->coset[2] = zeta2;
    d->coset_inv[0] = fr_t::one(); d->coset_inv[1] = zeta2; d->coset_inv[2] = zeta;
    // t(X) = X^n - 1 on the extended coset: 2^(ext-k) distinct values, inverted
    d->n_t = 1u << (ext - k);
    if (d->n_t > 8) {
        delete d;
        set_last_error("EvaluationDomain: extended_k - k > 3 is not supported");
        return SB_ERR_ARG;
    }
    fr_t cur = fr_pow_host(zeta, 1ull << k);
    fr_t step = fr_pow_host(d->ext_omega, 1ull << k);
    for (uint32_t i = 0; i < d->n_t; i++) {
        d->t_inv[i] = fr_inv_host(sub(cur, fr_t::one()));
        cur = mul(cur, step);
    }
    *out_domain = d;
    return SB_OK;
}
int32_t sb_domain_destroy(sb_domain *domain) { delete domain; return SB_OK; }
int32_t sb_domain_extended_k(const sb_domain *domain, uint32_t *out) {
    if (!domain || !out) return SB_ERR_ARG;
    *out = domain->ext_k;
    return SB_OK;
}

// EvaluationDomain's scalings ride inside the transform (ntt.cu NttFuse): no separate pass over the data for n^-1, the zeta-coset pattern,
// the zero padding to the extended size, t(X)^-1 or the truncation to (j - 1) n coefficients
static int32_t l2c_dev(sb_ctx *ctx, const sb_domain *d, void *d_a, cudaStream_t st) {
    NttFuse f;
    f.has_scale = true;
    f.scale = d->ifft_divisor;
    return ntt_run_fused(ctx, d_a, d_a, (const uint8_t *)d->omega_inv.v, d->k, &f, st);
}
static int32_t c2e_dev(sb_ctx *ctx, const sb_domain *d, const void *d_coeff, void *d_ext, cudaStream_t st) {
    // coefficient i is scaled by zeta^(i mod 3) as it is loaded; coefficients n .. 2^ext_k are zeros that are never read
    NttFuse f;
    f.pre_m = 3;
    for (int i = 0; i < 3; i++) f.pre_pat[i] = d->coset[i];
    f.n_in = (uint64_t)1 << d->k;
    return ntt_run_fused(ctx, d_coeff, d_ext, (const uint8_t *)d->ext_omega.v, d->ext_k, &f, st);
}
// divide != 0: the division by t(X) = X^n - 1 (a pattern of 2^(ext_k - k) constants) is applied to the values as they are loaded
static int32_t e2c_dev(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *d_coeff, int divide, cudaStream_t st) {
    NttFuse f;
    f.has_scale = true;
    f.scale = d->ext_ifft_divisor;
    f.post_m = 3;
    for (int i = 0; i < 3; i++) f.post_pat[i] = d->coset_inv[i];  // zeta^-(i mod 3)
    f.n_out = (uint64_t)d->quotient_degree << d->k;                // truncated to (j - 1) n coefficients
    if (divide) {
        f.pre_m = d->n_t;
        for (uint32_t i = 0; i < d->n_t; i++) f.pre_pat[i] = d->t_inv[i];
    }
    return ntt_run_fused(ctx, d_ext, d_coeff, (const uint8_t *)d->ext_omega_inv.v, d->ext_k, &f, st);
}

int32_t sb_lagrange_to_coeff_dev(sb_ctx *ctx, const sb_domain *d, void *d_a, void *stream) {
    if (!ctx || !d || !d_a) return SB_ERR_ARG;
    Guard g(ctx);
    return l2c_dev(ctx, d, d_a, pick_stream(ctx, stream));
}
int32_t sb_coeff_to_lagrange_dev(sb_ctx *ctx, const sb_domain *d, void *d_a, void *stream) {
    if (!ctx || !d || !d_a) return SB_ERR_ARG;
    Guard g(ctx);
    return ntt_run(ctx, d_a, (const uint8_t *)d->omega.v, d->k, pick_stream(ctx, stream));
}
int32_t sb_coeff_to_extended_dev(sb_ctx *ctx, const sb_domain *d, const void *d_coeff, void *d_ext, void *stream) {
    if (!ctx || !d || !d_coeff || !d_ext) return SB_ERR_ARG;
    Guard g(ctx);
    return c2e_dev(ctx, d, d_coeff, d_ext, pick_stream(ctx, stream));
}
int32_t sb_extended_to_coeff_dev(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *d_coeff, void *stream) {
    if (!ctx || !d || !d_coeff || !d_ext) return SB_ERR_ARG;
    Guard g(ctx);
    return e2c_dev(ctx, d, d_ext, d_coeff, 0, pick_stream(ctx, stream));
}
int32_t sb_divide_by_vanishing_poly_dev(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *stream) {
    if (!ctx || !d || !d_ext) return SB_ERR_ARG;
    Guard g(ctx);
    return fr_scale_pattern(ctx, d_ext, (size_t)1 << d->ext_k, d->t_inv, d->n_t, pick_stream(ctx, stream));
}

}  // extern "C"
namespace sb {
int32_t dom_l2c(sb_ctx *ctx, const sb_domain *d, void *d_a, cudaStream_t st) { return l2c_dev(ctx, d, d_a, st); }
int32_t dom_c2e(sb_ctx *ctx, const sb_domain *d, const void *d_coeff, void *d_ext, cudaStream_t st) { return c2e_dev(ctx, d, d_coeff, d_ext, st); }
int32_t dom_e2c(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *d_coeff, cudaStream_t st) { return e2c_dev(ctx, d, d_ext, d_coeff, 0, st); }
int32_t dom_div_e2c(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *d_coeff, cudaStream_t st) { return e2c_dev(ctx, d, d_ext, d_coeff, 1, st); }
int32_t dom_div_vanishing(sb_ctx *ctx, const sb_domain *d, void *d_ext, cudaStream_t st) { return fr_scale_pattern(ctx, d_ext, (size_t)1 << d->ext_k, d->t_inv, d->n_t, st); }
}  // namespace sb
extern "C" {
// host-buffer forms: stage through scratch on the context stream
static int32_t host_roundtrip(sb_ctx *ctx, const uint8_t *in, size_t n_in, uint8_t *out, size_t n_out, size_t n_dev, void **d_out) {
    void *dv;
    SB_TRY(scratch_get(ctx, "dom_host", n_dev * 32, &dv));
    if (in) SB_CUDA_TRY(cudaMemcpyAsync(dv, in, n_in * 32, cudaMemcpyHostToDevice, ctx->stream));
    (void)out; (void)n_out;
    *d_out = dv;
    return SB_OK;
}
int32_t sb_lagrange_to_coeff(sb_ctx *ctx, const sb_domain *d, uint8_t *a) {
    if (!ctx || !d || !a) return SB_ERR_ARG;
    Guard g(ctx);
    const size_t n = (size_t)1 << d->k;
    void *dv;
    SB_TRY(host_roundtrip(ctx, a, n, nullptr, 0, n, &dv));
    SB_TRY(l2c_dev(ctx, d, dv, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, dv, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_coeff_to_lagrange(sb_ctx *ctx, const sb_domain *d, uint8_t *a) {
    if (!ctx || !d || !a) return SB_ERR_ARG;
    Guard g(ctx);
    const size_t n = (size_t)1 << d->k;
    void *dv;
    SB_TRY(host_roundtrip(ctx, a, n, nullptr, 0, n, &dv));
    SB_TRY(ntt_run(ctx, dv, (const uint8_t *)d->omega.v, d->k, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, dv, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_coeff_to_extended(sb_ctx *ctx, const sb_domain *d, const uint8_t *coeff, uint8_t *ext) {
    if (!ctx || !d || !coeff || !ext) return SB_ERR_ARG;
    Guard g(ctx);
    const size_t n = (size_t)1 << d->k, ne = (size_t)1 << d->ext_k;
    void *dc, *de;
    SB_TRY(scratch_get(ctx, "dom_host_c", n * 32, &dc));
    SB_TRY(scratch_get(ctx, "dom_host", ne * 32, &de));
    SB_CUDA_TRY(cudaMemcpyAsync(dc, coeff, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(c2e_dev(ctx, d, dc, de, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(ext, de, ne * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_extended_to_coeff(sb_ctx *ctx, const sb_domain *d, const uint8_t *ext, uint8_t *coeff) {
    if (!ctx || !d || !coeff || !ext) return SB_ERR_ARG;
    Guard g(ctx);
    const size_t ne = (size_t)1 << d->ext_k, n_out = (size_t)d->quotient_degree << d->k;
    void *de;
    SB_TRY(host_roundtrip(ctx, ext, ne, nullptr, 0, ne, &de));
    SB_TRY(e2c_dev(ctx, d, de, de, 0, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(coeff, de, n_out * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_divide_by_vanishing_poly(sb_ctx *ctx, const sb_domain *d, uint8_t *ext) {
    if (!ctx || !d || !ext) return SB_ERR_ARG;
    Guard g(ctx);
    const size_t ne = (size_t)1 << d->ext_k;
    void *de;
    SB_TRY(host_roundtrip(ctx, ext, ne, nullptr, 0, ne, &de));
    SB_TRY(fr_scale_pattern(ctx, de, ne, d->t_inv, d->n_t, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(ext, de, ne * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// ---- instrumentation ----------------------------------------------------------------------------
int32_t sb_launch_count(const sb_ctx *ctx, uint64_t *out) {
    if (!ctx || !out) return SB_ERR_ARG;
    *out = ctx->launches;
    return SB_OK;
}

int32_t sb_msm_phase_times(const sb_ctx *ctx, float out_ms[5], uint32_t out_shape[4]) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    for (int i = 0; i < 5; i++) out_ms[i] = ctx->msm_phase_ms[i];
    if (out_shape)
        for (int i = 0; i < 4; i++) out_shape[i] = ctx->msm_last_shape[i];
    return SB_OK;
}

static int32_t timed(sb_ctx *ctx, float *out_ms, int which, uint32_t blocks, uint32_t threads, uint32_t iters) {
    void *d;
    SB_TRY(scratch_get(ctx, "bench_out", (size_t)blocks * threads * 32 + 64, &d));
    cudaEvent_t e0, e1;
    SB_CUDA_TRY(cudaEventCreate(&e0));
    SB_CUDA_TRY(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; rep++) {  // first repetition is the warm-up
        SB_CUDA_TRY(cudaEventRecord(e0, ctx->stream));
        if (which == 0) SB_LAUNCH(ctx, bench_field_mul_kernel<FrParams>, blocks, threads, 0, ctx->stream, (uint4 *)d, iters);
        else if (which == 1) SB_LAUNCH(ctx, bench_field_mul_kernel<FqParams>, blocks, threads, 0, ctx->stream, (uint4 *)d, iters);
        else if (which == 2) SB_LAUNCH(ctx, bench_imad_kernel, blocks, threads, 0, ctx->stream, (uint32_t *)d, iters);
        else if (which == 3) SB_LAUNCH(ctx, bench_imad_wide_kernel, blocks, threads, 0, ctx->stream, (uint64_t *)d, iters);
        else SB_LAUNCH(ctx, bench_imad_hi_kernel, blocks, threads, 0, ctx->stream, (uint32_t *)d, iters);
        SB_CUDA_TRY(cudaEventRecord(e1, ctx->stream));
        SB_CUDA_TRY(cudaEventSynchronize(e1));
    }
    SB_CUDA_TRY(cudaEventElapsedTime(out_ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return SB_OK;
}
int32_t sb_bench_field_mul(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, int32_t field, float *out_ms) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    Guard g(ctx);
    return timed(ctx, out_ms, field ? 1 : 0, blocks, threads, iters);
}
int32_t sb_bench_imad(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, float *out_ms) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    Guard g(ctx);
    return timed(ctx, out_ms, 2, blocks, threads, iters);
}
int32_t sb_bench_imad_wide(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, float *out_ms) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    Guard g(ctx);
    return timed(ctx, out_ms, 3, blocks, threads, iters);
}
int32_t sb_bench_imad_hi(sb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, float *out_ms) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    Guard g(ctx);
    return timed(ctx, out_ms, 4, blocks, threads, iters);
}
int32_t sb_test_msm_host_tail(const uint8_t *fin, int32_t n_bits, int32_t shift, int32_t x_slot, int32_t log_mod, int32_t res, uint8_t out_affine[64]) {
    if (!fin || !out_affine || n_bits < 0 || n_bits > 16 || shift < 0 || x_slot < 0 || x_slot > 17 || log_mod < 0 || log_mod > 8 || res < 0 || res >= (1 << log_mod)) return SB_ERR_ARG;
    uint8_t pt[128];
    sb::host_bucket_combine(fin, n_bits, shift, x_slot, pt);
    if (log_mod) sb::host_residue_fixup(pt, fin, log_mod, res);
    sb::host_fold_windows(pt, 1, 0, out_affine);
    return SB_OK;
}
int32_t sb_g1_sum_affine(const uint8_t *pts, size_t n, uint8_t out_affine[64]) {
    if (!pts || !out_affine) return SB_ERR_ARG;
    sb::host_sum_affine(pts, (int)n, out_affine);
    return SB_OK;
}

}  // extern "C"
