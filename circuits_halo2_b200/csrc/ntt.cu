// Multi-pass NTT over BN254 Fr for sm_100a: register radix-8 tiles, cp.async staging, fused domain scalings.
//
// Replaces halo2_proofs::arithmetic::best_fft (reference call sites reach it through EvaluationDomain from
// zk_prover/src/circuits/utils.rs:75-76,94-102) and carries the scalings EvaluationDomain wraps around it: n^-1 of the
// inverse transforms, the zeta-coset patterns, t(X)^-1, zero padding and truncation are folded into the first pass' loads, the inter-pass
// twiddle table or the last pass' stores.  Natural order in, natural order out, any power-of-two size up to 2^28.
//
// Decomposition (mixed-radix Cooley-Tukey, "four-step" generalised to P passes):
//   N = R_1 R_2 ... R_P.  Before pass t the array is indexed [a][j][c] with a = (k_1 .. k_{t-1}) the digits already transformed, j < R_t the
//   digit transformed now and c < C_t the still-contiguous remainder.  A CTA owns a tile of R_t x G elements (2^11 elements = 64 KB of shared
//   memory with two CTAs per SM, or 2^12 = 128 KB with one), transforms digit j in registers (ntt_core.cuh), multiplies by the inter-pass
//   twiddle omega^(A_t c k) and writes row k back in place.  The last pass reads G rows whose leading digit k_1 is consecutive (contiguous rows,
//   staged by cp.async) and scatters them to out[k_1 + R_1 (k_2 + R_2 (...)) + A k]: one HBM round trip per pass, TWO passes up to 2^22 (2^24
//   with the 128 KB tile), three beyond.
// Twiddles: per-pass omega_R^e table (R/2 entries, cp.async'ed into shared memory, swizzled); inter-pass factors from ONE table per pass boundary
//   laid out exactly like the pass' output (coalesced 32-byte reads, one product per element, the plan's scale n^-1 folded in); plans whose
//   tables would exceed SB_NTT_TW_MB (default 1024 MiB) fall back to a two-level omega^lo * omega^(hi << h) pair (one more product).
// HBM is not the binding roof for a 254-bit field (~11 products = 1500 wide multiply-adds per 64 B moved per pass); what the layout buys
// is fewer products (11 per element at 2^22 instead of 13.5), fewer barriers and 3x less shared-memory traffic.
// The index algebra is pinned on the CPU: tests/host/host_ntt_harness.cpp runs ntt_core.cuh for every thread of every tile.
#include "common.cuh"
#include "ntt_core.cuh"

namespace sb {

static const uint32_t SMALL_MAX_LOG = 10;  // single-pass sizes up to here use the simple shared-memory kernel

struct NttPlan {
    uint32_t log_n = 0, tile_log = 11;
    int npass = 0;
    uint32_t radix[NTT_MAX_PASS];
    uint4 *w_block[NTT_MAX_PASS];
    uint4 *tw_full[NTT_MAX_PASS];  // per strided pass, or null
    uint4 *t_lo = nullptr, *t_hi = nullptr;
    uint32_t log_tlo = 0;
    fr_t scale;                   // folded into the first inter-pass boundary (multi-pass plans)
    void *tables = nullptr;       // w_block + t_lo + t_hi
    void *full_tables = nullptr;  // the tw_full tables
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

struct SmemTwiddle {
    const uint4 *lo, *hi;
    __device__ __forceinline__ nfr_t operator()(uint32_t e) const {
        const uint32_t s = ntt_swz(e);
        const uint4 a = lo[s], b = hi[s];
        nfr_t r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
        r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        return r;
    }
};

__device__ __forceinline__ nfr_t lds_fr(const uint4 *lo, const uint4 *hi, uint32_t i) {
    const uint4 a = lo[i], b = hi[i];
    nfr_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr(uint4 *lo, uint4 *hi, uint32_t i, const nfr_t &x) {
    lo[i] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    hi[i] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

// One tile of one pass.  blockDim.x = 2^(r + g - EB): 2^EB elements per thread (EB = 3: radix-8 stages, 128 registers, 16 warps / SM;
// EB = 2: radix-4 stages, one more exchange per pass, 64 registers, 32 warps / SM).
template <int T_LOG, int EB>
__global__ void __launch_bounds__(1 << (T_LOG - EB), T_LOG == 11 ? 2 : 1) ntt_pass_kernel(const NttPassArgs p) {
    extern __shared__ uint4 smem[];
    const NttGeom G(p);
    constexpr int E = 1 << EB;
    const uint32_t tile = 1u << G.t, T = tile >> EB, tid = threadIdx.x, half_r = 1u << (G.r - 1);
    uint4 *s_lo = smem, *s_hi = smem + tile;
    uint4 *w_lo = s_hi + tile, *w_hi = w_lo + half_r;
    const NttTileCoord tc(p, blockIdx.x + p.tile0);
    const NttBatch bo(p, blockIdx.y);

    // butterfly twiddles of this pass: global -> shared through cp.async (no register staging), under the data loads below
    for (uint32_t e = tid; e < half_r; e += T) {
        const uint32_t s = ntt_swz(e);
        cp_async16(w_lo + s, p.w + 2 * e);
        cp_async16(w_hi + s, p.w + 2 * e + 1);
    }

    nfr_t x[E];
    uint32_t pw = G.window(0);
    if (p.kind == NTT_LAST) {
        // the tile's G rows are contiguous in HBM: stage them with cp.async, coalesced, straight to their swizzled slots
        for (uint32_t i = tid; i < tile; i += T) {
            const uint64_t gi = tc.in_index(p, G.j_of(i), G.gg_of(i));
            const uint32_t s = ntt_swz(i);
            cp_async16(s_lo + s, p.src + 2 * (bo.src + gi));
            cp_async16(s_hi + s, p.src + 2 * (bo.src + gi) + 1);
        }
        cp_async_wait_all();
        __syncthreads();
#pragma unroll
        for (int b = 0; b < E; b++) x[b] = lds_fr(s_lo, s_hi, ntt_swz(G.idx(tid, pw, (uint32_t)b)));
    } else {
        const bool first = (p.log_a == 0);
#pragma unroll
        for (int b = 0; b < E; b++) {
            const uint32_t i = G.idx(tid, pw, (uint32_t)b);
            const uint64_t gi = tc.in_index(p, G.j_of(i), G.gg_of(i));
            x[b] = first ? ntt_fetch_input(p, bo, gi) : ntt_load_fr(p.src + 2 * (bo.src + gi));
        }
        cp_async_wait_all();
        __syncthreads();
    }

    const SmemTwiddle tw{w_lo, w_hi};
    uint32_t low = G.r, prev = pw;
    for (uint32_t s = 0; s < G.n_stages; s++) {
        pw = G.window(s);
        if (s > 0) {
            __syncthreads();  // every thread is done reading the previous contents of the exchange buffer
#pragma unroll
            for (int b = 0; b < E; b++) sts_fr(s_lo, s_hi, ntt_swz(G.idx(tid, prev, (uint32_t)b)), x[b]);
            __syncthreads();
#pragma unroll
            for (int b = 0; b < E; b++) x[b] = lds_fr(s_lo, s_hi, ntt_swz(G.idx(tid, pw, (uint32_t)b)));
        }
        ntt_stage_butterflies<EB>(x, G, tid, pw, low, tw);
        low = pw - G.jshift;
        prev = pw;
    }

    // ---- write back: register b holds output digit k = bitrev(j) of column gg ----
    if (p.kind == NTT_LAST && p.g > 0) {
        // contiguous runs on the output side need gg fastest across lanes: one more exchange
        __syncthreads();
#pragma unroll
        for (int b = 0; b < E; b++) sts_fr(s_lo, s_hi, ntt_swz(G.idx(tid, prev, (uint32_t)b)), x[b]);
        __syncthreads();
#pragma unroll
        for (int b = 0; b < E; b++) {
            const uint32_t m = tid + T * (uint32_t)b;
            const uint32_t gg = m & ((1u << p.g) - 1u), jj = m >> p.g;
            ntt_emit(p, bo, tc, ntt_brev(jj, G.r), gg, lds_fr(s_lo, s_hi, ntt_swz((gg << G.r) | jj)));
        }
    } else {
#pragma unroll
        for (int b = 0; b < E; b++) {
            const uint32_t i = G.idx(tid, prev, (uint32_t)b);
            ntt_emit(p, bo, tc, ntt_brev(G.j_of(i), G.r), G.gg_of(i), x[b]);
        }
    }
}

// sizes below 2^11: the whole transform in one CTA, one level per barrier (launch-bound territory)
__global__ void __launch_bounds__(256) ntt_small_kernel(const NttPassArgs p) {
    extern __shared__ uint4 smem[];
    const uint32_t r = p.r, R = 1u << r;
    uint4 *s_lo = smem, *s_hi = smem + R;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const NttBatch bo(p, blockIdx.y);
    for (uint32_t j = tid; j < R; j += nt) sts_fr(s_lo, s_hi, j, ntt_fetch_input(p, bo, j));
    __syncthreads();
    for (uint32_t l = 0; l < r; l++) {
        const uint32_t sh = r - 1 - l, h = 1u << sh;
        for (uint32_t q = tid; q < R / 2; q += nt) {
            const uint32_t i = ((q >> sh) << (sh + 1)) | (q & (h - 1));
            const uint32_t e = (q & (h - 1)) << l;
            const nfr_t u = lds_fr(s_lo, s_hi, i), v = lds_fr(s_lo, s_hi, i + h);
            sts_fr(s_lo, s_hi, i, add(u, v));
            nfr_t d = sub(u, v);
            if (l != r - 1) d = mul(d, ntt_load_fr(p.w + 2 * e));
            sts_fr(s_lo, s_hi, i + h, d);
        }
        __syncthreads();
    }
    const NttTileCoord tc(p, 0);
    for (uint32_t j = tid; j < R; j += nt) ntt_emit(p, bo, tc, r ? (__brev(j) >> (32 - r)) : 0, 0, lds_fr(s_lo, s_hi, j));
}

// out[i] = base^i  (i < count), per-thread square-and-multiply; table setup only
__global__ void gen_powers_kernel(uint4 *out, fr_t base, uint64_t count) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    fr_t acc = fr_t::one(), b = base;
    uint64_t e = i;
    while (e) {
        if (e & 1) acc = mul(acc, b);
        b = sqr(b);
        e >>= 1;
    }
    store_fp(out + 2 * i, acc);
}
// a[i] *= s
__global__ void scale_table_kernel(uint4 *a, fr_t s, uint64_t count) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < count) store_fp(a + 2 * i, mul(load_fp<FrParams>(a + 2 * i), s));
}
// out[(k << log_c) + c] = t_lo[E & mask] * t_hi[E >> log_tlo],  E = (c k) << log_a   (t_lo carries the plan's scale)
__global__ void gen_tw_full_kernel(uint4 *out, const uint4 *t_lo, const uint4 *t_hi, uint32_t log_tlo, uint32_t log_a, uint32_t log_c, uint64_t count) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t c = i & ((1ull << log_c) - 1), k = i >> log_c;
    const uint64_t E = (c * k) << log_a;
    fr_t tw = load_fp<FrParams>(t_lo + 2 * (E & ((1ull << log_tlo) - 1)));
    const uint64_t eh = E >> log_tlo;
    if (eh) tw = mul(tw, load_fp<FrParams>(t_hi + 2 * eh));
    store_fp(out + 2 * i, tw);
}

int32_t fr_gen_powers(sb_ctx *ctx, void *d_out, const fr_t &base, size_t count, cudaStream_t st) {
    if (count == 0) return SB_OK;
    SB_LAUNCH(ctx, gen_powers_kernel, (unsigned)((count + 127) / 128), 128, 0, st, (uint4 *)d_out, base, (uint64_t)count);
    return SB_OK;
}

// ---- host side ------------------------------------------------------------------------------
fr_t fr_pow_host(const fr_t &base, uint64_t e) {
    fr_t acc = fr_t::one(), b = base;
    while (e) {
        if (e & 1) acc = mul(acc, b);
        b = sqr(b);
        e >>= 1;
    }
    return acc;
}
fr_t fr_from_u64_host(uint64_t x) {
    fr_t a = fr_t::zero();
    a.v[0] = (uint32_t)x;
    a.v[1] = (uint32_t)(x >> 32);
    return to_mont(a);
}

// pass radices, most significant digit first; mirrored by tests/host/host_ntt_harness.cpp through sb_test_ntt_plan
void ntt_make_plan(uint32_t log_n, uint32_t tile_log, uint32_t max_passes_hint, int *npass, uint32_t *radix) {
    if (log_n <= SMALL_MAX_LOG || log_n <= tile_log) {
        *npass = 1;
        radix[0] = log_n;
        return;
    }
    uint32_t p = (log_n + tile_log - 1) / tile_log;
    if (max_passes_hint > p) p = max_passes_hint;
    const uint32_t base = log_n / p, extra = log_n % p;
    *npass = (int)p;
    for (uint32_t t = 0; t < p; t++) radix[t] = base + (t < extra ? 1 : 0);
}

static uint32_t pick_tile_log(const sb_ctx *ctx, uint32_t log_n) {
    if (ctx->tune.ntt_tile >= 9 && ctx->tune.ntt_tile <= 12) return (uint32_t)ctx->tune.ntt_tile;
    if (log_n == 12) return 12;              // one 4096-element tile instead of two passes
    if (log_n == 11) return 11;
    // up to 2^20 two passes fit 2^10-element tiles: 128-thread CTAs, four per SM instead of two -- barriers and load phases of one tile hide under
    // three others (2^20: 0.265 vs 0.279 ms, 2^16: 0.054 vs 0.071 ms; profiles/r02v_ntt_tile_sweep.txt)
    if (log_n <= 20) return 10;
    return (log_n == 23 || log_n == 24) ? 12 : 11;  // two passes up to 2^24
}

static int32_t plan_get(sb_ctx *ctx, const uint8_t omega[32], uint32_t log_n, const fr_t &scale, cudaStream_t st, NttPlan **out) {
    std::string key((const char *)omega, 32);
    key.append((const char *)scale.v, 32);
    key.push_back((char)log_n);
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) {
        *out = it->second;
        return SB_OK;
    }
    NttPlan *pl = new NttPlan();
    pl->log_n = log_n;
    pl->scale = scale;
    pl->tile_log = pick_tile_log(ctx, log_n);
    ntt_make_plan(log_n, pl->tile_log, (uint32_t)(ctx->tune.ntt_passes > 0 ? ctx->tune.ntt_passes : 0), &pl->npass, pl->radix);
    pl->log_tlo = (log_n + 1) / 2;
    const uint64_t n_lo = 1ull << pl->log_tlo, n_hi = 1ull << (log_n - pl->log_tlo);
    uint64_t total = n_lo + n_hi;
    uint64_t off_w[NTT_MAX_PASS];
    for (int t = 0; t < pl->npass; t++) {
        off_w[t] = total;
        uint64_t cnt = (1ull << pl->radix[t]) / 2;
        total += cnt ? cnt : 1;
    }
    cudaError_t e = cudaMalloc(&pl->tables, total * 32);
    if (e != cudaSuccess) {
        delete pl;
        set_last_error("ntt plan: cudaMalloc(%llu) failed: %s", (unsigned long long)(total * 32), cudaGetErrorString(e));
        return SB_ERR_ALLOC;
    }
    auto fail = [&](int32_t rc) {
        cudaStreamSynchronize(st);
        cudaFree(pl->tables);
        if (pl->full_tables) cudaFree(pl->full_tables);
        delete pl;
        return rc;
    };
    uint4 *tb = (uint4 *)pl->tables;
    pl->t_lo = tb;
    pl->t_hi = tb + 2 * n_lo;
    fr_t w;
    memcpy(w.v, omega, 32);
    auto gen = [&](uint4 *dst, const fr_t &b, uint64_t cnt) -> int32_t {
        unsigned blocks = (unsigned)((cnt + 127) / 128);
        SB_LAUNCH(ctx, gen_powers_kernel, blocks, 128, 0, st, dst, b, cnt);
        return SB_OK;
    };
    int32_t rc = gen(pl->t_lo, w, n_lo);
    if (rc == SB_OK) rc = gen(pl->t_hi, fr_pow_host(w, n_lo), n_hi);
    for (int t = 0; t < pl->npass && rc == SB_OK; t++) {
        pl->w_block[t] = tb + 2 * off_w[t];
        pl->tw_full[t] = nullptr;
        uint64_t cnt = (1ull << pl->radix[t]) / 2;
        if (cnt == 0) cnt = 1;
        rc = gen(pl->w_block[t], fr_pow_host(w, 1ull << (log_n - pl->radix[t])), cnt);  // omega_R = omega^(N / R)
    }
    if (rc != SB_OK) return fail(rc);
    // one inter-pass twiddle table per strided pass, in the pass' output layout, when they fit the budget
    if (pl->npass > 1) {
        uint64_t full = 0, log_a = 0;
        for (int t = 0; t + 1 < pl->npass; t++) { full += 1ull << (log_n - log_a); log_a += pl->radix[t]; }
        if (full * 32 <= (uint64_t)ctx->tune.ntt_tw_mb << 20) {
            e = cudaMalloc(&pl->full_tables, full * 32);
            if (e == cudaSuccess) {
                uint64_t off = 0;
                log_a = 0;
                for (int t = 0; t + 1 < pl->npass; t++) {
                    const uint64_t cnt = 1ull << (log_n - log_a);
                    const uint32_t log_c = log_n - (uint32_t)log_a - pl->radix[t];
                    pl->tw_full[t] = (uint4 *)pl->full_tables + 2 * off;
                    gen_tw_full_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(pl->tw_full[t], pl->t_lo, pl->t_hi, pl->log_tlo, (uint32_t)log_a, log_c, cnt);
                    ctx->launches++;
                    if (t == 0 && !(scale == fr_t::one())) {
                        // the plan's scale (n^-1 of an inverse transform) rides in the FIRST boundary's twiddles: no pass of its own
                        scale_table_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, st>>>(pl->tw_full[t], scale, cnt);
                        ctx->launches++;
                    }
                    off += cnt;
                    log_a += pl->radix[t];
                }
            } else {
                cudaGetLastError();  // no room: the two-level tables do the job
                pl->full_tables = nullptr;
            }
        }
    }
    if (cudaGetLastError() != cudaSuccess) return fail(SB_ERR_CUDA);
    // the tables are generated on `st` but a plan is used from any stream of the context afterwards: make them visible once, here
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail(SB_ERR_CUDA);
    ctx->ntt_plans[key] = pl;
    *out = pl;
    return SB_OK;
}

void ntt_plans_free(sb_ctx *ctx) {
    for (auto &kv : ctx->ntt_plans) {
        if (kv.second->tables) cudaFree(kv.second->tables);
        if (kv.second->full_tables) cudaFree(kv.second->full_tables);
        delete kv.second;
    }
    ctx->ntt_plans.clear();
}

static size_t pass_smem(uint32_t r, uint32_t g) { return ((size_t)32 << (r + g)) + ((size_t)16 << r); }

int32_t ntt_run_fused(sb_ctx *ctx, const void *d_in, void *d_out, const uint8_t omega[32], uint32_t log_n, const NttFuse *fuse, cudaStream_t st) {
    SB_REQUIRE(log_n <= 28, "best_fft: log_n > 28 (Fr two-adicity is 28)");
    const uint64_t n = 1ull << log_n;
    NttFuse f0;
    if (!fuse) fuse = &f0;
    SB_REQUIRE(fuse->pre_m <= 8 && fuse->post_m <= 8, "ntt: pattern lengths must be <= 8");
    const uint64_t n_in = fuse->n_in ? fuse->n_in : n, n_out = fuse->n_out ? fuse->n_out : n;
    SB_REQUIRE(n_in <= n && n_out <= n, "ntt: n_in / n_out exceed the transform size");
    const bool has_scale = fuse->has_scale;
    const uint32_t batch = fuse->batch ? fuse->batch : 1;
    SB_REQUIRE(batch <= 65535, "ntt: batch too large");
    if (log_n == 0) {
        SB_REQUIRE(batch == 1, "ntt: batched size-1 transforms are not supported");
        // the identity transform: only the fused scalings remain
        if (d_out != d_in) SB_CUDA_TRY(cudaMemcpyAsync(d_out, d_in, 32, cudaMemcpyDeviceToDevice, st));
        fr_t s = has_scale ? fuse->scale : fr_t::one();
        if (fuse->pre_m) s = mul(s, fuse->pre_pat[0]);
        if (fuse->post_m) s = mul(s, fuse->post_pat[0]);
        SB_REQUIRE(!fuse->pre_vec && !fuse->post_vec, "ntt: vector scalings on a size-1 transform are not supported");
        return (s == fr_t::one()) ? SB_OK : fr_scale(ctx, d_out, 1, s, st);
    }
    NttPlan *pl = nullptr;
    // multi-pass plans carry the scale in their inter-pass twiddles; single-pass plans apply it with the post pattern
    const fr_t one = fr_t::one();
    SB_TRY(plan_get(ctx, omega, log_n, (has_scale && log_n > SMALL_MAX_LOG) ? fuse->scale : one, st, &pl));
    const bool scale_in_post = has_scale && pl->npass == 1;
    static bool attr_set[64] = {false};  // per device
    if (ctx->device >= 64 || !attr_set[ctx->device]) {
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<11, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(11, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<12, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(12, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<11, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(11, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<12, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(12, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)32 << SMALL_MAX_LOG)));
        if (ctx->device < 64) attr_set[ctx->device] = true;
    }
    const size_t bytes = (size_t)32 << log_n;
    // strided passes run in place on a work buffer, the last pass goes work -> d_out.  The input itself is never written unless d_out == d_in.
    void *d_work = nullptr;
    if (pl->npass > 1) {
        // one ping-pong buffer per stream: calls on different streams of a context may be in flight together
        char slot[48];
        snprintf(slot, sizeof slot, "ntt_tmp_%llx", (unsigned long long)(uintptr_t)st);
        SB_TRY(scratch_get(ctx, slot, bytes * batch, &d_work));
    }

    uint32_t log_a = 0;
    for (int t = 0; t < pl->npass; t++) {
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        const bool last = (t == pl->npass - 1);
        a.log_n = log_n;
        a.r = pl->radix[t];
        a.log_a = log_a;
        a.log_c = log_n - log_a - a.r;
        a.log_r1 = pl->radix[0];
        a.log_tlo = pl->log_tlo;
        a.npass = (uint32_t)pl->npass;
        a.kind = pl->npass == 1 ? NTT_SINGLE : (last ? NTT_LAST : NTT_STRIDED);
        a.g = pl->npass == 1 ? 0 : pl->tile_log - a.r;
        if (a.kind == NTT_STRIDED && a.g > a.log_c) a.g = a.log_c;
        if (a.kind == NTT_LAST && a.g > a.log_r1) a.g = a.log_r1;
        a.n_mid = 0;
        for (int m = 1; m < pl->npass - 1; m++) a.mid_bits[a.n_mid++] = pl->radix[m];
        a.w = pl->w_block[t];
        a.tw_full = last ? nullptr : pl->tw_full[t];
        a.t_lo = pl->t_lo;
        a.t_hi = pl->t_hi;
        if (t == 0 && !last && !a.tw_full && !(pl->scale == one)) { a.has_tw_scale = 1; a.tw_scale = pl->scale; }
        a.n_in = n;
        a.n_out = n;
        if (t == 0) {
            a.pre_vec = (const uint4 *)fuse->pre_vec;
            a.pre_m = fuse->pre_m;
            for (uint32_t i = 0; i < fuse->pre_m; i++) a.pre_pat[i] = fuse->pre_pat[i];
            a.n_in = n_in;
        }
        if (last) {
            a.n_out = n_out;
            a.post_vec = (const uint4 *)fuse->post_vec;
            a.post_m = fuse->post_m;
            for (uint32_t i = 0; i < fuse->post_m; i++) a.post_pat[i] = scale_in_post ? mul(fuse->post_pat[i], fuse->scale) : fuse->post_pat[i];
            if (scale_in_post && fuse->post_m == 0) { a.post_m = 1; a.post_pat[0] = fuse->scale; }
        }
        a.pre_bstride = fuse->pre_stride;
        a.post_bstride = fuse->post_stride;
        if (pl->npass == 1) {
            a.src = (const uint4 *)d_in;
            a.dst = (uint4 *)d_out;
            a.src_bstride = fuse->src_stride;
            a.dst_bstride = fuse->dst_stride ? fuse->dst_stride : n;
        } else {
            // first pass reads the caller's input and writes the work buffer at the same indices; middle passes stay in place
            a.src = (const uint4 *)(t == 0 ? d_in : d_work);
            a.dst = (uint4 *)(last ? d_out : d_work);
            a.src_bstride = t == 0 ? fuse->src_stride : n;
            a.dst_bstride = last ? (fuse->dst_stride ? fuse->dst_stride : n) : n;
        }
        if (a.r <= SMALL_MAX_LOG && pl->npass == 1) {
            const unsigned nt = a.r >= 9 ? 256u : (a.r >= 6 ? 64u : 32u);
            SB_LAUNCH(ctx, ntt_small_kernel, dim3(1, batch), nt, (size_t)32 << a.r, st, a);
        } else {
            const uint64_t tiles = 1ull << (log_n - a.r - a.g);
            a.eb = (ctx->tune.ntt_eb == 2 && a.r + a.g >= 4) ? 2 : 3;
            const unsigned threads = 1u << (a.r + a.g - a.eb);
            if (a.eb == 2) {
                if (a.r + a.g == 12) SB_LAUNCH(ctx, (ntt_pass_kernel<12, 2>), dim3((unsigned)tiles, batch), threads, pass_smem(a.r, a.g), st, a);
                else SB_LAUNCH(ctx, (ntt_pass_kernel<11, 2>), dim3((unsigned)tiles, batch), threads, pass_smem(a.r, a.g), st, a);
            } else if (a.r + a.g == 12) SB_LAUNCH(ctx, (ntt_pass_kernel<12, 3>), dim3((unsigned)tiles, batch), threads, pass_smem(a.r, a.g), st, a);
            else SB_LAUNCH(ctx, (ntt_pass_kernel<11, 3>), dim3((unsigned)tiles, batch), threads, pass_smem(a.r, a.g), st, a);
        }
        log_a += a.r;
    }
    return SB_OK;
}

// ---- distributed four-step NTT (the one place an all-to-all belongs) ------------------------------------------------------------
// One transform of a vector that every rank holds in full (the sharded prover's polynomials are replicated), each rank doing 1 / world of both
// passes of the two-pass plan:
//   pass 1  the rank's share of the column tiles (columns c in [rank C / world, (rank + 1) C / world)), twiddles applied, in place in a work buffer;
//   A       all-to-all: the rank receives rows k1 in [rank R1 / world, ...) of everybody's columns (blocks packed / unpacked by 2D device copies);
//   pass 2  the rank's share of the row tiles -> its stripe out[k1 + R1 k2] of the natural-order result;
//   B       all-gather of the stripes: every rank ends with the whole transform, like the local ntt_run.
// The exchanges go through the caller's collective library (sb_comm: NCCL all_to_all / all_gather over NVLink); world = 1 skips them.
int32_t ntt_run_dist(sb_ctx *ctx, const sb_comm *comm, void *d_a, const uint8_t omega[32], uint32_t log_n, const fr_t *scale, cudaStream_t st) {
    const uint32_t W = comm ? (uint32_t)comm->world : 1u, rk = comm ? (uint32_t)comm->rank : 0u;
    const fr_t one = fr_t::one();
    NttFuse f;
    if (scale) { f.has_scale = true; f.scale = *scale; }
    SB_REQUIRE(log_n <= 28, "ntt_dist: log_n > 28");
    if (log_n < 16) return ntt_run_fused(ctx, d_a, d_a, omega, log_n, &f, st);  // too small to be worth two exchanges: every rank transforms it locally
    SB_REQUIRE((W & (W - 1)) == 0 && W <= 8, "ntt_dist: world must be 1, 2, 4 or 8");
    SB_REQUIRE(W == 1 || (comm->alltoall_dev && comm->allgather_dev), "ntt_dist: sb_comm needs alltoall_dev and allgather_dev");
    NttPlan *pl = nullptr;
    SB_TRY(plan_get(ctx, omega, log_n, scale ? *scale : one, st, &pl));
    SB_REQUIRE(pl->npass == 2, "ntt_dist: the size needs a two-pass plan (2^16 .. 2^24)");
    static bool attr_set[64] = {false};
    if (ctx->device >= 64 || !attr_set[ctx->device]) {
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<11, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(11, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<12, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(12, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<11, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(11, 0)));
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel<12, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(12, 0)));
        if (ctx->device < 64) attr_set[ctx->device] = true;
    }
    const uint64_t n = 1ull << log_n;
    const uint32_t r1 = pl->radix[0], r2 = pl->radix[1];
    const uint64_t R1 = 1ull << r1, Cc = 1ull << r2;
    uint8_t *d_work, *d_send, *d_recv;
    char slot[3][48];   // per stream, like the local transform's ping-pong buffer
    snprintf(slot[0], 48, "nttd_work_%llx", (unsigned long long)(uintptr_t)st);
    snprintf(slot[1], 48, "nttd_send_%llx", (unsigned long long)(uintptr_t)st);
    snprintf(slot[2], 48, "nttd_recv_%llx", (unsigned long long)(uintptr_t)st);
    SB_TRY(scratch_get(ctx, slot[0], n * 32, (void **)&d_work));
    SB_TRY(scratch_get(ctx, slot[1], n * 32, (void **)&d_send));   // exchange A: n / W; exchange B: the stripes of all ranks (n)
    SB_TRY(scratch_get(ctx, slot[2], (n / W) * 32, (void **)&d_recv));
    NttPassArgs a;
    auto base_args = [&](int t) {
        memset(&a, 0, sizeof a);
        a.log_n = log_n; a.r = pl->radix[t]; a.log_a = t == 0 ? 0 : r1; a.log_c = log_n - a.log_a - a.r; a.log_r1 = r1; a.log_tlo = pl->log_tlo;
        a.npass = 2; a.kind = t == 0 ? NTT_STRIDED : NTT_LAST;
        a.g = pl->tile_log - a.r;
        if (a.kind == NTT_STRIDED && a.g > a.log_c) a.g = a.log_c;
        if (a.kind == NTT_LAST && a.g > a.log_r1) a.g = a.log_r1;
        a.w = pl->w_block[t]; a.tw_full = t == 0 ? pl->tw_full[0] : nullptr; a.t_lo = pl->t_lo; a.t_hi = pl->t_hi;
        if (t == 0 && !a.tw_full && !(pl->scale == one)) { a.has_tw_scale = 1; a.tw_scale = pl->scale; }
        a.n_in = n; a.n_out = n;
    };
    auto launch = [&](uint64_t tiles_all) -> int32_t {
        SB_REQUIRE(tiles_all % W == 0 && tiles_all / W >= 1, "ntt_dist: world does not divide the tiles of a pass");
        const uint64_t mine = tiles_all / W;
        a.tile0 = (uint32_t)(rk * mine);
        a.eb = 3;
        const unsigned threads = 1u << (a.r + a.g - 3);
        if (a.r + a.g == 12) SB_LAUNCH(ctx, (ntt_pass_kernel<12, 3>), dim3((unsigned)mine, 1), threads, pass_smem(a.r, a.g), st, a);
        else SB_LAUNCH(ctx, (ntt_pass_kernel<11, 3>), dim3((unsigned)mine, 1), threads, pass_smem(a.r, a.g), st, a);
        return SB_OK;
    };
    // ---- pass 1 on my columns
    base_args(0);
    a.src = (const uint4 *)d_a; a.dst = (uint4 *)d_work;
    SB_TRY(launch(1ull << (log_n - a.r - a.g)));
    // ---- exchange A
    const uint64_t rows_w = R1 / W, cols_w = Cc / W, blk = rows_w * cols_w * 32;
    if (W > 1) {
        for (uint32_t q = 0; q < W; q++)   // block for rank q: its rows of my columns
            SB_CUDA_TRY(cudaMemcpy2DAsync(d_send + q * blk, cols_w * 32, d_work + ((uint64_t)q * rows_w * Cc + (uint64_t)rk * cols_w) * 32, Cc * 32, cols_w * 32, rows_w,
                                          cudaMemcpyDeviceToDevice, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
        if (comm->alltoall_dev(comm->user, d_send, d_recv, blk, (void *)st) != 0) { set_last_error("sb_comm.alltoall_dev failed"); return SB_ERR_ARG; }
        for (uint32_t q = 0; q < W; q++)   // from rank q: my rows of its columns
            SB_CUDA_TRY(cudaMemcpy2DAsync(d_work + ((uint64_t)rk * rows_w * Cc + (uint64_t)q * cols_w) * 32, Cc * 32, d_recv + q * blk, cols_w * 32, cols_w * 32, rows_w,
                                          cudaMemcpyDeviceToDevice, st));
    }
    // ---- pass 2 on my rows
    base_args(1);
    a.src = (const uint4 *)d_work; a.dst = (uint4 *)d_a;
    SB_TRY(launch(1ull << (log_n - a.r - a.g)));
    // ---- exchange B: my stripe out[k1 + R1 k2], k1 in my range, to everybody
    if (W > 1) {
        const uint64_t stripe = (n / W) * 32;
        SB_CUDA_TRY(cudaMemcpy2DAsync(d_send + rk * stripe, rows_w * 32, (uint8_t *)d_a + (uint64_t)rk * rows_w * 32, R1 * 32, rows_w * 32, Cc, cudaMemcpyDeviceToDevice, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
        if (comm->allgather_dev(comm->user, d_send, stripe, (void *)st) != 0) { set_last_error("sb_comm.allgather_dev failed"); return SB_ERR_ARG; }
        for (uint32_t q = 0; q < W; q++) {
            if (q == rk) continue;
            SB_CUDA_TRY(cudaMemcpy2DAsync((uint8_t *)d_a + (uint64_t)q * rows_w * 32, R1 * 32, d_send + q * stripe, rows_w * 32, rows_w * 32, Cc, cudaMemcpyDeviceToDevice, st));
        }
    }
    return SB_OK;
}

int32_t ntt_run(sb_ctx *ctx, void *d_a, const uint8_t omega[32], uint32_t log_n, cudaStream_t st) {
    return ntt_run_fused(ctx, d_a, d_a, omega, log_n, nullptr, st);
}

}  // namespace sb
