// Batched-tile multi-pass NTT over BN254 Fr for sm_100a.
//
// Replaces halo2_proofs::arithmetic::best_fft (reference call sites reach it through
// EvaluationDomain from zk_prover/src/circuits/utils.rs:75-76,94-102).  Natural order in, natural
// order out, any power-of-two size up to 2^28.
//
// Decomposition (mixed-radix Cooley-Tukey, "four-step" generalised to P passes):
//   N = R_1 R_2 ... R_P, every R_t <= 256.  Before pass t the array is indexed [a][j][c] with
//   a = (k_1 .. k_{t-1}) the digits already transformed, j < R_t the digit transformed now and
//   c < C_t the still-contiguous remainder.  A CTA owns a tile of R_t x G elements (G consecutive
//   c's, 2048 elements = 64 KB of shared memory), runs log2(R_t) decimation-in-frequency levels in
//   shared memory, multiplies by the inter-pass twiddle omega^(A_t c k) and writes row k back.
//   The last pass reads G rows whose leading digit k_1 is consecutive and scatters them to
//   out[k_1 + R_1 (k_2 + R_2 (...)) + A k], i.e. every global access of every pass moves G x 32 B
//   contiguous bytes and each pass is exactly one HBM round trip.
// Twiddles: per-pass omega_R^e table (<= 128 entries, staged in shared memory) for the butterflies;
//   inter-pass factors from a two-level table omega^lo * omega^(hi << h) (2 x 2^(n/2) entries,
//   L2-resident) -- one extra product per element per pass boundary.
// The index algebra is pinned on the CPU by tests/models/ntt_model.py (same plan, same formulas).
#include "common.cuh"

namespace sb {

static const uint32_t TILE_LOG = 11;  // elements per tile (2^11 x 32 B = 64 KB)
static const uint32_t RMAX_LOG = 8;   // largest per-pass radix (multi-pass plans)
static const int NTT_THREADS = 256;
static const int MAX_PASS = 8;

struct NttPassArgs {
    const uint4 *src;
    uint4 *dst;
    const uint4 *w_block;  // omega_R^e, e < R/2 (32 B each)
    const uint4 *t_lo;     // omega^i, i < 2^log_tlo
    const uint4 *t_hi;     // omega^(i << log_tlo)
    uint32_t log_n, log_r, log_g, log_a, log_c, log_r1, log_tlo;
    uint32_t last, npass, n_mid;
    uint32_t mid_bits[MAX_PASS];  // radices of passes 2 .. P-1 (for the last pass' digit reversal)
};

struct NttPlan {
    uint32_t log_n = 0;
    int npass = 0;
    uint32_t radix[MAX_PASS];
    uint4 *w_block[MAX_PASS];
    uint4 *t_lo = nullptr, *t_hi = nullptr;
    uint32_t log_tlo = 0;
    void *tables = nullptr;  // one allocation backing all of the above
};

__device__ __forceinline__ fr_t lds_fr(const uint4 *lo, const uint4 *hi, uint32_t i) {
    uint4 a = lo[i], b = hi[i];
    fr_t r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void sts_fr(uint4 *lo, uint4 *hi, uint32_t i, const fr_t &x) {
    lo[i] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    hi[i] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}

__global__ void __launch_bounds__(NTT_THREADS, 3) ntt_pass_kernel(const NttPassArgs p) {
    extern __shared__ uint4 smem[];
    const uint32_t r = p.log_r, g = p.log_g;
    const uint32_t R = 1u << r, G = 1u << g, tile = R << g, gmask = G - 1;
    uint4 *s_lo = smem, *s_hi = smem + tile;
    uint4 *w_lo = s_hi + tile, *w_hi = w_lo + (R > 1 ? R / 2 : 1);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;

    for (uint32_t e = tid; e < R / 2; e += nt) {
        w_lo[e] = __ldg(p.w_block + 2 * e);
        w_hi[e] = __ldg(p.w_block + 2 * e + 1);
    }

    // ---- tile coordinates ------------------------------------------------------------
    const uint64_t tile_id = blockIdx.x;
    uint64_t base = 0;       // strided / single pass: element (j, gg) at base + (j << log_c) + gg
    uint32_t c0 = 0;         // first column of the tile (strided pass)
    uint64_t k1_0 = 0, rest = 0, rev = 0;
    uint32_t log_rest = 0;
    if (!p.last) {
        const uint32_t log_cg = p.log_c - g;
        const uint64_t a_idx = tile_id >> log_cg;
        c0 = (uint32_t)(tile_id & ((1ull << log_cg) - 1)) << g;
        base = (a_idx << (r + p.log_c)) + c0;
    } else if (p.npass > 1) {
        log_rest = p.log_a - p.log_r1;
        rest = tile_id & ((1ull << log_rest) - 1);
        k1_0 = (tile_id >> log_rest) << g;
        // digit-reverse rest = (k_2 .. k_{P-1}), most significant first -> k_2 + R_2 k_3 + ...
        uint64_t tmp = rest;
        for (int m = (int)p.n_mid - 1; m >= 0; m--) {
            const uint32_t bits = p.mid_bits[m];
            const uint64_t d = tmp & ((1ull << bits) - 1);
            tmp >>= bits;
            uint32_t shift = 0;  // digit m lands above digits 0 .. m-1
            for (int q = 0; q < m; q++) shift += p.mid_bits[q];
            rev |= d << shift;
        }
    }

    // ---- load tile into shared memory (swizzled so both access orders are conflict-free) ---
    if (!p.last || p.npass == 1) {
        for (uint32_t idx = tid; idx < tile; idx += nt) {
            const uint32_t gg = idx & gmask, j = idx >> g;
            const uint4 *q = p.src + 2 * (base + ((uint64_t)j << p.log_c) + gg);
            const uint32_t s = (j << g) | (gg ^ (j & gmask));
            s_lo[s] = q[0];
            s_hi[s] = q[1];
        }
    } else {
        for (uint32_t idx = tid; idx < tile; idx += nt) {
            const uint32_t j = idx & (R - 1), gg = idx >> r;
            const uint4 *q = p.src + 2 * (((((k1_0 + gg) << log_rest) + rest) << r) + j);
            const uint32_t s = (j << g) | (gg ^ (j & gmask));
            s_lo[s] = q[0];
            s_hi[s] = q[1];
        }
    }
    __syncthreads();

    // ---- log2(R) decimation-in-frequency levels ---------------------------------------------
    const uint32_t nbf = tile >> 1;
    for (uint32_t l = 0; l < r; l++) {
        const uint32_t sh = r - 1 - l, h = 1u << sh;
        const bool trivial = (l == r - 1);
        for (uint32_t b = tid; b < nbf; b += nt) {
            const uint32_t gg = b & gmask, q = b >> g;
            const uint32_t i = ((q >> sh) << (sh + 1)) | (q & (h - 1));
            const uint32_t e = (q & (h - 1)) << l;
            const uint32_t s0 = (i << g) | (gg ^ (i & gmask));
            const uint32_t i1 = i + h;
            const uint32_t s1 = (i1 << g) | (gg ^ (i1 & gmask));
            fr_t u = lds_fr(s_lo, s_hi, s0), v = lds_fr(s_lo, s_hi, s1);
            sts_fr(s_lo, s_hi, s0, add(u, v));
            fr_t d = sub(u, v);
            if (!trivial) d = mul(d, lds_fr(w_lo, w_hi, e));
            sts_fr(s_lo, s_hi, s1, d);
        }
        __syncthreads();
    }

    // ---- write back: row i of the tile holds output digit k = bitrev(i) -----------------------
    if (!p.last) {
        const uint32_t lo_mask = (1u << p.log_tlo) - 1;
        for (uint32_t idx = tid; idx < tile; idx += nt) {
            const uint32_t gg = idx & gmask, i = idx >> g;
            const uint32_t k = r ? (__brev(i) >> (32 - r)) : 0;
            fr_t x = lds_fr(s_lo, s_hi, (i << g) | (gg ^ (i & gmask)));
            const uint64_t E = ((uint64_t)(c0 + gg) * k) << p.log_a;  // < N
            if (E != 0) {
                fr_t tw = ldg_fp<FrParams>(p.t_lo + 2 * (E & lo_mask));
                const uint64_t eh = E >> p.log_tlo;
                if (eh) tw = mul(tw, ldg_fp<FrParams>(p.t_hi + 2 * eh));
                x = mul(x, tw);
            }
            store_fp(p.dst + 2 * (base + ((uint64_t)k << p.log_c) + gg), x);
        }
    } else if (p.npass == 1) {
        for (uint32_t idx = tid; idx < tile; idx += nt) {
            const uint32_t k = r ? (__brev(idx) >> (32 - r)) : 0;
            store_fp(p.dst + 2 * (uint64_t)k, lds_fr(s_lo, s_hi, idx));
        }
    } else {
        const uint32_t sh_k = p.log_a - p.log_r1;
        for (uint32_t idx = tid; idx < tile; idx += nt) {
            const uint32_t gg = idx & gmask, i = idx >> g;
            const uint32_t k = r ? (__brev(i) >> (32 - r)) : 0;
            fr_t x = lds_fr(s_lo, s_hi, (i << g) | (gg ^ (i & gmask)));
            const uint64_t o = (k1_0 + gg) + ((rev + ((uint64_t)k << sh_k)) << p.log_r1);
            store_fp(p.dst + 2 * o, x);
        }
    }
}

// out[i] = base^(i * stride)  (i < count), per-thread square-and-multiply; table setup only
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

static void make_radices(uint32_t log_n, int *npass, uint32_t *radix) {
    if (log_n <= TILE_LOG) {
        *npass = 1;
        radix[0] = log_n;
        return;
    }
    uint32_t p = (log_n + RMAX_LOG - 1) / RMAX_LOG;
    uint32_t base = log_n / p, extra = log_n % p;
    *npass = (int)p;
    for (uint32_t t = 0; t < p; t++) radix[t] = base + (t < extra ? 1 : 0);
}

static int32_t plan_get(sb_ctx *ctx, const uint8_t omega[32], uint32_t log_n, cudaStream_t st, NttPlan **out) {
    std::string key((const char *)omega, 32);
    key.push_back((char)log_n);
    auto it = ctx->ntt_plans.find(key);
    if (it != ctx->ntt_plans.end()) {
        *out = it->second;
        return SB_OK;
    }
    NttPlan *pl = new NttPlan();
    pl->log_n = log_n;
    make_radices(log_n, &pl->npass, pl->radix);
    pl->log_tlo = (log_n + 1) / 2;
    const uint64_t n_lo = 1ull << pl->log_tlo, n_hi = 1ull << (log_n - pl->log_tlo);
    uint64_t total = n_lo + n_hi;
    uint64_t off_w[MAX_PASS];
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
    SB_TRY(gen(pl->t_lo, w, n_lo));
    SB_TRY(gen(pl->t_hi, fr_pow_host(w, n_lo), n_hi));
    for (int t = 0; t < pl->npass; t++) {
        pl->w_block[t] = tb + 2 * off_w[t];
        uint64_t cnt = (1ull << pl->radix[t]) / 2;
        if (cnt == 0) cnt = 1;
        // omega_R = omega^(N / R)
        SB_TRY(gen(pl->w_block[t], fr_pow_host(w, 1ull << (log_n - pl->radix[t])), cnt));
    }
    // the tables are generated on `st` but a plan is used from any stream of the context afterwards (create_proof runs its coset NTTs
    // on a side stream): make them visible once, here
    SB_CUDA_TRY(cudaStreamSynchronize(st));
    ctx->ntt_plans[key] = pl;
    *out = pl;
    return SB_OK;
}

void ntt_plans_free(sb_ctx *ctx) {
    for (auto &kv : ctx->ntt_plans) {
        if (kv.second->tables) cudaFree(kv.second->tables);
        delete kv.second;
    }
    ctx->ntt_plans.clear();
}

static size_t pass_smem(uint32_t log_r, uint32_t log_g) {
    size_t tile = (size_t)1 << (log_r + log_g);
    size_t w = ((size_t)1 << log_r) / 2;
    if (w == 0) w = 1;
    return (tile + w) * 32;
}

int32_t ntt_run(sb_ctx *ctx, void *d_a, const uint8_t omega[32], uint32_t log_n, cudaStream_t st) {
    SB_REQUIRE(log_n <= 28, "best_fft: log_n > 28 (Fr two-adicity is 28)");
    if (log_n == 0) return SB_OK;
    NttPlan *pl = nullptr;
    SB_TRY(plan_get(ctx, omega, log_n, st, &pl));
    static bool attr_set[64] = {false};  // per device
    if (ctx->device >= 64 || !attr_set[ctx->device]) {
        SB_CUDA_TRY(cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pass_smem(TILE_LOG, 0)));
        if (ctx->device < 64) attr_set[ctx->device] = true;
    }
    const size_t bytes = (size_t)32 << log_n;
    void *d_tmp = nullptr;
    // the ping-pong buffer is per stream: create_proof runs coset NTTs on the side stream while the main stream transforms other columns
    if (pl->npass > 1) SB_TRY(scratch_get(ctx, st == ctx->side_stream ? "ntt_tmp_side" : "ntt_tmp", bytes, &d_tmp));

    uint32_t log_a = 0;
    for (int t = 0; t < pl->npass; t++) {
        NttPassArgs a;
        memset(&a, 0, sizeof a);
        const bool last = (t == pl->npass - 1);
        a.log_n = log_n;
        a.log_r = pl->radix[t];
        a.log_g = pl->npass == 1 ? 0 : TILE_LOG - a.log_r;
        a.log_a = log_a;
        a.log_c = log_n - log_a - a.log_r;
        a.log_r1 = pl->radix[0];
        a.log_tlo = pl->log_tlo;
        a.last = last ? 1 : 0;
        a.npass = (uint32_t)pl->npass;
        a.n_mid = 0;
        for (int m = 1; m < pl->npass - 1; m++) a.mid_bits[a.n_mid++] = pl->radix[m];
        a.w_block = pl->w_block[t];
        a.t_lo = pl->t_lo;
        a.t_hi = pl->t_hi;
        // ping-pong: first pass d_a -> tmp, middle passes tmp -> tmp (in place), last pass tmp -> d_a
        if (pl->npass == 1) {
            a.src = (const uint4 *)d_a;
            a.dst = (uint4 *)d_a;
        } else {
            a.src = (const uint4 *)(t == 0 ? d_a : d_tmp);
            a.dst = (uint4 *)(last ? d_a : d_tmp);
        }
        const uint64_t tiles = 1ull << (log_n - a.log_r - a.log_g);
        SB_LAUNCH(ctx, ntt_pass_kernel, (unsigned)tiles, NTT_THREADS, pass_smem(a.log_r, a.log_g), st, a);
        log_a += a.log_r;
    }
    return SB_OK;
}

}  // namespace sb
