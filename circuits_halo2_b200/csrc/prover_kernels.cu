// O(n) / O(n log n) building blocks of halo2 `create_proof` that sit between the MSMs and NTTs
// (kernels K6-K8):
//   fr_batch_invert      halo2 `batch_invert` (zeros stay zero)          -- permutation / lookup denominators
//   fr_running_product   z[0] = init, z[i] = z[i-1] * a[i-1]             -- grand-product columns Z
//   fr_eval_polys        halo2 `eval_polynomial` for many (poly, point)   -- the 35 openings + h(x)
//   fr_sort_canonical    bitonic sort by canonical value (Fr `Ord`)       -- lookup permuted columns
//   lookup_permute       halo2 `permute_expression_pair`
//   fr_axpy / fr_sub_head linear combinations of coefficient vectors      -- SHPLONK numerators
// All operate on device arrays of 32-byte Montgomery elements; grids are sized to the data, one
// contiguous chunk per thread where a recurrence is involved (HBM-streaming, no atomics).
#include "common.cuh"
#include "prover.h"

namespace sb {

// ------------------------------------------------------------------ u32 exclusive scan (flags -> positions)
static const int PSCAN_ITEMS = 8, PSCAN_THREADS = 256, PSCAN_TILE = PSCAN_ITEMS * PSCAN_THREADS;

__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *total) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
        uint32_t wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= (uint32_t)o) wi += t;
        }
        warp_sums[lane] = wi - ws;
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    uint32_t r = incl - v + warp_sums[wid];
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(PSCAN_THREADS) pscan_tile_sums(const uint32_t *in, uint64_t m, uint32_t *tile_sums) {
    __shared__ uint32_t total;
    const uint64_t base = (uint64_t)blockIdx.x * PSCAN_TILE + threadIdx.x * PSCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < PSCAN_ITEMS; q++)
        if (base + q < m) s += in[base + q];
    block_excl_scan_u32(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) pscan_of_tiles(uint32_t *tile_sums, uint64_t ntiles, uint32_t *grand_total) {
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (uint64_t base = 0; base < ntiles; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0;
        uint32_t ex = block_excl_scan_u32(v, &total);
        if (i < ntiles) tile_sums[i] = ex + carry;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}
__global__ void __launch_bounds__(PSCAN_THREADS) pscan_apply(const uint32_t *in, uint64_t m, const uint32_t *tile_sums, uint32_t *out) {
    __shared__ uint32_t total;
    const uint64_t base = (uint64_t)blockIdx.x * PSCAN_TILE + threadIdx.x * PSCAN_ITEMS;
    uint32_t v[PSCAN_ITEMS], s = 0;
#pragma unroll
    for (int q = 0; q < PSCAN_ITEMS; q++) {
        v[q] = base + q < m ? in[base + q] : 0;
        s += v[q];
    }
    uint32_t ex = block_excl_scan_u32(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int q = 0; q < PSCAN_ITEMS; q++) {
        if (base + q < m) out[base + q] = ex;
        ex += v[q];
    }
}
// out[i] = sum_{j<i} in[i]; *d_total = sum of all (d_total: device u32)
static int32_t scan_u32(sb_ctx *ctx, const uint32_t *d_in, uint32_t *d_out, uint64_t m, uint32_t *d_total, cudaStream_t st) {
    const uint64_t ntiles = (m + PSCAN_TILE - 1) / PSCAN_TILE;
    uint32_t *d_tiles;
    SB_TRY(scratch_get(ctx, "pscan_tiles", (ntiles + 4) * 4, (void **)&d_tiles));
    SB_LAUNCH(ctx, pscan_tile_sums, (unsigned)ntiles, PSCAN_THREADS, 0, st, d_in, m, d_tiles);
    SB_LAUNCH(ctx, pscan_of_tiles, 1, 1024, 0, st, d_tiles, ntiles, d_total);
    SB_LAUNCH(ctx, pscan_apply, (unsigned)ntiles, PSCAN_THREADS, 0, st, d_in, m, d_tiles, d_out);
    return SB_OK;
}

// ------------------------------------------------------------------ batch inversion
static const int BINV_CHUNK = 32;
__global__ void __launch_bounds__(128) fr_batch_invert_kernel(uint4 *a, uint4 *tmp, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t lo = t * BINV_CHUNK;
    if (lo >= n) return;
    const uint64_t hi = lo + BINV_CHUNK < n ? lo + BINV_CHUNK : n;
    fr_t acc = fr_t::one();
    for (uint64_t i = lo; i < hi; i++) {
        fr_t x = load_fp<FrParams>(a + 2 * i);
        store_fp(tmp + 2 * i, acc);
        if (!x.is_zero()) acc = mul(acc, x);
    }
    fr_t iv = inv(acc);
    for (uint64_t i = hi; i-- > lo;) {
        fr_t x = load_fp<FrParams>(a + 2 * i);
        if (x.is_zero()) continue;
        fr_t p = load_fp<FrParams>(tmp + 2 * i);
        store_fp(a + 2 * i, mul(iv, p));
        iv = mul(iv, x);
    }
}
// Two-level form for long vectors.  The single kernel above spends 310 products of a Fermat inversion per 32 elements (3 n + 10 n products); here a
// thread owns only 8 elements, strided by 32 so that a warp reads 1 KB rows, multiplies them up (n products), the 8x shorter vector of thread totals
// is inverted by the kernel above, and a second pass rebuilds the prefixes in registers and back-substitutes (3 n products): 4 n + 1.6 n products.
static const int BINV_C1 = 8;
__global__ void __launch_bounds__(128) fr_binv_totals_kernel(const uint4 *a, uint4 *tot, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t base = (t >> 5) * (32 * BINV_C1) + (t & 31);
    if (base >= n) return;
    fr_t acc = fr_t::one();
#pragma unroll
    for (int q = 0; q < BINV_C1; q++) {
        const uint64_t i = base + 32 * q;
        if (i < n) {
            const fr_t x = load_fp<FrParams>(a + 2 * i);
            if (!x.is_zero()) acc = mul(acc, x);
        }
    }
    store_fp(tot + 2 * t, acc);
}
__global__ void __launch_bounds__(128) fr_binv_apply_kernel(uint4 *a, const uint4 *tot_inv, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t base = (t >> 5) * (32 * BINV_C1) + (t & 31);
    if (base >= n) return;
    fr_t pre[BINV_C1];
    fr_t acc = fr_t::one();
#pragma unroll
    for (int q = 0; q < BINV_C1; q++) {
        const uint64_t i = base + 32 * q;
        pre[q] = acc;
        if (i < n) {
            const fr_t x = load_fp<FrParams>(a + 2 * i);
            if (!x.is_zero()) acc = mul(acc, x);
        }
    }
    fr_t iv = load_fp<FrParams>(tot_inv + 2 * t);
#pragma unroll
    for (int q = BINV_C1 - 1; q >= 0; q--) {
        const uint64_t i = base + 32 * q;
        if (i < n) {
            const fr_t x = load_fp<FrParams>(a + 2 * i);
            if (!x.is_zero()) {
                store_fp(a + 2 * i, mul(iv, pre[q]));
                iv = mul(iv, x);
            }
        }
    }
}
int32_t fr_batch_invert(sb_ctx *ctx, void *d_a, size_t n, cudaStream_t st) {
    if (n == 0) return SB_OK;
    if (n >= (1u << 16) && !ctx->tune.no_binv2) {
        const uint64_t warps = (n + 32 * BINV_C1 - 1) / (32 * BINV_C1), t1 = warps * 32;
        void *d_tot, *d_tmp2;
        SB_TRY(scratch_get(ctx, "binv_tot", t1 * 32, &d_tot));
        SB_TRY(scratch_get(ctx, "binv_tmp", t1 * 32, &d_tmp2));
        SB_LAUNCH(ctx, fr_binv_totals_kernel, (unsigned)((t1 + 127) / 128), 128, 0, st, (const uint4 *)d_a, (uint4 *)d_tot, (uint64_t)n);
        const uint64_t t2 = (t1 + BINV_CHUNK - 1) / BINV_CHUNK;
        SB_LAUNCH(ctx, fr_batch_invert_kernel, (unsigned)((t2 + 127) / 128), 128, 0, st, (uint4 *)d_tot, (uint4 *)d_tmp2, t1);
        SB_LAUNCH(ctx, fr_binv_apply_kernel, (unsigned)((t1 + 127) / 128), 128, 0, st, (uint4 *)d_a, (const uint4 *)d_tot, (uint64_t)n);
        return SB_OK;
    }
    void *d_tmp;
    SB_TRY(scratch_get(ctx, "binv_tmp", n * 32, &d_tmp));
    const uint64_t threads = (n + BINV_CHUNK - 1) / BINV_CHUNK;
    SB_LAUNCH(ctx, fr_batch_invert_kernel, (unsigned)((threads + 127) / 128), 128, 0, st, (uint4 *)d_a, (uint4 *)d_tmp, (uint64_t)n);
    return SB_OK;
}

// ------------------------------------------------------------------ running product
static const int RP_CHUNK = 64;
// phase 1: product of a over each chunk of indices [t*C, (t+1)*C) intersected with [0, n_a)
__global__ void __launch_bounds__(128) rp_chunk_products(const uint4 *a, uint64_t n_a, uint4 *chunk_prod, uint64_t n_chunks) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_chunks) return;
    const uint64_t lo = t * RP_CHUNK, hi = lo + RP_CHUNK < n_a ? lo + RP_CHUNK : n_a;
    fr_t acc = fr_t::one();
    for (uint64_t i = lo; i < hi; i++) acc = mul(acc, load_fp<FrParams>(a + 2 * i));
    store_fp(chunk_prod + 2 * t, acc);
}
// phase 2 (single CTA): exclusive multiplicative scan of the chunk products, in place
__global__ void __launch_bounds__(1024) rp_scan_chunks(uint4 *chunk_prod, uint64_t n_chunks) {
    __shared__ uint4 s_lo[1024], s_hi[1024];
    const uint32_t tid = threadIdx.x;
    const uint64_t per = (n_chunks + blockDim.x - 1) / blockDim.x;
    const uint64_t lo = tid * per, hi = lo + per < n_chunks ? lo + per : n_chunks;
    fr_t local = fr_t::one();
    for (uint64_t i = lo; i < hi; i++) local = mul(local, load_fp<FrParams>(chunk_prod + 2 * i));
    fr_t incl = local;
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        s_lo[tid] = make_uint4(incl.v[0], incl.v[1], incl.v[2], incl.v[3]);
        s_hi[tid] = make_uint4(incl.v[4], incl.v[5], incl.v[6], incl.v[7]);
        __syncthreads();
        if (tid >= d) {
            uint4 a = s_lo[tid - d], b = s_hi[tid - d];
            fr_t o;
            o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
            incl = mul(incl, o);
        }
        __syncthreads();
    }
    // exclusive prefix of this thread's segment = inclusive of previous thread
    s_lo[tid] = make_uint4(incl.v[0], incl.v[1], incl.v[2], incl.v[3]);
    s_hi[tid] = make_uint4(incl.v[4], incl.v[5], incl.v[6], incl.v[7]);
    __syncthreads();
    fr_t run = fr_t::one();
    if (tid > 0) {
        uint4 a = s_lo[tid - 1], b = s_hi[tid - 1];
        run.v[0] = a.x; run.v[1] = a.y; run.v[2] = a.z; run.v[3] = a.w; run.v[4] = b.x; run.v[5] = b.y; run.v[6] = b.z; run.v[7] = b.w;
    }
    for (uint64_t i = lo; i < hi; i++) {
        fr_t c = load_fp<FrParams>(chunk_prod + 2 * i);
        store_fp(chunk_prod + 2 * i, run);
        run = mul(run, c);
    }
}
// phase 3: z[i] = init * prefix[chunk] * prod_{j in chunk, j < i} a[j]   for i < n_z
__global__ void __launch_bounds__(128) rp_write(const uint4 *a, uint64_t n_a, const uint4 *chunk_prefix, uint64_t n_chunks, fr_t init, uint4 *z, uint64_t n_z) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_chunks) return;
    const uint64_t lo = t * RP_CHUNK;
    fr_t cur = mul(init, load_fp<FrParams>(chunk_prefix + 2 * t));
    for (uint64_t i = lo; i < lo + RP_CHUNK && i < n_z; i++) {
        store_fp(z + 2 * i, cur);
        if (i < n_a) cur = mul(cur, load_fp<FrParams>(a + 2 * i));
    }
}
int32_t fr_running_product(sb_ctx *ctx, const void *d_a, size_t n_a, const fr_t &init, void *d_z, size_t n_z, cudaStream_t st) {
    if (n_z == 0) return SB_OK;
    const uint64_t n_chunks = (n_z + RP_CHUNK - 1) / RP_CHUNK;
    void *d_cp;
    SB_TRY(scratch_get(ctx, "rp_chunks", n_chunks * 32, &d_cp));
    SB_LAUNCH(ctx, rp_chunk_products, (unsigned)((n_chunks + 127) / 128), 128, 0, st, (const uint4 *)d_a, (uint64_t)n_a, (uint4 *)d_cp, n_chunks);
    SB_LAUNCH(ctx, rp_scan_chunks, 1, 1024, 0, st, (uint4 *)d_cp, n_chunks);
    SB_LAUNCH(ctx, rp_write, (unsigned)((n_chunks + 127) / 128), 128, 0, st, (const uint4 *)d_a, (uint64_t)n_a, (const uint4 *)d_cp, n_chunks, init, (uint4 *)d_z, (uint64_t)n_z);
    return SB_OK;
}

// ------------------------------------------------------------------ polynomial evaluation (many (poly, point) pairs)
static const int EV_CHUNK = 256, EV_THREADS = 128;
struct EvalJob {
    const uint4 *poly;
    fr_t x;
    fr_t x_chunk;  // x^EV_CHUNK
};
__global__ void __launch_bounds__(EV_THREADS) eval_poly_partial(const EvalJob *jobs, uint64_t n, uint4 *partials, uint32_t blocks_per_job) {
    __shared__ uint4 s_lo[EV_THREADS], s_hi[EV_THREADS];
    const EvalJob job = jobs[blockIdx.y];
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t lo = t * EV_CHUNK;
    fr_t acc = fr_t::zero();
    if (lo < n) {
        const uint64_t hi = lo + EV_CHUNK < n ? lo + EV_CHUNK : n;
        for (uint64_t i = hi; i-- > lo;) acc = add(mul(acc, job.x), load_fp<FrParams>(job.poly + 2 * i));
        // times x^(t * EV_CHUNK) = (x^EV_CHUNK)^t
        fr_t p = fr_t::one(), b = job.x_chunk;
        uint64_t e = t;
        while (e) {
            if (e & 1) p = mul(p, b);
            b = sqr(b);
            e >>= 1;
        }
        acc = mul(acc, p);
    }
    const uint32_t tid = threadIdx.x;
    s_lo[tid] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
    s_hi[tid] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
    __syncthreads();
    for (uint32_t d = EV_THREADS / 2; d >= 1; d >>= 1) {
        if (tid < d) {
            uint4 a = s_lo[tid + d], b2 = s_hi[tid + d];
            fr_t o;
            o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b2.x; o.v[5] = b2.y; o.v[6] = b2.z; o.v[7] = b2.w;
            acc = add(acc, o);
            s_lo[tid] = make_uint4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
            s_hi[tid] = make_uint4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
        }
        __syncthreads();
    }
    if (tid == 0) store_fp(partials + 2 * ((uint64_t)blockIdx.y * blocks_per_job + blockIdx.x), acc);
}
__global__ void eval_poly_finish(const uint4 *partials, uint32_t blocks_per_job, uint4 *out) {
    // one thread per job: blocks_per_job is small (n / 32768)
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= gridDim.x * blockDim.x) return;
    fr_t acc = fr_t::zero();
    for (uint32_t b = 0; b < blocks_per_job; b++) acc = add(acc, load_fp<FrParams>(partials + 2 * ((uint64_t)j * blocks_per_job + b)));
    store_fp(out + 2 * j, acc);
}
// host: polys[j] device pointers (n coefficients each), xs[j] points -> out[j] on the host
int32_t fr_eval_polys(sb_ctx *ctx, const std::vector<const void *> &polys, const std::vector<fr_t> &xs, size_t n, std::vector<fr_t> &out, cudaStream_t st) {
    const size_t m = polys.size();
    out.resize(m);
    if (m == 0) return SB_OK;
    std::vector<EvalJob> jobs(m);
    for (size_t j = 0; j < m; j++) {
        jobs[j].poly = (const uint4 *)polys[j];
        jobs[j].x = xs[j];
        jobs[j].x_chunk = fr_pow_host(xs[j], EV_CHUNK);
    }
    const uint64_t threads = (n + EV_CHUNK - 1) / EV_CHUNK;
    const uint32_t bpj = (uint32_t)((threads + EV_THREADS - 1) / EV_THREADS);
    void *d_jobs, *d_part, *d_out;
    SB_TRY(scratch_get(ctx, "ev_jobs", m * sizeof(EvalJob), &d_jobs));
    SB_TRY(scratch_get(ctx, "ev_part", (size_t)m * bpj * 32, &d_part));
    SB_TRY(scratch_get(ctx, "ev_out", m * 32, &d_out));
    SB_CUDA_TRY(cudaMemcpyAsync(d_jobs, jobs.data(), m * sizeof(EvalJob), cudaMemcpyHostToDevice, st));
    SB_LAUNCH(ctx, eval_poly_partial, dim3(bpj, (unsigned)m), EV_THREADS, 0, st, (const EvalJob *)d_jobs, (uint64_t)n, (uint4 *)d_part, bpj);
    SB_LAUNCH(ctx, eval_poly_finish, (unsigned)m, 1, 0, st, (const uint4 *)d_part, bpj, (uint4 *)d_out);
    SB_CUDA_TRY(cudaMemcpyAsync(out.data(), d_out, m * 32, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(sync_stream(ctx, st));
    return SB_OK;
}

// ------------------------------------------------------------------ canonical <-> Montgomery, sort
__global__ void fr_convert_kernel(uint4 *a, uint64_t n, int to_mont_flag) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t x = load_fp<FrParams>(a + 2 * i);
    store_fp(a + 2 * i, to_mont_flag ? to_mont(x) : from_mont(x));
}
__global__ void fill_ones_kernel(uint4 *a, uint64_t from, uint64_t to) {
    uint64_t i = from + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= to) return;
    a[2 * i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    a[2 * i + 1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
}
__device__ __forceinline__ bool less256(const uint4 &alo, const uint4 &ahi, const uint4 &blo, const uint4 &bhi) {
    if (ahi.w != bhi.w) return ahi.w < bhi.w;
    if (ahi.z != bhi.z) return ahi.z < bhi.z;
    if (ahi.y != bhi.y) return ahi.y < bhi.y;
    if (ahi.x != bhi.x) return ahi.x < bhi.x;
    if (alo.w != blo.w) return alo.w < blo.w;
    if (alo.z != blo.z) return alo.z < blo.z;
    if (alo.y != blo.y) return alo.y < blo.y;
    return alo.x < blo.x;
}
__device__ __forceinline__ bool eq256(const uint4 &alo, const uint4 &ahi, const uint4 &blo, const uint4 &bhi) {
    return alo.x == blo.x && alo.y == blo.y && alo.z == blo.z && alo.w == blo.w && ahi.x == bhi.x && ahi.y == bhi.y && ahi.z == bhi.z && ahi.w == bhi.w;
}
// one compare-exchange step of the bitonic network on raw 256-bit little-endian integers
__global__ void bitonic_step_kernel(uint4 *a, uint64_t n, uint64_t j, uint64_t k) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    // t enumerates the pairs: insert a zero bit at position log2(j)
    const uint64_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const uint64_t l = i | j;
    uint4 ilo = a[2 * i], ihi = a[2 * i + 1], llo = a[2 * l], lhi = a[2 * l + 1];
    const bool asc = (i & k) == 0;
    const bool swap = asc ? less256(llo, lhi, ilo, ihi) : less256(ilo, ihi, llo, lhi);
    if (swap) {
        a[2 * i] = llo; a[2 * i + 1] = lhi;
        a[2 * l] = ilo; a[2 * l + 1] = ihi;
    }
}
// in-shared-memory phase: all steps with j < 2 * TILE handled by one CTA on a tile of 2*blockDim elements
static const int BITONIC_SMEM_THREADS = 512;  // tile = 1024 elements = 32 KB
__global__ void __launch_bounds__(BITONIC_SMEM_THREADS) bitonic_smem_kernel(uint4 *a, uint64_t n, uint64_t k, uint64_t j_start) {
    __shared__ uint4 s_lo[2 * BITONIC_SMEM_THREADS], s_hi[2 * BITONIC_SMEM_THREADS];
    const uint32_t tid = threadIdx.x;
    const uint64_t base = (uint64_t)blockIdx.x * (2 * BITONIC_SMEM_THREADS);
    for (uint32_t q = tid; q < 2 * BITONIC_SMEM_THREADS; q += BITONIC_SMEM_THREADS) {
        s_lo[q] = a[2 * (base + q)];
        s_hi[q] = a[2 * (base + q) + 1];
    }
    __syncthreads();
    for (uint64_t j = j_start; j > 0; j >>= 1) {
        const uint32_t i = (uint32_t)(((tid & ~(j - 1)) << 1) | (tid & (j - 1)));
        const uint32_t l = i | (uint32_t)j;
        const bool asc = ((base + i) & k) == 0;
        uint4 ilo = s_lo[i], ihi = s_hi[i], llo = s_lo[l], lhi = s_hi[l];
        const bool swap = asc ? less256(llo, lhi, ilo, ihi) : less256(ilo, ihi, llo, lhi);
        if (swap) {
            s_lo[i] = llo; s_hi[i] = lhi;
            s_lo[l] = ilo; s_hi[l] = ihi;
        }
        __syncthreads();
    }
    for (uint32_t q = tid; q < 2 * BITONIC_SMEM_THREADS; q += BITONIC_SMEM_THREADS) {
        a[2 * (base + q)] = s_lo[q];
        a[2 * (base + q) + 1] = s_hi[q];
    }
}
// ---- small-key fast path: when every key is below 2^16 (range-check inputs, byte tables, selector-gated zeros) the sorted array is
//      fully described by a 65 536-bin histogram: count, scan, and let every output position look its value up in the offsets.
static const uint32_t SMALLKEY_BINS = 1u << 16;
__global__ void smallkey_hist_kernel(const uint4 *a, uint64_t n, uint32_t *hist, uint32_t *too_big) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 lo = a[2 * i], hi = a[2 * i + 1];
    if ((lo.x >> 16) | lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w) {
        *too_big = 1;
        return;
    }
    // warp-aggregated: constant columns put a whole warp into one bin
    const uint32_t peers = __match_any_sync(__activemask(), lo.x);
    if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(hist + lo.x, __popc(peers));
}
__global__ void smallkey_scan_kernel(uint32_t *hist /* in: counts, out: exclusive offsets; [BINS] = total */) {
    __shared__ uint32_t part[1024];
    const uint32_t t = threadIdx.x, per = SMALLKEY_BINS / 1024;
    uint32_t s = 0;
    for (uint32_t q = 0; q < per; q++) s += hist[t * per + q];
    part[t] = s;
    __syncthreads();
    for (uint32_t d = 1; d < 1024; d <<= 1) {
        uint32_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for (uint32_t q = 0; q < per; q++) {
        const uint32_t c = hist[t * per + q];
        hist[t * per + q] = run;
        run += c;
    }
    if (t == 1023) hist[SMALLKEY_BINS] = run;
}
__global__ void smallkey_expand_kernel(const uint32_t *offsets, uint4 *out, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t lo = 0, hi = SMALLKEY_BINS;  // largest v with offsets[v] <= i
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= (uint32_t)i) lo = mid;
        else hi = mid;
    }
    out[2 * i] = make_uint4(lo, 0, 0, 0);
    out[2 * i + 1] = make_uint4(0, 0, 0, 0);
}

// sorts d_a[0 .. count) ascending as raw 256-bit integers; d_a must have room for the next power of two
int32_t sort_u256(sb_ctx *ctx, void *d_a, size_t count, size_t capacity_pow2, cudaStream_t st) {
    if (count >= 4096 && count < (1ull << 32) && !ctx->tune.no_smallkey_sort) {
        uint32_t *d_hist;
        SB_TRY(scratch_get(ctx, "sort_hist", (SMALLKEY_BINS + 2) * 4, (void **)&d_hist));
        SB_CUDA_TRY(cudaMemsetAsync(d_hist, 0, (SMALLKEY_BINS + 2) * 4, st));
        SB_LAUNCH(ctx, smallkey_hist_kernel, (unsigned)((count + 255) / 256), 256, 0, st, (const uint4 *)d_a, (uint64_t)count, d_hist, d_hist + SMALLKEY_BINS + 1);
        uint32_t too_big = 1;
        SB_CUDA_TRY(cudaMemcpyAsync(&too_big, d_hist + SMALLKEY_BINS + 1, 4, cudaMemcpyDeviceToHost, st));
        SB_CUDA_TRY(sync_stream(ctx, st));
        if (!too_big) {
            SB_LAUNCH(ctx, smallkey_scan_kernel, 1, 1024, 0, st, d_hist);
            SB_LAUNCH(ctx, smallkey_expand_kernel, (unsigned)((count + 255) / 256), 256, 0, st, (const uint32_t *)d_hist, (uint4 *)d_a, (uint64_t)count);
            return SB_OK;
        }
    }
    const uint64_t N = capacity_pow2;
    if (N > count) SB_LAUNCH(ctx, fill_ones_kernel, (unsigned)((N - count + 255) / 256), 256, 0, st, (uint4 *)d_a, (uint64_t)count, N);
    if (N < 2) return SB_OK;
    const uint64_t tile = 2 * BITONIC_SMEM_THREADS;
    for (uint64_t k = 2; k <= N; k <<= 1) {
        uint64_t j = k >> 1;
        if (N >= tile) {
            for (; j >= tile; j >>= 1) SB_LAUNCH(ctx, bitonic_step_kernel, (unsigned)((N / 2 + 255) / 256), 256, 0, st, (uint4 *)d_a, N, j, k);
            SB_LAUNCH(ctx, bitonic_smem_kernel, (unsigned)(N / tile), BITONIC_SMEM_THREADS, 0, st, (uint4 *)d_a, N, k, j);
        } else {
            for (; j > 0; j >>= 1) SB_LAUNCH(ctx, bitonic_step_kernel, (unsigned)((N / 2 + 255) / 256), 256, 0, st, (uint4 *)d_a, N, j, k);
        }
    }
    return SB_OK;
}

// ------------------------------------------------------------------ lookup: permute_expression_pair
// a_sorted / t_sorted: canonical, ascending, u entries each.
// rep[r]  = 1 iff row r of the sorted input repeats the previous value (r > 0)
// left[p] = 1 iff sorted table entry p is NOT consumed by a first occurrence of an input value
__global__ void lookup_flags_kernel(const uint4 *a_sorted, const uint4 *t_sorted, uint64_t u, uint32_t *rep, uint32_t *left) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= u) return;
    uint4 alo = a_sorted[2 * i], ahi = a_sorted[2 * i + 1];
    rep[i] = (i > 0 && eq256(alo, ahi, a_sorted[2 * (i - 1)], a_sorted[2 * (i - 1) + 1])) ? 1u : 0u;
    uint4 tlo = t_sorted[2 * i], thi = t_sorted[2 * i + 1];
    bool consumed = false;
    if (i == 0 || !eq256(tlo, thi, t_sorted[2 * (i - 1)], t_sorted[2 * (i - 1) + 1])) {
        // first occurrence of this table value: consumed iff the value occurs among the inputs
        uint64_t lo = 0, hi = u;
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (less256(a_sorted[2 * mid], a_sorted[2 * mid + 1], tlo, thi)) lo = mid + 1;
            else hi = mid;
        }
        consumed = lo < u && eq256(a_sorted[2 * lo], a_sorted[2 * lo + 1], tlo, thi);
    }
    left[i] = consumed ? 0u : 1u;
}
__global__ void lookup_compact_kernel(const uint4 *t_sorted, uint64_t u, const uint32_t *rep, const uint32_t *rep_pos, const uint32_t *left, const uint32_t *left_pos,
                                      uint32_t *rep_rows, uint4 *left_vals) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= u) return;
    if (rep[i]) rep_rows[rep_pos[i]] = (uint32_t)i;
    if (left[i]) {
        left_vals[2 * left_pos[i]] = t_sorted[2 * i];
        left_vals[2 * left_pos[i] + 1] = t_sorted[2 * i + 1];
    }
}
// permuted table: first occurrences take the input value; repeated rows, from the LAST one backwards, take the
// leftover table values in ascending order
__global__ void lookup_fill_kernel(const uint4 *a_sorted, uint64_t u, const uint32_t *rep, const uint32_t *rep_rows, const uint4 *left_vals, uint32_t m, uint4 *s_perm) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < u && !rep[i]) {
        s_perm[2 * i] = a_sorted[2 * i];
        s_perm[2 * i + 1] = a_sorted[2 * i + 1];
    }
    if (i < m) {
        const uint32_t row = rep_rows[m - 1 - i];
        s_perm[2 * row] = left_vals[2 * i];
        s_perm[2 * row + 1] = left_vals[2 * i + 1];
    }
}
// d_in / d_tab: compressed input / table columns (Montgomery, n rows).  Outputs (Montgomery, first u rows
// written): d_a_perm, d_s_perm.  Returns SB_ERR_ARG if an input value is missing from the table.
int32_t lookup_permute(sb_ctx *ctx, const void *d_in, const void *d_tab, size_t n, size_t u, void *d_a_perm, void *d_s_perm, cudaStream_t st) {
    size_t N = 1;
    while (N < u) N <<= 1;
    void *d_as, *d_ts, *d_leftv;
    uint32_t *d_flags;
    SB_TRY(scratch_get(ctx, "lk_as", N * 32, &d_as));
    SB_TRY(scratch_get(ctx, "lk_ts", N * 32, &d_ts));
    SB_TRY(scratch_get(ctx, "lk_leftv", (u + 1) * 32, &d_leftv));
    SB_TRY(scratch_get(ctx, "lk_flags", (5 * (u + 4) + 8) * 4, (void **)&d_flags));
    uint32_t *rep = d_flags, *rep_pos = rep + (u + 4), *left = rep_pos + (u + 4), *left_pos = left + (u + 4), *rep_rows = left_pos + (u + 4), *totals = rep_rows + (u + 4);
    SB_CUDA_TRY(cudaMemcpyAsync(d_as, d_in, u * 32, cudaMemcpyDeviceToDevice, st));
    SB_CUDA_TRY(cudaMemcpyAsync(d_ts, d_tab, u * 32, cudaMemcpyDeviceToDevice, st));
    const unsigned gu = (unsigned)((u + 255) / 256);
    SB_LAUNCH(ctx, fr_convert_kernel, gu, 256, 0, st, (uint4 *)d_as, (uint64_t)u, 0);
    SB_LAUNCH(ctx, fr_convert_kernel, gu, 256, 0, st, (uint4 *)d_ts, (uint64_t)u, 0);
    SB_TRY(sort_u256(ctx, d_as, u, N, st));
    SB_TRY(sort_u256(ctx, d_ts, u, N, st));
    SB_LAUNCH(ctx, lookup_flags_kernel, gu, 256, 0, st, (const uint4 *)d_as, (const uint4 *)d_ts, (uint64_t)u, rep, left);
    SB_TRY(scan_u32(ctx, rep, rep_pos, u, totals, st));
    SB_TRY(scan_u32(ctx, left, left_pos, u, totals + 1, st));
    uint32_t h_tot[2];
    SB_CUDA_TRY(cudaMemcpyAsync(h_tot, totals, 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(sync_stream(ctx, st));
    if (h_tot[0] != h_tot[1]) {
        set_last_error("lookup: an input value is not contained in the table (repeated rows %u, leftover table values %u)", h_tot[0], h_tot[1]);
        return SB_ERR_ARG;
    }
    SB_LAUNCH(ctx, lookup_compact_kernel, gu, 256, 0, st, (const uint4 *)d_ts, (uint64_t)u, rep, rep_pos, left, left_pos, rep_rows, (uint4 *)d_leftv);
    SB_LAUNCH(ctx, lookup_fill_kernel, gu, 256, 0, st, (const uint4 *)d_as, (uint64_t)u, rep, rep_rows, (const uint4 *)d_leftv, h_tot[0], (uint4 *)d_s_perm);
    SB_CUDA_TRY(cudaMemcpyAsync(d_a_perm, d_as, u * 32, cudaMemcpyDeviceToDevice, st));
    SB_LAUNCH(ctx, fr_convert_kernel, gu, 256, 0, st, (uint4 *)d_a_perm, (uint64_t)u, 1);
    SB_LAUNCH(ctx, fr_convert_kernel, gu, 256, 0, st, (uint4 *)d_s_perm, (uint64_t)u, 1);
    (void)n;
    return SB_OK;
}

// ------------------------------------------------------------------ ChaCha20 -> Fr (vanishing argument's random polynomial)
// halo2curves `Fr::random` consumes 8 x next_u64 = 16 keystream words = exactly ONE ChaCha20 block, so coefficient i of the
// random polynomial is block (counter0 + i) of the child stream: counter-mode makes the whole column one parallel kernel.
struct ChaChaKey { uint32_t k[8]; };
__device__ __forceinline__ uint32_t rotl32d(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define SB_CQR(a, b, c, d) a += b; d = rotl32d(d ^ a, 16); c += d; b = rotl32d(b ^ c, 12); a += b; d = rotl32d(d ^ a, 8); c += d; b = rotl32d(b ^ c, 7);
__global__ void chacha_fr_kernel(ChaChaKey key, uint64_t counter0, uint4 *out, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t ctr = counter0 + i;
    uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6], key.k[7],
                         (uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t s[16];
#pragma unroll
    for (int q = 0; q < 16; q++) s[q] = init[q];
    for (int r = 0; r < 10; r++) {
        SB_CQR(s[0], s[4], s[8], s[12]) SB_CQR(s[1], s[5], s[9], s[13]) SB_CQR(s[2], s[6], s[10], s[14]) SB_CQR(s[3], s[7], s[11], s[15])
        SB_CQR(s[0], s[5], s[10], s[15]) SB_CQR(s[1], s[6], s[11], s[12]) SB_CQR(s[2], s[7], s[8], s[13]) SB_CQR(s[3], s[4], s[9], s[14])
    }
    fr_t lo, hi;
#pragma unroll
    for (int q = 0; q < 8; q++) {
        lo.v[q] = s[q] + init[q];
        hi.v[q] = s[8 + q] + init[8 + q];
    }
    // lo, hi are arbitrary 256-bit integers (< 6 r): bring them below r first, the Montgomery product needs operands < 2^254
#pragma unroll
    for (int q = 0; q < 5; q++) {
        fr_t t = lo;
        final_sub<FrParams>(lo.v, t.v);
        t = hi;
        final_sub<FrParams>(hi.v, t.v);
    }
    // from_u512: (lo + hi * 2^256) mod r in Montgomery form = lo * R^2 * R^-1 + hi * R^3 * R^-1
    const fr_t r2 = fr_t::r2();
    const fr_t r3 = mul(r2, r2);
    store_fp(out + 2 * i, add(mul(lo, r2), mul(hi, r3)));
}
#undef SB_CQR
int32_t chacha_fr_fill(sb_ctx *ctx, const uint32_t key[8], uint64_t counter0, void *d_out, size_t n, cudaStream_t st) {
    if (n == 0) return SB_OK;
    ChaChaKey k;
    for (int i = 0; i < 8; i++) k.k[i] = key[i];
    SB_LAUNCH(ctx, chacha_fr_kernel, (unsigned)((n + 255) / 256), 256, 0, st, k, counter0, (uint4 *)d_out, (uint64_t)n);
    return SB_OK;
}

// ------------------------------------------------------------------ linear combinations
__global__ void fr_axpy_kernel(uint4 *acc, const uint4 *p, fr_t s, uint64_t n, int first) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t v = mul(load_fp<FrParams>(p + 2 * i), s);
    if (!first) v = add(v, load_fp<FrParams>(acc + 2 * i));
    store_fp(acc + 2 * i, v);
}
// acc = (first ? 0 : acc) + s * p
int32_t fr_axpy(sb_ctx *ctx, void *d_acc, const void *d_p, const fr_t &s, size_t n, bool first, cudaStream_t st) {
    if (n == 0) return SB_OK;
    SB_LAUNCH(ctx, fr_axpy_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (uint4 *)d_acc, (const uint4 *)d_p, s, (uint64_t)n, first ? 1 : 0);
    return SB_OK;
}
struct HeadArgs {
    fr_t c[4];
    uint32_t k;
};
__global__ void fr_sub_head_kernel(uint4 *acc, HeadArgs h) {
    uint32_t i = threadIdx.x;
    if (i >= h.k) return;
    store_fp(acc + 2 * i, sub(load_fp<FrParams>(acc + 2 * i), h.c[i]));
}
// acc[i] -= c[i] for i < k <= 4
int32_t fr_sub_head(sb_ctx *ctx, void *d_acc, const fr_t *c, uint32_t k, cudaStream_t st) {
    if (k == 0) return SB_OK;
    SB_REQUIRE(k <= 4, "fr_sub_head: k > 4");
    HeadArgs h;
    for (uint32_t i = 0; i < 4; i++) h.c[i] = i < k ? c[i] : fr_t::zero();
    h.k = k;
    SB_LAUNCH(ctx, fr_sub_head_kernel, 1, 4, 0, st, (uint4 *)d_acc, h);
    return SB_OK;
}


// ------------------------------------------------------------------ SHPLONK helpers
// out = sum_m coeff[m] * poly[m] over up to LINCOMB_MAX coefficient vectors in ONE pass (each input read once, one write),
// minus up to 4 leading coefficients (the low-degree interpolant r(X) of the rotation set).
static const int LINCOMB_MAX = 24;
struct LinCombArgs {
    const uint4 *poly[LINCOMB_MAX];
    fr_t coeff[LINCOMB_MAX];
    fr_t head[4];
    uint32_t m, n_head, accumulate;
};
__global__ void fr_lincomb_kernel(uint4 *out, const LinCombArgs a, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t acc = a.accumulate ? load_fp<FrParams>(out + 2 * i) : fr_t::zero();
    for (uint32_t m = 0; m < a.m; m++) acc = add(acc, mul(load_fp<FrParams>(a.poly[m] + 2 * i), a.coeff[m]));
    if (i < a.n_head) acc = sub(acc, a.head[i]);
    store_fp(out + 2 * i, acc);
}
int32_t fr_lincomb(sb_ctx *ctx, void *d_out, const std::vector<const void *> &polys, const std::vector<fr_t> &coeffs, const std::vector<fr_t> &head, size_t n, bool accumulate,
                   cudaStream_t st) {
    SB_REQUIRE(polys.size() == coeffs.size() && head.size() <= 4, "fr_lincomb: bad arguments");
    size_t done = 0;
    bool acc = accumulate;
    do {
        LinCombArgs a;
        const size_t take = std::min((size_t)LINCOMB_MAX, polys.size() - done);
        for (size_t m = 0; m < take; m++) { a.poly[m] = (const uint4 *)polys[done + m]; a.coeff[m] = coeffs[done + m]; }
        for (size_t m = take; m < (size_t)LINCOMB_MAX; m++) { a.poly[m] = nullptr; a.coeff[m] = fr_t::zero(); }
        a.m = (uint32_t)take;
        const bool last = done + take == polys.size();
        a.n_head = last ? (uint32_t)head.size() : 0;
        for (size_t h = 0; h < 4; h++) a.head[h] = h < head.size() ? head[h] : fr_t::zero();
        a.accumulate = acc ? 1 : 0;
        SB_LAUNCH(ctx, fr_lincomb_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (uint4 *)d_out, a, (uint64_t)n);
        done += take;
        acc = true;
    } while (done < polys.size());
    return SB_OK;
}

// out[j] = prod_r (x[j] - root[r])   (vanishing polynomial of a rotation set on the division coset)
struct RootArgs {
    fr_t r[8];
    uint32_t k;
};
__global__ void fr_vanish_kernel(const uint4 *x, RootArgs ra, uint4 *out, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t xv = load_fp<FrParams>(x + 2 * i);
    fr_t p = sub(xv, ra.r[0]);
    for (uint32_t q = 1; q < ra.k; q++) p = mul(p, sub(xv, ra.r[q]));
    store_fp(out + 2 * i, p);
}
int32_t fr_vanish(sb_ctx *ctx, const void *d_x, const std::vector<fr_t> &roots, void *d_out, size_t n, cudaStream_t st) {
    SB_REQUIRE(!roots.empty() && roots.size() <= 8, "fr_vanish: 1..8 roots");
    RootArgs ra;
    for (size_t q = 0; q < 8; q++) ra.r[q] = q < roots.size() ? roots[q] : fr_t::zero();
    ra.k = (uint32_t)roots.size();
    SB_LAUNCH(ctx, fr_vanish_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (const uint4 *)d_x, ra, (uint4 *)d_out, (uint64_t)n);
    return SB_OK;
}

// acc[j] (+)= scale * f[j] * inv_d[j] * prod_{r in comp} (x[j] - r):  f / Z_S on the coset, with 1 / Z_S = (1 / Z_T) * prod over the
// points of the super set T that are NOT in S (one batch inversion of Z_T serves every rotation set)
// `head` (evaluation-domain form of the sum): f is first reduced by the low-degree interpolant r(x[j]) = sum_q head.r[q] x[j]^q, which the
// coefficient-domain form subtracts from the leading coefficients instead
__global__ void fr_div_combine_kernel(uint4 *acc, const uint4 *f, const uint4 *inv_d, const uint4 *x, RootArgs comp, RootArgs head, fr_t scale, uint64_t n, int first) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t fv = load_fp<FrParams>(f + 2 * i);
    fr_t xv = fr_t::zero();
    if (comp.k || head.k) xv = load_fp<FrParams>(x + 2 * i);
    if (head.k) {
        fr_t r = head.r[head.k - 1];
        for (uint32_t q = head.k - 1; q-- > 0;) r = add(mul(r, xv), head.r[q]);
        fv = sub(fv, r);
    }
    fr_t v = mul(mul(fv, load_fp<FrParams>(inv_d + 2 * i)), scale);
    for (uint32_t q = 0; q < comp.k; q++) v = mul(v, sub(xv, comp.r[q]));
    if (!first) v = add(v, load_fp<FrParams>(acc + 2 * i));
    store_fp(acc + 2 * i, v);
}
int32_t fr_div_combine(sb_ctx *ctx, void *d_acc, const void *d_f, const void *d_inv_d, const void *d_x, const std::vector<fr_t> &comp, const fr_t &scale, size_t n, bool first,
                       cudaStream_t st, const std::vector<fr_t> *head) {
    SB_REQUIRE(comp.size() <= 8 && (!head || head->size() <= 8), "fr_div_combine: at most 8 complement roots / interpolant coefficients");
    RootArgs ra, ha;
    for (size_t q = 0; q < 8; q++) ra.r[q] = q < comp.size() ? comp[q] : fr_t::zero();
    ra.k = (uint32_t)comp.size();
    for (size_t q = 0; q < 8; q++) ha.r[q] = head && q < head->size() ? (*head)[q] : fr_t::zero();
    ha.k = head ? (uint32_t)head->size() : 0u;
    SB_LAUNCH(ctx, fr_div_combine_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (uint4 *)d_acc, (const uint4 *)d_f, (const uint4 *)d_inv_d, (const uint4 *)d_x, ra, ha, scale,
              (uint64_t)n, first ? 1 : 0);
    return SB_OK;
}

// out[j] = (f[j] - cst) * inv_d[j] * scale: the opening quotient (L(X) - L(u)) / (X - u) in the evaluation domain (inv_d[j] = 1 / (x[j] - u))
__global__ void fr_open_quotient_kernel(uint4 *out, const uint4 *f, const uint4 *inv_d, fr_t cst, fr_t scale, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const fr_t v = mul(sub(load_fp<FrParams>(f + 2 * i), cst), mul(load_fp<FrParams>(inv_d + 2 * i), scale));
    store_fp(out + 2 * i, v);
}
int32_t fr_open_quotient(sb_ctx *ctx, void *d_out, const void *d_f, const void *d_inv_d, const fr_t &cst, const fr_t &scale, size_t n, cudaStream_t st) {
    SB_LAUNCH(ctx, fr_open_quotient_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (uint4 *)d_out, (const uint4 *)d_f, (const uint4 *)d_inv_d, cst, scale, (uint64_t)n);
    return SB_OK;
}

// ---- sparse cells -> dense columns (keygen's fixed cells, the witness' advice cells, the permutation's moved cells) ----
// cells: (col, row) pairs; column c of the output starts at cols.p[c]; one 32-byte value per cell
struct ColPtrs { uint4 *p[32]; };
__global__ void scatter_cells_kernel(ColPtrs cols, const uint32_t *cells, const uint4 *values, uint64_t n_cells) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    uint4 *dst = cols.p[cells[2 * i]] + 2 * (uint64_t)cells[2 * i + 1];
    dst[0] = values[2 * i];
    dst[1] = values[2 * i + 1];
}
// sigma_col(omega^row) = delta^to_col * omega^to_row for the cells copy constraints moved: (col, row, to_col, to_row) quadruples
struct DeltaPows { fr_t d[16]; };
__global__ void sigma_patch_kernel(ColPtrs cols, const uint32_t *cells, uint64_t n_cells, const uint4 *omega_pows, DeltaPows dp) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const uint32_t c = cells[4 * i], row = cells[4 * i + 1], tc = cells[4 * i + 2], trow = cells[4 * i + 3];
    store_fp(cols.p[c] + 2 * (uint64_t)row, mul(dp.d[tc], load_fp<FrParams>(omega_pows + 2 * (uint64_t)trow)));
}
// host validates the indices (column < n_cols <= 32, row < n) before the launch; d_cells / d_values are device copies
int32_t scatter_cells(sb_ctx *ctx, void *const *col_ptrs, uint32_t n_cols, const void *d_cells, const void *d_values, size_t n_cells, cudaStream_t st) {
    if (n_cells == 0) return SB_OK;
    SB_REQUIRE(n_cols <= 32, "scatter_cells: more than 32 columns");
    ColPtrs cp;
    for (uint32_t c = 0; c < 32; c++) cp.p[c] = c < n_cols ? (uint4 *)col_ptrs[c] : nullptr;
    SB_LAUNCH(ctx, scatter_cells_kernel, (unsigned)((n_cells + 255) / 256), 256, 0, st, cp, (const uint32_t *)d_cells, (const uint4 *)d_values, (uint64_t)n_cells);
    return SB_OK;
}
int32_t sigma_patch(sb_ctx *ctx, void *const *col_ptrs, uint32_t n_cols, const void *d_cells, size_t n_cells, const void *d_omega_pows, const fr_t *delta_pows, cudaStream_t st) {
    if (n_cells == 0) return SB_OK;
    SB_REQUIRE(n_cols <= 16, "sigma_patch: more than 16 permutation columns");
    ColPtrs cp;
    DeltaPows dp;
    for (uint32_t c = 0; c < 32; c++) cp.p[c] = c < n_cols ? (uint4 *)col_ptrs[c] : nullptr;
    for (uint32_t c = 0; c < 16; c++) dp.d[c] = c < n_cols ? delta_pows[c] : fr_t::zero();
    SB_LAUNCH(ctx, sigma_patch_kernel, (unsigned)((n_cells + 255) / 256), 256, 0, st, cp, (const uint32_t *)d_cells, (uint64_t)n_cells, (const uint4 *)d_omega_pows, dp);
    return SB_OK;
}

// ---- quotient: per-coset data -> coefficients of h.  out[q * n + i] = sum_s m[q][s] * d_s[i]  (q, s < n_cos <= 8; m = inverse Vandermonde of the
// cosets' g_s^n with 1 / t(g_s) folded in, sb_pk::combine) ----
struct CombineArgs {
    const uint4 *d[8];
    fr_t m[64];
    uint32_t n_cos;
};
__global__ void __launch_bounds__(128) fr_coset_combine_kernel(const CombineArgs a, uint4 *out, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t d[8];
    for (uint32_t s = 0; s < a.n_cos; s++) d[s] = load_fp<FrParams>(a.d[s] + 2 * i);
    for (uint32_t q = 0; q < a.n_cos; q++) {
        fr_t acc = mul(a.m[q * 8], d[0]);
        for (uint32_t s = 1; s < a.n_cos; s++) acc = add(acc, mul(a.m[q * 8 + s], d[s]));
        store_fp(out + 2 * ((uint64_t)q * n + i), acc);
    }
}
int32_t fr_coset_combine(sb_ctx *ctx, const std::vector<const void *> &slots, const fr_t *m, void *d_out, size_t n, cudaStream_t st) {
    SB_REQUIRE(slots.size() >= 1 && slots.size() <= 8, "fr_coset_combine: 1..8 cosets");
    CombineArgs a;
    a.n_cos = (uint32_t)slots.size();
    for (uint32_t s = 0; s < 8; s++) a.d[s] = s < a.n_cos ? (const uint4 *)slots[s] : nullptr;
    for (int i = 0; i < 64; i++) a.m[i] = m[i];
    SB_LAUNCH(ctx, fr_coset_combine_kernel, (unsigned)((n + 127) / 128), 128, 0, st, a, (uint4 *)d_out, (uint64_t)n);
    return SB_OK;
}

// ---- instance column on a coset without a transform.  The column is sum_i v_i L_i(X) with a handful of values, and L_i(X) = L_0(omega^-i X): on
// the coset g H its value at g omega^j is sum_i v_i * l0_coset[(j - i) mod n], a few products per row against the key's l_0 coset (the transform
// pair it replaces -- lagrange_to_coeff + one size-n NTT per coset -- costs 1.4 ms at k = 20) ----
__global__ void __launch_bounds__(128) instance_coset_kernel(const uint4 *l0_coset, const uint4 *vals, uint32_t n_vals, uint64_t n, uint4 *out) {
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    fr_t acc = fr_t::zero();
    for (uint32_t i = 0; i < n_vals; i++) {
        const fr_t v = ldg_fp<FrParams>(vals + 2 * i);
        const fr_t l = load_fp<FrParams>(l0_coset + 2 * ((j - i) & (n - 1)));
        acc = add(acc, mul(v, l));
    }
    store_fp(out + 2 * j, acc);
}
int32_t instance_coset(sb_ctx *ctx, const void *d_l0_coset, const void *d_vals, uint32_t n_vals, size_t n, void *d_out, cudaStream_t st) {
    SB_LAUNCH(ctx, instance_coset_kernel, (unsigned)((n + 127) / 128), 128, 0, st, (const uint4 *)d_l0_coset, (const uint4 *)d_vals, n_vals, (uint64_t)n, (uint4 *)d_out);
    return SB_OK;
}

}  // namespace sb
