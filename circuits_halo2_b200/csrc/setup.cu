// Fixed-base scalar multiplication s_i * G over BN254 G1, one thread per scalar.
//
// This is the data-parallel core of halo2 `ParamsKZG::setup` (reference call site
// zk_prover/src/circuits/utils.rs:70: g[i] = [tau^i] G), and what bench.py uses to synthesise
// valid random bases on the device.  A table T[j] = 2^j * G (j < 254, affine, built once on the host)
// turns every product into ~127 mixed additions with no doublings; each thread normalises its own
// result (one Fermat inversion), so the output is the affine halo2curves layout.
#include "common.cuh"
#include "ec.cuh"

namespace sb {

void host_pow2_table(uint8_t *out /* 254 x 64 B */);

__global__ void __launch_bounds__(128) g1_fixed_base_mul_kernel(const uint4 *scalars, uint64_t n, const uint4 *table, uint4 *out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t s = from_mont(load_fp<FrParams>(scalars + 2 * i));
    xyzz_t acc = xyzz_t::identity();
    for (int j = 0; j < 254; j++) {
        if ((s.v[j >> 5] >> (j & 31)) & 1) {
            affine_t t;
            t.x = ldg_fp<FqParams>(table + 4 * j);
            t.y = ldg_fp<FqParams>(table + 4 * j + 2);
            madd(acc, t, false);
        }
    }
    affine_t r = to_affine(acc);
    store_fp(out + 4 * i, r.x);
    store_fp(out + 4 * i + 2, r.y);
}

int32_t g1_fixed_base_mul(sb_ctx *ctx, const void *d_scalars, size_t n, void *d_out, cudaStream_t st) {
    if (n == 0) return SB_OK;
    void *d_table = nullptr;
    const bool fresh = ctx->scratch.find("g1_pow2_table") == ctx->scratch.end();
    SB_TRY(scratch_get(ctx, "g1_pow2_table", 254 * 64, &d_table));
    if (fresh) {
        std::vector<uint8_t> host(254 * 64);
        host_pow2_table(host.data());
        SB_CUDA_TRY(cudaMemcpyAsync(d_table, host.data(), host.size(), cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
    }
    SB_LAUNCH(ctx, g1_fixed_base_mul_kernel, (unsigned)((n + 127) / 128), 128, 0, st, (const uint4 *)d_scalars, (uint64_t)n, (const uint4 *)d_table, (uint4 *)d_out);
    return SB_OK;
}


// ------------------------------------------------------------------ EC-NTT: ParamsKZG::downsize
// halo2 `ParamsKZG::downsize(k)` (zk_prover/src/circuits/utils.rs:62-66) keeps the first 2^k monomial bases and recomputes the Lagrange
// bases as `g_to_lagrange`: an inverse FFT over GROUP elements, g_lagrange[i] = (1/n) sum_j omega^(-i j) g[j].  Radix-2 DIT over XYZZ
// points: bit-reversed load, log n passes of n/2 butterflies (a, b) -> (a + [w] b, a - [w] b) with the twiddle applied by a 254-bit
// double-and-add, then [1/n] and the normalisation to affine.  Off the proving path (once per (SRS, k)).
__device__ __forceinline__ xyzz_t g1_scalar_mul(const xyzz_t &p, const fr_t &s_canonical) {
    xyzz_t acc = xyzz_t::identity();
    if (p.is_identity()) return acc;
    int top = 255;
    while (top >= 0 && !((s_canonical.v[top >> 5] >> (top & 31)) & 1)) top--;
    for (int j = top; j >= 0; j--) {
        acc = dbl(acc);
        if ((s_canonical.v[j >> 5] >> (j & 31)) & 1) add(acc, p);
    }
    return acc;
}
__global__ void __launch_bounds__(128) ecntt_load_kernel(const uint4 *affine_in, uint4 *xyzz_out, uint32_t log_n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_n)) return;
    const uint32_t r = log_n ? (__brev(i) >> (32 - log_n)) : 0;
    affine_t p;
    p.x = load_fp<FqParams>(affine_in + 4 * (uint64_t)i);
    p.y = load_fp<FqParams>(affine_in + 4 * (uint64_t)i + 2);
    const xyzz_t q = xyzz_t::from_affine(p);
    uint4 *o = xyzz_out + 8 * (uint64_t)r;
    store_fp(o, q.x); store_fp(o + 2, q.y); store_fp(o + 4, q.zz); store_fp(o + 6, q.zzz);
}
__device__ __forceinline__ xyzz_t ld_xyzz(const uint4 *p) {
    xyzz_t r;
    r.x = load_fp<FqParams>(p); r.y = load_fp<FqParams>(p + 2); r.zz = load_fp<FqParams>(p + 4); r.zzz = load_fp<FqParams>(p + 6);
    return r;
}
__device__ __forceinline__ void st_xyzz(uint4 *p, const xyzz_t &a) { store_fp(p, a.x); store_fp(p + 2, a.y); store_fp(p + 4, a.zz); store_fp(p + 6, a.zzz); }
// pass s: half = 2^s; twiddle of butterfly j inside a group = w_pows[j * (n >> (s + 1))]  (w_pows[i] = omega_inv^i, Montgomery)
__global__ void __launch_bounds__(128) ecntt_pass_kernel(uint4 *a, const uint4 *w_pows, uint32_t log_n, uint32_t s) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = 1u << log_n, half = 1u << s;
    if (t >= n / 2) return;
    const uint32_t grp = t >> s, j = t & (half - 1);
    const uint64_t i0 = (uint64_t)grp * 2 * half + j, i1 = i0 + half;
    xyzz_t x = ld_xyzz(a + 8 * i0), y = ld_xyzz(a + 8 * i1);
    if (j) {
        const fr_t w = from_mont(load_fp<FrParams>(w_pows + 2 * ((uint64_t)j * (n >> (s + 1)))));
        y = g1_scalar_mul(y, w);
    }
    xyzz_t sum = x, diff = x;
    add(sum, y);
    add(diff, neg(y));
    st_xyzz(a + 8 * i0, sum);
    st_xyzz(a + 8 * i1, diff);
}
__global__ void __launch_bounds__(128) ecntt_finish_kernel(const uint4 *a, fr_t n_inv_mont, uint4 *affine_out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const xyzz_t p = g1_scalar_mul(ld_xyzz(a + 8 * (uint64_t)i), from_mont(n_inv_mont));
    const affine_t r = to_affine(p);
    store_fp(affine_out + 4 * (uint64_t)i, r.x);
    store_fp(affine_out + 4 * (uint64_t)i + 2, r.y);
}
// d_lagrange_out[i] = (1/n) sum_j omega^(-i j) d_g[j]  for n = 2^log_n affine points
int32_t g1_to_lagrange(sb_ctx *ctx, const void *d_g, uint32_t log_n, const fr_t &omega_inv, const fr_t &n_inv, void *d_lagrange_out, cudaStream_t st) {
    const size_t n = (size_t)1 << log_n;
    void *d_pts, *d_w;
    SB_TRY(scratch_get(ctx, "ecntt_pts", n * 128, &d_pts));
    SB_TRY(scratch_get(ctx, "ecntt_w", (n / 2 + 1) * 32, &d_w));
    SB_TRY(fr_gen_powers(ctx, d_w, omega_inv, n / 2 + 1, st));
    SB_LAUNCH(ctx, ecntt_load_kernel, (unsigned)((n + 127) / 128), 128, 0, st, (const uint4 *)d_g, (uint4 *)d_pts, log_n);
    for (uint32_t s = 0; s < log_n; s++)
        SB_LAUNCH(ctx, ecntt_pass_kernel, (unsigned)((n / 2 + 127) / 128), 128, 0, st, (uint4 *)d_pts, (const uint4 *)d_w, log_n, s);
    SB_LAUNCH(ctx, ecntt_finish_kernel, (unsigned)((n + 127) / 128), 128, 0, st, (const uint4 *)d_pts, n_inv, (uint4 *)d_lagrange_out, (uint32_t)n);
    return SB_OK;
}

}  // namespace sb
