// Element-wise Fr / Fq vector kernels: the O(n) glue of halo2 `EvaluationDomain`
// (ifft divisor, zeta-coset scaling + zero padding of `coeff_to_extended`, the 2^(ext_k-k)-periodic
// `divide_by_vanishing_poly`) and the field-op parity probes.  HBM-bound streaming:
// one 32 B element per thread, 128-bit accesses, grid sized to the array.
#include "common.cuh"

namespace sb {

struct PatArgs {
    fr_t pat[8];
    uint32_t m;
};

__global__ void fr_scale_pattern_kernel(uint4 *a, uint64_t n, const PatArgs pa) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t x = load_fp<FrParams>(a + 2 * i);
    const uint32_t sel = (uint32_t)(i % pa.m);
    fr_t s = pa.pat[0];
#pragma unroll
    for (int q = 1; q < 8; q++)
        if (sel == (uint32_t)q) s = pa.pat[q];
    store_fp(a + 2 * i, mul(x, s));
}

__global__ void fr_scale_pattern_pad_kernel(const uint4 *src, uint64_t n_src, uint4 *dst, uint64_t n_dst, const PatArgs pa) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n_dst) return;
    fr_t x = fr_t::zero();
    if (i < n_src) {
        x = load_fp<FrParams>(src + 2 * i);
        const uint32_t sel = (uint32_t)(i % pa.m);
        fr_t s = pa.pat[0];
#pragma unroll
        for (int q = 1; q < 8; q++)
            if (sel == (uint32_t)q) s = pa.pat[q];
        x = mul(x, s);
    }
    store_fp(dst + 2 * i, x);
}

template <class P>
__global__ void fp_vec_op_kernel(const uint4 *a, const uint4 *b, uint4 *out, uint64_t n, int op) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<P> x = load_fp<P>(a + 2 * i), y = load_fp<P>(b + 2 * i), z;
    if (op == 0) z = mul(x, y);
    else if (op == 1) z = add(x, y);
    else z = sub(x, y);
    store_fp(out + 2 * i, z);
}

int32_t fr_scale_pattern(sb_ctx *ctx, void *d_a, size_t n, const fr_t *pat, uint32_t m, cudaStream_t st) {
    SB_REQUIRE(m >= 1 && m <= 8, "scale pattern length must be 1..8");
    if (n == 0) return SB_OK;
    PatArgs pa;
    for (uint32_t i = 0; i < 8; i++) pa.pat[i] = pat[i < m ? i : 0];
    pa.m = m;
    SB_LAUNCH(ctx, fr_scale_pattern_kernel, (unsigned)((n + 255) / 256), 256, 0, st, (uint4 *)d_a, (uint64_t)n, pa);
    return SB_OK;
}

int32_t fr_scale(sb_ctx *ctx, void *d_a, size_t n, const fr_t &s, cudaStream_t st) { return fr_scale_pattern(ctx, d_a, n, &s, 1, st); }

int32_t fr_scale_pattern_pad(sb_ctx *ctx, const void *d_src, size_t n_src, void *d_dst, size_t n_dst, const fr_t *pat, uint32_t m, cudaStream_t st) {
    SB_REQUIRE(m >= 1 && m <= 8, "scale pattern length must be 1..8");
    if (n_dst == 0) return SB_OK;
    PatArgs pa;
    for (uint32_t i = 0; i < 8; i++) pa.pat[i] = pat[i < m ? i : 0];
    pa.m = m;
    SB_LAUNCH(ctx, fr_scale_pattern_pad_kernel, (unsigned)((n_dst + 255) / 256), 256, 0, st, (const uint4 *)d_src, (uint64_t)n_src, (uint4 *)d_dst, (uint64_t)n_dst, pa);
    return SB_OK;
}

// out[(r << shift) + j] = in[j * n + r] * pat[j]   (coset-major -> extended order; pat = t_inv of divide_by_vanishing_poly)
__global__ void fr_coset_interleave_scale_kernel(const uint4 *in, uint4 *out, uint32_t log_n, uint32_t shift, const PatArgs pa) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;  // output index
    if (i >= (1ull << (log_n + shift))) return;
    const uint32_t j = (uint32_t)(i & ((1u << shift) - 1));
    const uint64_t r = i >> shift;
    fr_t s = pa.pat[0];
#pragma unroll
    for (int q = 1; q < 8; q++)
        if (j == (uint32_t)q) s = pa.pat[q];
    store_fp(out + 2 * i, mul(load_fp<FrParams>(in + 2 * (((uint64_t)j << log_n) + r)), s));
}
int32_t fr_coset_interleave_scale(sb_ctx *ctx, const void *d_in, void *d_out, uint32_t log_n, uint32_t shift, const fr_t *pat, cudaStream_t st) {
    SB_REQUIRE(shift <= 3, "coset interleave: at most 8 cosets");
    PatArgs pa;
    for (uint32_t i = 0; i < 8; i++) pa.pat[i] = pat[i < (1u << shift) ? i : 0];
    pa.m = 1u << shift;
    const uint64_t total = 1ull << (log_n + shift);
    SB_LAUNCH(ctx, fr_coset_interleave_scale_kernel, (unsigned)((total + 255) / 256), 256, 0, st, (const uint4 *)d_in, (uint4 *)d_out, log_n, shift, pa);
    return SB_OK;
}

int32_t fp_vec_op(sb_ctx *ctx, int field, int op, const void *d_a, const void *d_b, void *d_out, size_t n, cudaStream_t st) {
    SB_REQUIRE(op >= 0 && op <= 2, "vec op must be 0 (mul), 1 (add) or 2 (sub)");
    if (n == 0) return SB_OK;
    unsigned grid = (unsigned)((n + 255) / 256);
    if (field == 0) SB_LAUNCH(ctx, fp_vec_op_kernel<FrParams>, grid, 256, 0, st, (const uint4 *)d_a, (const uint4 *)d_b, (uint4 *)d_out, (uint64_t)n, op);
    else SB_LAUNCH(ctx, fp_vec_op_kernel<FqParams>, grid, 256, 0, st, (const uint4 *)d_a, (const uint4 *)d_b, (uint4 *)d_out, (uint64_t)n, op);
    return SB_OK;
}

}  // namespace sb
