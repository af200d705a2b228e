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

}  // namespace sb
