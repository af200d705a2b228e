// Opaque handle layouts shared by api.cu and prover.cu.
#pragma once
#include "common.cuh"

struct sb_srs {
    std::atomic<int> refs{1};  // a proving key holds a reference to its SRS: sb_srs_destroy before sb_pk_destroy is legal
    uint32_t k = 0;
    void *d_g = nullptr;
    void *d_g_lagrange = nullptr;
    bool borrowed = false;  // sb_srs_wrap_dev: the caller owns the device arrays
    // fixed-base window tables per basis (sb_srs_precompute): [0] monomial, [1] Lagrange; d_tables == nullptr -> plain Pippenger
    sb::MsmTables tab[2];
    void *tab_slab = nullptr;  // both tables in one allocation (monomial first): lets one launch set commit over both bases
    std::mutex tab_mu;  // two keys created concurrently over one SRS build its tables once
};

namespace sb {
int32_t srs_precompute_impl(sb_ctx *ctx, sb_srs *srs, int32_t basis_mask, uint32_t window_bits);  // no locking
// commit through the SRS handle: the table path when the basis has been precomputed, plain Pippenger otherwise
inline int32_t srs_msm(sb_ctx *ctx, const sb_srs *srs, int basis, const void *d_scalars, size_t n, uint8_t out_affine[64], cudaStream_t st) {
    if (srs->tab[basis].d_tables) return msm_run_tables(ctx, &srs->tab[basis], d_scalars, n, 0, -1, out_affine, st);
    return msm_run(ctx, basis == 0 ? srs->d_g : srs->d_g_lagrange, d_scalars, n, out_affine, st);
}
// m scalar vectors (contiguous, n each) against one basis: one launch set with tables, a loop otherwise; out = m x 64 B
inline int32_t srs_msm_batch(sb_ctx *ctx, const sb_srs *srs, int basis, const void *d_scalars, size_t n, uint32_t m, uint8_t *out_affine, cudaStream_t st) {
    if (srs->tab[basis].d_tables) return msm_run_tables_batch(ctx, &srs->tab[basis], d_scalars, n, m, out_affine, st);
    for (uint32_t j = 0; j < m; j++) {
        int32_t rc = msm_run(ctx, basis == 0 ? srs->d_g : srs->d_g_lagrange, (const uint8_t *)d_scalars + (size_t)j * n * 32, n, out_affine + (size_t)j * 64, st);
        if (rc != SB_OK) return rc;
    }
    return SB_OK;
}
}

struct sb_domain {
    uint32_t j = 0, k = 0, ext_k = 0, quotient_degree = 0;
    sb::fr_t omega, omega_inv, ext_omega, ext_omega_inv;
    sb::fr_t ifft_divisor, ext_ifft_divisor;
    sb::fr_t coset[3], coset_inv[3];  // zeta^(i mod 3), zeta^-(i mod 3)
    sb::fr_t t_inv[8];
    uint32_t n_t = 0;
};


namespace sb {
// EvaluationDomain transforms on device pointers, no locking (callers hold the context)
int32_t dom_l2c(sb_ctx *ctx, const sb_domain *d, void *d_a, cudaStream_t st);
int32_t dom_c2e(sb_ctx *ctx, const sb_domain *d, const void *d_coeff, void *d_ext, cudaStream_t st);
int32_t dom_e2c(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *d_coeff, cudaStream_t st);
int32_t dom_div_vanishing(sb_ctx *ctx, const sb_domain *d, void *d_ext, cudaStream_t st);
int32_t dom_div_e2c(sb_ctx *ctx, const sb_domain *d, void *d_ext, void *d_coeff, cudaStream_t st);  // divide_by_vanishing_poly + extended_to_coeff in one transform

struct CtxGuard {
    std::lock_guard<std::mutex> lk;
    int prev = -1;
    explicit CtxGuard(sb_ctx *c) : lk(c->mu) {
        cudaGetDevice(&prev);
        if (prev != c->device) cudaSetDevice(c->device);
        else prev = -1;
    }
    ~CtxGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace sb
