// Internal interfaces of the create_proof pipeline (prover_kernels.cu, expr.cu, prover.cu).
#pragma once
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "hostfr.h"

namespace sb {

// ---- prover_kernels.cu ------------------------------------------------------------------
int32_t fr_batch_invert(sb_ctx *ctx, void *d_a, size_t n, cudaStream_t st);
int32_t fr_running_product(sb_ctx *ctx, const void *d_a, size_t n_a, const fr_t &init, void *d_z, size_t n_z, cudaStream_t st);
int32_t fr_eval_polys(sb_ctx *ctx, const std::vector<const void *> &polys, const std::vector<fr_t> &xs, size_t n, std::vector<fr_t> &out, cudaStream_t st);
int32_t sort_u256(sb_ctx *ctx, void *d_a, size_t count, size_t capacity_pow2, cudaStream_t st);
int32_t lookup_permute(sb_ctx *ctx, const void *d_in, const void *d_tab, size_t n, size_t u, void *d_a_perm, void *d_s_perm, cudaStream_t st);
int32_t chacha_fr_fill(sb_ctx *ctx, const uint32_t key[8], uint64_t counter0, void *d_out, size_t n, cudaStream_t st);
int32_t fr_axpy(sb_ctx *ctx, void *d_acc, const void *d_p, const fr_t &s, size_t n, bool first, cudaStream_t st);
int32_t fr_sub_head(sb_ctx *ctx, void *d_acc, const fr_t *c, uint32_t k, cudaStream_t st);
// out (+)= sum_m coeffs[m] * polys[m], then out[i] -= head[i] for i < head.size() (<= 4); one pass per 24 polynomials
int32_t fr_lincomb(sb_ctx *ctx, void *d_out, const std::vector<const void *> &polys, const std::vector<fr_t> &coeffs, const std::vector<fr_t> &head, size_t n, bool accumulate,
                   cudaStream_t st);
// out[j] = prod_r (x[j] - roots[r])
int32_t fr_vanish(sb_ctx *ctx, const void *d_x, const std::vector<fr_t> &roots, void *d_out, size_t n, cudaStream_t st);
// acc[j] (+)= scale * f[j] * inv_d[j] * prod_{r in comp} (x[j] - r)
int32_t fr_div_combine(sb_ctx *ctx, void *d_acc, const void *d_f, const void *d_inv_d, const void *d_x, const std::vector<fr_t> &comp, const fr_t &scale, size_t n, bool first,
                       cudaStream_t st, const std::vector<fr_t> *head = nullptr);
int32_t fr_open_quotient(sb_ctx *ctx, void *d_out, const void *d_f, const void *d_inv_d, const fr_t &cst, const fr_t &scale, size_t n, cudaStream_t st);

// sparse cells -> dense columns (device copies of the cell / value arrays; indices validated by the caller)
int32_t scatter_cells(sb_ctx *ctx, void *const *col_ptrs, uint32_t n_cols, const void *d_cells, const void *d_values, size_t n_cells, cudaStream_t st);
int32_t sigma_patch(sb_ctx *ctx, void *const *col_ptrs, uint32_t n_cols, const void *d_cells, size_t n_cells, const void *d_omega_pows, const fr_t *delta_pows, cudaStream_t st);

// out[q * n + i] = sum_s m[q * 8 + s] * slots[s][i]: the per-coset data of the quotient -> h's coefficient pieces
int32_t instance_coset(sb_ctx *ctx, const void *d_l0_coset, const void *d_vals, uint32_t n_vals, size_t n, void *d_out, cudaStream_t st);
int32_t fr_coset_combine(sb_ctx *ctx, const std::vector<const void *> &slots, const fr_t *m, void *d_out, size_t n, cudaStream_t st);

// ---- expr.cu: expression DAGs compiled to a register program, evaluated over whole columns ---
struct Expr;
typedef std::shared_ptr<Expr> ExprP;
struct Expr {
    enum Kind { CONST, COL, NEG, ADD, SUB, MUL } kind;
    fr_t c;        // CONST
    int col, rot;  // COL: index into the column-pointer table, row rotation
    ExprP a, b;
};
ExprP e_const(const fr_t &c);
ExprP e_col(int col, int rot);
ExprP e_neg(ExprP a);
ExprP e_add(ExprP a, ExprP b);
ExprP e_sub(ExprP a, ExprP b);
ExprP e_mul(ExprP a, ExprP b);

struct Program {
    std::vector<uint32_t> code;    // 3 words per instruction: (op << 16 | dst), operand a, operand b
    std::vector<fr_t> consts;
    std::vector<int32_t> inputs;   // (col, rot) pairs
    uint32_t n_slots = 0;
    uint32_t out_slot = 0;
    uint32_t n_mul = 0, n_addsub = 0;
};
// terms folded Horner-style: acc = acc * fold + term (fold == nullptr: a single term, no folding)
Program compile_terms(const std::vector<ExprP> &terms, const fr_t *fold);
// the same program on the host for one row: inputs[i] = value of the (column, rotation) pair i of prog.inputs
fr_t program_eval_host(const Program &prog, const std::vector<fr_t> &inputs);
// out[i] = program(columns at row i) for i < 2^log_n; rotations move by rot << rot_scale_log rows (cyclic).
// col_shift (optional): column c is read at element ((row + rot) mod 2^log_n) << col_shift[c], i.e. with a power-of-two stride
// (one coset of an extended-domain column, the pointer already offset to the coset's first element).
int32_t expr_eval(sb_ctx *ctx, const Program &prog, const std::vector<const void *> &cols, uint32_t log_n, uint32_t rot_scale_log, void *d_out, cudaStream_t st,
                  const std::vector<uint8_t> *col_shift = nullptr);

// ---- expr_jit.cu: the same program as straight-line code compiled by NVRTC (nullptr when NVRTC is unavailable: the interpreter runs instead) ----
struct ExprJit;
ExprJit *expr_jit_compile(const Program &prog, std::string *why);
void expr_jit_free(ExprJit *j);
std::string expr_jit_source(const Program &prog);  // the generated CUDA source (CPU test: it must compile offline with nvcc)
int32_t expr_jit_run(sb_ctx *ctx, const ExprJit *j, const void *d_consts, const std::vector<const void *> &cols, uint32_t log_n, uint32_t rot_scale_log, void *d_out,
                     cudaStream_t st, const std::vector<uint8_t> *col_shift = nullptr);

inline fr_t to_dev(const hfr::Fr &x) { fr_t r; memcpy(r.v, x.v, 32); return r; }
inline hfr::Fr to_host(const fr_t &x) { hfr::Fr r; memcpy(r.v, x.v, 32); return r; }

}  // namespace sb
