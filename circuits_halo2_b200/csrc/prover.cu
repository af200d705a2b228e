// halo2 `create_proof` (KZG commitments, SHPLONK multi-open) on one B200.
//
// Replaces halo2_proofs::plonk::{keygen_pk (the device-resident part), create_proof}
// (reference call sites zk_prover/src/circuits/utils.rs:75-76, :94-102 Blake2b transcript,
// :171-178 Keccak transcript).  The sequence of transcript writes, challenge squeezes and RNG draws is
// the one SURVEY.md A.5-A.10/A.13 restates; the oracle twin is oracle/halo2_prover.py and the two
// produce byte-identical proofs for the same seed.
//
// Division of labour: everything that touches n or 8n field elements runs on the GPU (MSM, NTT,
// expression programs over columns, batch inversion, running products, sort, Horner sums, SHPLONK
// numerators and their exact division by Z_S(X) via a coset NTT); the host keeps the Fiat-Shamir
// transcript, the ChaCha20 blinding RNG, and O(1) scalar algebra (rotations of x, interpolation of
// <= 3 points, vanishing products).  All per-circuit constants (fixed / sigma columns in Lagrange,
// coefficient and extended-coset form, l_0 / l_last / l_active, the coset X column) stay resident in
// HBM inside the `sb_pk` handle and are reused by every proof.
#include <algorithm>
#include <chrono>
#include <set>

#include "handles.h"
#include "hostcrypto.h"
#include "json.h"
#include "prover.h"

namespace sb {
void host_fq_to_canonical(const uint8_t mont[32], uint8_t canon_le[32]);
}
using namespace sb;
using hfr::Fr;

namespace {

// ------------------------------------------------------------------ constraint system
struct JExpr;  // constraint-system expression before column indices are bound
struct ConstraintSystem {
    int A = 0, F = 0, I = 0, n_instances = 0;
    std::vector<std::pair<int, int>> advice_q, fixed_q;
    std::vector<json::ValueP> gates;
    struct Lookup { std::vector<json::ValueP> input, table; };
    std::vector<Lookup> lookups;
    std::vector<std::pair<std::string, int>> perm_cols;
    int degree = 0, blinding = 0;
    json::ValueP root;
};

Fr fr_from_hex(const std::string &hex) {
    // canonical value as a hex string ("0x..."), < r
    uint64_t l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t start = (hex.size() > 2 && hex[1] == 'x') ? 2 : 0;
    int nib = 0;
    for (size_t i = hex.size(); i-- > start; nib++) {
        char c = hex[i];
        uint64_t v = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : 0;
        if (nib < 128) l[nib >> 4] |= v << (4 * (nib & 15));
    }
    return hfr::from_u512(l);
}

ConstraintSystem parse_cs(const std::string &text) {
    ConstraintSystem cs;
    cs.root = json::parse(text);
    const json::Value &r = *cs.root;
    cs.A = (int)r.at("num_advice_columns").as_int();
    cs.F = (int)r.at("num_fixed_columns").as_int();
    cs.I = (int)r.at("num_instance_columns").as_int();
    cs.n_instances = r.has("num_instances") ? (int)r.at("num_instances").as_int() : 0;
    for (auto &q : r.at("advice_queries").arr) cs.advice_q.push_back({(int)(*q)[0].as_int(), (int)(*q)[1].as_int()});
    for (auto &q : r.at("fixed_queries").arr) cs.fixed_q.push_back({(int)(*q)[0].as_int(), (int)(*q)[1].as_int()});
    for (auto &g : r.at("gates").arr) cs.gates.push_back(g);
    for (auto &l : r.at("lookups").arr) {
        ConstraintSystem::Lookup lk;
        for (auto &e : l->at("input").arr) lk.input.push_back(e);
        for (auto &e : l->at("table").arr) lk.table.push_back(e);
        if (lk.input.size() != lk.table.size() || lk.input.empty()) throw std::runtime_error("lookup: input/table arity mismatch");
        cs.lookups.push_back(lk);
    }
    for (auto &c : r.at("permutation_columns").arr) cs.perm_cols.push_back({(*c)[0].as_str(), (int)(*c)[1].as_int()});
    cs.degree = (int)r.at("degree").as_int();
    cs.blinding = (int)r.at("blinding_factors").as_int();
    if (cs.I != 1) throw std::runtime_error("exactly one instance column is supported");
    return cs;
}

// column-index binding: (kind, column) -> index into a device pointer table
struct ColMap {
    int advice0 = 0, fixed0 = 0, instance0 = 0;
    int col(const std::string &kind, int c) const {
        if (kind == "advice") return advice0 + c;
        if (kind == "fixed") return fixed0 + c;
        if (kind == "instance") return instance0 + c;
        throw std::runtime_error("unknown column kind " + kind);
    }
};

ExprP bind_expr(const json::Value &e, const ColMap &m) {
    const std::string &k = e[0].as_str();
    if (k == "const") return e_const(to_dev(fr_from_hex(e[1].as_str())));
    if (k == "advice" || k == "fixed" || k == "instance") return e_col(m.col(k, (int)e[1].as_int()), (int)e[2].as_int());
    if (k == "neg") return e_neg(bind_expr(e[1], m));
    if (k == "add") return e_add(bind_expr(e[1], m), bind_expr(e[2], m));
    if (k == "sub") return e_sub(bind_expr(e[1], m), bind_expr(e[2], m));
    if (k == "mul") return e_mul(bind_expr(e[1], m), bind_expr(e[2], m));
    throw std::runtime_error("unknown expression node " + k);
}

ExprP ec(const Fr &x) { return e_const(to_dev(x)); }

// ------------------------------------------------------------------ transcripts
struct Transcript {
    std::vector<uint8_t> proof;
    virtual ~Transcript() {}
    virtual void common_scalar(const Fr &s) = 0;
    virtual void common_point(const uint8_t aff_mont[64]) = 0;
    virtual void emit_point(const uint8_t aff_mont[64]) = 0;
    virtual void emit_scalar(const Fr &s) = 0;
    virtual Fr squeeze() = 0;
    bool write_point(const uint8_t aff_mont[64]) {
        bool id = true;
        for (int i = 0; i < 64; i++) id = id && aff_mont[i] == 0;
        if (id) return false;  // both reference transcripts refuse the identity
        common_point(aff_mont);
        emit_point(aff_mont);
        return true;
    }
    void write_scalar(const Fr &s) {
        common_scalar(s);
        emit_scalar(s);
    }
};

struct KeccakTranscript : Transcript {
    std::vector<uint8_t> buf;
    static void be32(const uint8_t le[32], uint8_t out[32]) { for (int i = 0; i < 32; i++) out[i] = le[31 - i]; }
    void common_scalar(const Fr &s) override {
        uint8_t b[32];
        hfr::to_bytes_be(s, b);
        buf.insert(buf.end(), b, b + 32);
    }
    void point_bytes(const uint8_t aff[64], uint8_t out[64]) {
        uint8_t le[32];
        host_fq_to_canonical(aff, le);
        be32(le, out);
        host_fq_to_canonical(aff + 32, le);
        be32(le, out + 32);
    }
    void common_point(const uint8_t aff[64]) override {
        uint8_t b[64];
        point_bytes(aff, b);
        buf.insert(buf.end(), b, b + 64);
    }
    void emit_point(const uint8_t aff[64]) override {
        uint8_t b[64];
        point_bytes(aff, b);
        proof.insert(proof.end(), b, b + 64);
    }
    void emit_scalar(const Fr &s) override {
        uint8_t b[32];
        hfr::to_bytes_be(s, b);
        proof.insert(proof.end(), b, b + 32);
    }
    Fr squeeze() override {
        std::vector<uint8_t> data(buf);
        if (buf.size() == 32) data.push_back(1);
        uint8_t h[32], le[32];
        keccak256(data.data(), data.size(), h);
        buf.assign(h, h + 32);
        for (int i = 0; i < 32; i++) le[i] = h[31 - i];
        return hfr::from_bytes_le_wide(le, 32);
    }
};

struct Blake2bTranscript : Transcript {
    Blake2b st;
    Blake2bTranscript() { st.init(64, (const uint8_t *)"Halo2-Transcript"); }
    void common_scalar(const Fr &s) override {
        uint8_t b[33];
        b[0] = 2;
        hfr::to_bytes_le(s, b + 1);
        st.update(b, 33);
    }
    void common_point(const uint8_t aff[64]) override {
        uint8_t b[65];
        b[0] = 1;
        host_fq_to_canonical(aff, b + 1);
        host_fq_to_canonical(aff + 32, b + 33);
        st.update(b, 65);
    }
    void emit_point(const uint8_t aff[64]) override {
        uint8_t x[32], y[32];
        host_fq_to_canonical(aff, x);
        host_fq_to_canonical(aff + 32, y);
        x[31] |= (uint8_t)((y[0] & 1) << 6);
        proof.insert(proof.end(), x, x + 32);
    }
    void emit_scalar(const Fr &s) override {
        uint8_t b[32];
        hfr::to_bytes_le(s, b);
        proof.insert(proof.end(), b, b + 32);
    }
    Fr squeeze() override {
        uint8_t z = 0, h[64];
        st.update(&z, 1);
        st.final(h);
        return hfr::from_bytes_le_wide(h, 64);
    }
};

// host scalar helpers
Fr fpow(const Fr &b, uint64_t e) { return hfr::pow_u64(b, e); }
Fr rotate(const Fr &x, const Fr &omega, const Fr &omega_inv, int r) { return r >= 0 ? hfr::mul(x, fpow(omega, (uint64_t)r)) : hfr::mul(x, fpow(omega_inv, (uint64_t)(-r))); }

// Lagrange basis of a point set as coefficient vectors: basis[j](pts[k]) = [j == k].  One field inversion per point (host inversions are
// Fermat powers, ~15 us each), shared by every polynomial opened at this set.
std::vector<std::vector<Fr>> lagrange_basis(const std::vector<Fr> &pts) {
    const size_t n = pts.size();
    std::vector<std::vector<Fr>> basis(n);
    for (size_t j = 0; j < n; j++) {
        std::vector<Fr> num(1, hfr::ONE);
        Fr den = hfr::ONE;
        for (size_t k = 0; k < n; k++) {
            if (k == j) continue;
            std::vector<Fr> nxt(num.size() + 1, hfr::ZERO);
            for (size_t i = 0; i < num.size(); i++) {
                nxt[i + 1] = hfr::add(nxt[i + 1], num[i]);
                nxt[i] = hfr::sub(nxt[i], hfr::mul(pts[k], num[i]));
            }
            num.swap(nxt);
            den = hfr::mul(den, hfr::sub(pts[j], pts[k]));
        }
        if (n > 1) {
            const Fr dinv = hfr::inv(den);
            for (Fr &c : num) c = hfr::mul(c, dinv);
        }
        basis[j] = num;
    }
    return basis;
}
std::vector<Fr> lagrange_interpolate(const std::vector<std::vector<Fr>> &basis, const std::vector<Fr> &evals) {
    const size_t n = basis.size();
    std::vector<Fr> out(n, hfr::ZERO);
    for (size_t j = 0; j < n; j++)
        for (size_t i = 0; i < basis[j].size(); i++) out[i] = hfr::add(out[i], hfr::mul(basis[j][i], evals[j]));
    return out;
}
Fr eval_small(const std::vector<Fr> &c, const Fr &x) {
    Fr acc = hfr::ZERO;
    for (size_t i = c.size(); i-- > 0;) acc = hfr::add(hfr::mul(acc, x), c[i]);
    return acc;
}

struct FrLess {
    bool operator()(const Fr &a, const Fr &b) const { return hfr::cmp(a, b) < 0; }
};

}  // namespace

// ------------------------------------------------------------------ proving key handle
struct sb_pk {
    ConstraintSystem cs;
    const sb_srs *srs = nullptr;
    sb_domain *dom = nullptr;
    uint32_t k = 0, ext_k = 0;
    size_t n = 0, ext_n = 0;
    Fr transcript_repr;
    int P = 0;
    // device-resident (all Montgomery Fr arrays)
    // The quotient h(X) has degree < (j - 1) n, so its numerator is evaluated on n_cos = j - 1 cosets g_s H of the size-n subgroup (g_s = zeta *
    // ext_omega^s, s < n_cos: the first n_cos of the 2^(ext_k - k) cosets that make up halo2's extended domain) instead of on all of them: every
    // "coset" array below is COSET-MAJOR, [n_cos][n], slot s holding the values on g_s H.  h's coefficients come back by one size-n inverse NTT per
    // coset and a constant n_cos x n_cos matrix (the inverse Vandermonde of g_s^n, with 1 / t(g_s) folded in): the same polynomial, 5/8 of the work.
    int n_cos = 0;
    fr_t combine[64];                                            // row-major [q][s]: h_{q n + i} = sum_s combine[q][s] * d_s[i]
    std::vector<void *> fixed_values, fixed_polys, fixed_cosets;
    std::vector<void *> sigma_values, sigma_polys, sigma_cosets;
    void *l0 = nullptr, *l_last = nullptr, *l_active = nullptr;  // coset-major
    void *x_coset = nullptr;                                     // X on the cosets: g_s * omega^i (coset-major)
    void *omega_pows = nullptr;                                  // omega^i (n)
    void *div_g_pows = nullptr, *div_x = nullptr, *div_ginv_scaled = nullptr;  // SHPLONK coset division: g^i, g*omega^i, g^-i / n
    std::vector<uint8_t> fixed_comms, sigma_comms;               // affine, 64 B each
    // evaluate_h program compiled once per key with placeholder challenges; per proof only the challenge-derived constants are patched
    struct HProgramCache {
        std::mutex mu;
        bool valid = false;
        Program prog;
        std::vector<uint32_t> theta, beta, gamma;               // constant slots holding theta / beta / gamma
        std::vector<std::pair<uint32_t, int>> delta_beta, ypow;  // (slot, j): delta^j * beta ; (slot, e): y^e
        // the program as straight-line sm_100a code (NVRTC, compiled on first use); nullptr + jit_tried: not available, the interpreter runs
        ExprJit *jit = nullptr;
        bool jit_tried = false;
        std::string jit_why;
    };
    mutable HProgramCache hcache;
    // g_s^m and g_s^-m (m < n) for coset slot s, coset-major slabs [n_cos][n]: coefficient scaling before the forward / after the inverse size-n NTT
    void *coset_pows = nullptr, *coset_pows_inv = nullptr;
    mutable std::vector<void *> owned;
};

namespace {

int32_t dalloc(const sb_pk *pk, size_t bytes, void **out) {
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 32);
    if (e != cudaSuccess) {
        set_last_error("pk: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return SB_ERR_ALLOC;
    }
    pk->owned.push_back(*out);
    return SB_OK;
}

const Fr DIV_G = hfr::from_u64(7);  // multiplicative generator: g * H is disjoint from H

struct SparseAssignment {  // keygen output in sparse form (cells that differ from the default column)
    const uint32_t *fixed_cells = nullptr;  // (col, row) pairs
    const uint8_t *fixed_values = nullptr;  // 32 B each
    size_t n_fixed = 0;
    const uint32_t *perm_cells = nullptr;   // (col, row, to_col, to_row): sigma_col(omega^row) = delta^to_col * omega^to_row
    size_t n_perm = 0;
};

int32_t pk_build(sb_ctx *ctx, sb_pk *pk, const uint8_t *fixed_values, const uint8_t *sigma_values, const SparseAssignment *sparse, cudaStream_t st) {
    const size_t n = pk->n;
    const sb_domain *d = pk->dom;
    const Fr DELTA = fr_from_hex("0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2");
    // ---- the cosets of the quotient argument and the matrix that turns per-coset data back into h's coefficients
    const int m = pk->n_cos = (int)d->quotient_degree;
    const size_t cn = (size_t)m * n;  // elements of a coset-major array
    SB_REQUIRE(m >= 1 && m <= 8 && (uint32_t)m <= d->n_t, "quotient degree must be within the cosets of the extended domain");
    {
        const Fr zeta = to_host(d->coset[1]), ext_omega = to_host(d->ext_omega);
        std::vector<Fr> g(m), v(m);
        for (int s2 = 0; s2 < m; s2++) {
            g[s2] = hfr::mul(zeta, hfr::pow_u64(ext_omega, (uint64_t)s2));
            v[s2] = hfr::pow_u64(g[s2], (uint64_t)n);  // g_s^n: the node of the Vandermonde system sum_q h_q(i) v_s^q = d_s(i)
        }
        // invert V[s][q] = v_s^q by Gauss-Jordan (m <= 8)
        std::vector<std::vector<Fr>> a(m, std::vector<Fr>(2 * m, hfr::ZERO));
        for (int s2 = 0; s2 < m; s2++) {
            Fr pw = hfr::ONE;
            for (int q = 0; q < m; q++) { a[s2][q] = pw; pw = hfr::mul(pw, v[s2]); }
            a[s2][m + s2] = hfr::ONE;
        }
        for (int c = 0; c < m; c++) {
            int piv = c;
            while (piv < m && a[piv][c] == hfr::ZERO) piv++;
            SB_REQUIRE(piv < m, "quotient cosets: singular Vandermonde system");
            std::swap(a[c], a[piv]);
            const Fr iv = hfr::inv(a[c][c]);
            for (int x = 0; x < 2 * m; x++) a[c][x] = hfr::mul(a[c][x], iv);
            for (int r2 = 0; r2 < m; r2++) {
                if (r2 == c || a[r2][c] == hfr::ZERO) continue;
                const Fr f = a[r2][c];
                for (int x = 0; x < 2 * m; x++) a[r2][x] = hfr::sub(a[r2][x], hfr::mul(f, a[c][x]));
            }
        }
        // combine[q][s] = Vinv[q][s] / t(g_s): the division by the vanishing polynomial (constant on a coset) rides along
        for (int q = 0; q < m; q++)
            for (int s2 = 0; s2 < m; s2++) pk->combine[q * 8 + s2] = to_dev(hfr::mul(a[q][m + s2], to_host(d->t_inv[s2])));
        SB_TRY(dalloc(pk, cn * 32, &pk->coset_pows));
        SB_TRY(dalloc(pk, cn * 32, &pk->coset_pows_inv));
        for (int s2 = 0; s2 < m; s2++) {
            SB_TRY(fr_gen_powers(ctx, (uint8_t *)pk->coset_pows + (size_t)s2 * n * 32, to_dev(g[s2]), n, st));
            SB_TRY(fr_gen_powers(ctx, (uint8_t *)pk->coset_pows_inv + (size_t)s2 * n * 32, to_dev(hfr::inv(g[s2])), n, st));
        }
    }
    // values of a coefficient-form polynomial on every coset, coset-major: ONE batched launch set (the same input, n_cos scalings, n_cos outputs)
    auto to_cosets = [&](const void *d_coeff, void *d_cm) -> int32_t {
        NttFuse f;
        f.pre_vec = pk->coset_pows;
        f.batch = (uint32_t)m;
        f.src_stride = 0; f.pre_stride = n; f.dst_stride = n;
        return ntt_run_fused(ctx, d_coeff, d_cm, (const uint8_t *)d->omega.v, pk->k, &f, st);
    };
    // omega^i first: the sparse permutation columns are built from it
    SB_TRY(dalloc(pk, n * 32, &pk->omega_pows));
    SB_TRY(fr_gen_powers(ctx, pk->omega_pows, d->omega, n, st));
    auto alloc_forms = [&](int count, std::vector<void *> &vals, std::vector<void *> &polys, std::vector<void *> &cosets) -> int32_t {
        for (int c = 0; c < count; c++) {
            void *v, *p, *e;
            SB_TRY(dalloc(pk, n * 32, &v));
            SB_TRY(dalloc(pk, n * 32, &p));
            SB_TRY(dalloc(pk, cn * 32, &e));
            vals.push_back(v); polys.push_back(p); cosets.push_back(e);
        }
        return SB_OK;
    };
    SB_TRY(alloc_forms(pk->cs.F, pk->fixed_values, pk->fixed_polys, pk->fixed_cosets));
    SB_TRY(alloc_forms(pk->P, pk->sigma_values, pk->sigma_polys, pk->sigma_cosets));
    if (!sparse) {
        for (int c = 0; c < pk->cs.F; c++) SB_CUDA_TRY(cudaMemcpyAsync(pk->fixed_values[c], fixed_values + (size_t)c * n * 32, n * 32, cudaMemcpyHostToDevice, st));
        for (int c = 0; c < pk->P; c++) SB_CUDA_TRY(cudaMemcpyAsync(pk->sigma_values[c], sigma_values + (size_t)c * n * 32, n * 32, cudaMemcpyHostToDevice, st));
    } else {
        // keygen's sparse output: ONE upload of the cell list and ONE scatter kernel per kind (the k = 20..23 keys have 10^5..10^6 assigned cells)
        SB_REQUIRE(pk->cs.F <= 32 && pk->P <= 16, "sparse key: at most 32 fixed and 16 permutation columns");
        for (size_t i = 0; i < sparse->n_fixed; i++)
            SB_REQUIRE((int)sparse->fixed_cells[2 * i] < pk->cs.F && sparse->fixed_cells[2 * i + 1] < n, "sparse fixed cell out of range");
        for (size_t i = 0; i < sparse->n_perm; i++)
            SB_REQUIRE((int)sparse->perm_cells[4 * i] < pk->P && sparse->perm_cells[4 * i + 1] < n && (int)sparse->perm_cells[4 * i + 2] < pk->P && sparse->perm_cells[4 * i + 3] < n,
                       "sparse permutation cell out of range");
        void *d_cells, *d_vals;
        SB_TRY(scratch_get(ctx, "pk_cells", std::max(sparse->n_fixed * 8, sparse->n_perm * 16) + 16, &d_cells));
        SB_TRY(scratch_get(ctx, "pk_cell_vals", sparse->n_fixed * 32 + 32, &d_vals));
        for (int c = 0; c < pk->cs.F; c++) SB_CUDA_TRY(cudaMemsetAsync(pk->fixed_values[c], 0, n * 32, st));
        if (sparse->n_fixed) {
            SB_CUDA_TRY(cudaMemcpyAsync(d_cells, sparse->fixed_cells, sparse->n_fixed * 8, cudaMemcpyHostToDevice, st));
            SB_CUDA_TRY(cudaMemcpyAsync(d_vals, sparse->fixed_values, sparse->n_fixed * 32, cudaMemcpyHostToDevice, st));
            SB_TRY(scatter_cells(ctx, pk->fixed_values.data(), (uint32_t)pk->cs.F, d_cells, d_vals, sparse->n_fixed, st));
        }
        // identity permutation delta^c * omega^row, then the cells moved by copy constraints: delta^to_col * omega^to_row
        std::vector<fr_t> delta_pows(pk->P);
        Fr dpow = hfr::ONE;
        for (int c = 0; c < pk->P; c++) {
            delta_pows[c] = to_dev(dpow);
            SB_CUDA_TRY(cudaMemcpyAsync(pk->sigma_values[c], pk->omega_pows, n * 32, cudaMemcpyDeviceToDevice, st));
            if (c) SB_TRY(fr_scale(ctx, pk->sigma_values[c], n, delta_pows[c], st));
            dpow = hfr::mul(dpow, DELTA);
        }
        if (sparse->n_perm) {
            SB_CUDA_TRY(cudaMemcpyAsync(d_cells, sparse->perm_cells, sparse->n_perm * 16, cudaMemcpyHostToDevice, st));
            SB_TRY(sigma_patch(ctx, pk->sigma_values.data(), (uint32_t)pk->P, d_cells, sparse->n_perm, pk->omega_pows, delta_pows.data(), st));
        }
        SB_CUDA_TRY(sync_stream(ctx, st));  // the caller's (pageable) cell arrays are free again
    }
    auto derive_forms = [&](int count, std::vector<void *> &vals, std::vector<void *> &polys, std::vector<void *> &cosets, std::vector<uint8_t> &comms) -> int32_t {
        comms.resize((size_t)count * 64);
        for (int c = 0; c < count; c++) {
            SB_CUDA_TRY(cudaMemcpyAsync(polys[c], vals[c], n * 32, cudaMemcpyDeviceToDevice, st));
            SB_TRY(dom_l2c(ctx, d, polys[c], st));
            SB_TRY(to_cosets(polys[c], cosets[c]));
            SB_TRY(srs_msm(ctx, pk->srs, 1, vals[c], n, comms.data() + (size_t)c * 64, st));
        }
        return SB_OK;
    };
    SB_TRY(derive_forms(pk->cs.F, pk->fixed_values, pk->fixed_polys, pk->fixed_cosets, pk->fixed_comms));
    SB_TRY(derive_forms(pk->P, pk->sigma_values, pk->sigma_polys, pk->sigma_cosets, pk->sigma_comms));
    // l_0, l_last, l_blind (Lagrange unit vectors) -> cosets; l_active = 1 - (l_last + l_blind)
    const int bf = pk->cs.blinding;
    std::vector<fr_t> tmp(n, fr_t::zero());
    void *d_tmp, *d_lblind;
    SB_TRY(scratch_get(ctx, "pk_tmp", n * 32, &d_tmp));
    SB_TRY(scratch_get(ctx, "pk_lblind", cn * 32, &d_lblind));
    auto unit_cosets = [&](const std::vector<size_t> &rows, void *d_cm) -> int32_t {
        std::fill(tmp.begin(), tmp.end(), fr_t::zero());
        for (size_t r : rows) tmp[r] = fr_t::one();
        SB_CUDA_TRY(cudaMemcpyAsync(d_tmp, tmp.data(), n * 32, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(sync_stream(ctx, st));
        SB_TRY(dom_l2c(ctx, d, d_tmp, st));
        return to_cosets(d_tmp, d_cm);
    };
    SB_TRY(dalloc(pk, cn * 32, &pk->l0));
    SB_TRY(dalloc(pk, cn * 32, &pk->l_last));
    SB_TRY(dalloc(pk, cn * 32, &pk->l_active));
    SB_TRY(unit_cosets({0}, pk->l0));
    SB_TRY(unit_cosets({n - (size_t)bf - 1}, pk->l_last));
    std::vector<size_t> blind_rows;
    for (size_t r = n - (size_t)bf; r < n; r++) blind_rows.push_back(r);
    SB_TRY(unit_cosets(blind_rows, d_lblind));
    {
        Program p = compile_terms({e_sub(e_const(fr_t::one()), e_add(e_col(0, 0), e_col(1, 0)))}, nullptr);
        for (int s2 = 0; s2 < m; s2++) {
            const size_t off = (size_t)s2 * n * 32;
            std::vector<const void *> cols = {(const uint8_t *)pk->l_last + off, (const uint8_t *)d_lblind + off};
            SB_TRY(expr_eval(ctx, p, cols, pk->k, 0, (uint8_t *)pk->l_active + off, st));
        }
    }
    // X on coset s: g_s * omega^i ; SHPLONK division helpers
    SB_TRY(dalloc(pk, cn * 32, &pk->x_coset));
    for (int s2 = 0; s2 < m; s2++) {
        void *xs = (uint8_t *)pk->x_coset + (size_t)s2 * n * 32;
        SB_CUDA_TRY(cudaMemcpyAsync(xs, pk->omega_pows, n * 32, cudaMemcpyDeviceToDevice, st));
        SB_TRY(fr_scale(ctx, xs, n, to_dev(hfr::mul(to_host(d->coset[1]), hfr::pow_u64(to_host(d->ext_omega), (uint64_t)s2))), st));
    }
    SB_TRY(dalloc(pk, n * 32, &pk->div_g_pows));
    SB_TRY(dalloc(pk, n * 32, &pk->div_x));
    SB_TRY(dalloc(pk, n * 32, &pk->div_ginv_scaled));
    SB_TRY(fr_gen_powers(ctx, pk->div_g_pows, to_dev(DIV_G), n, st));
    SB_CUDA_TRY(cudaMemcpyAsync(pk->div_x, pk->omega_pows, n * 32, cudaMemcpyDeviceToDevice, st));
    SB_TRY(fr_scale(ctx, pk->div_x, n, to_dev(DIV_G), st));
    SB_TRY(fr_gen_powers(ctx, pk->div_ginv_scaled, to_dev(hfr::inv(DIV_G)), n, st));
    SB_TRY(fr_scale(ctx, pk->div_ginv_scaled, n, d->ifft_divisor, st));
    SB_CUDA_TRY(sync_stream(ctx, st));
    return SB_OK;
}

// q(X) = p(X) / prod (X - root): exact division through the coset g*H (p is overwritten by q, n coefficients)
int32_t poly_div_by_roots_k(sb_ctx *ctx, uint32_t k, const fr_t &omega_d, const fr_t &omega_inv_d, const void *g_pows, const void *div_x, const void *ginv_scaled,
                            void *d_p, const std::vector<Fr> &roots, cudaStream_t st) {
    const size_t n = (size_t)1 << k;
    void *d_den;
    SB_TRY(scratch_get(ctx, "div_den", n * 32, &d_den));
    {
        NttFuse f;
        f.pre_vec = g_pows;
        SB_TRY(ntt_run_fused(ctx, d_p, d_p, (const uint8_t *)omega_d.v, k, &f, st));
    }
    std::vector<fr_t> rd;
    for (const Fr &r : roots) rd.push_back(to_dev(r));
    SB_TRY(fr_vanish(ctx, div_x, rd, d_den, n, st));
    SB_TRY(fr_batch_invert(ctx, d_den, n, st));
    SB_TRY(fp_vec_op(ctx, 0, 0, d_p, d_den, d_p, n, st));
    NttFuse fi;
    fi.post_vec = ginv_scaled;
    return ntt_run_fused(ctx, d_p, d_p, (const uint8_t *)omega_inv_d.v, k, &fi, st);
}
int32_t poly_div_by_roots(sb_ctx *ctx, const sb_pk *pk, void *d_p, const std::vector<Fr> &roots, cudaStream_t st) {
    return poly_div_by_roots_k(ctx, pk->k, pk->dom->omega, pk->dom->omega_inv, pk->div_g_pows, pk->div_x, pk->div_ginv_scaled, d_p, roots, st);
}

int32_t upload_frs(sb_ctx *ctx, void *d_dst, const std::vector<Fr> &v, cudaStream_t st) {
    if (v.empty()) return SB_OK;
    return h2d_staged(ctx, d_dst, v.data(), v.size() * 32, st);
}

int32_t msm_commit(sb_ctx *ctx, const sb_comm *comm, const sb_srs *srs, int basis, const void *d_scalars, size_t n, uint8_t out[64], cudaStream_t st);

struct Query {
    int poly_id;
    Fr point;
    const void *d_poly;
    Fr eval;
    const void *d_lagr = nullptr;  // the same polynomial's values on the size-n subgroup, when the prover still holds them
};

int32_t shplonk(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, Transcript &tr, const std::vector<Query> &queries, cudaStream_t st) {
    const size_t n = pk->n;
    const Fr y = tr.squeeze();
    // construct_intermediate_sets: first-appearance order of polynomials and of rotation sets,
    // points inside a set in ascending field order
    std::vector<int> poly_order;
    std::map<int, std::set<Fr, FrLess>> poly_points;
    std::map<int, const void *> polys;
    std::set<Fr, FrLess> super_points;
    auto find_eval = [&](int pid, const Fr &pt) -> Fr {
        for (const Query &q : queries)
            if (q.poly_id == pid && q.point == pt) return q.eval;
        return hfr::ZERO;
    };
    for (const Query &q : queries) {
        super_points.insert(q.point);
        if (!poly_points.count(q.poly_id)) {
            poly_order.push_back(q.poly_id);
            polys[q.poly_id] = q.d_poly;
        }
        poly_points[q.poly_id].insert(q.point);
    }
    struct RSet { std::vector<Fr> pts; std::vector<int> members; };
    std::vector<RSet> sets;
    for (int pid : poly_order) {
        std::vector<Fr> key(poly_points[pid].begin(), poly_points[pid].end());
        bool found = false;
        for (RSet &s : sets)
            if (s.pts.size() == key.size() && std::equal(key.begin(), key.end(), s.pts.begin())) { s.members.push_back(pid); found = true; break; }
        if (!found) sets.push_back({key, {pid}});
    }
    const Fr v = tr.squeeze();
    void *d_hx, *d_nx, *d_lx, *d_invd;
    SB_TRY(scratch_get(ctx, "sh_hx", n * 32, &d_hx));
    SB_TRY(scratch_get(ctx, "sh_nx", n * 32, &d_nx));
    SB_TRY(scratch_get(ctx, "sh_lx", n * 32, &d_lx));
    SB_TRY(scratch_get(ctx, "sh_invd", n * 32, &d_invd));
    // The whole argument in the evaluation domain H itself (default).  Almost every opened polynomial still exists as its values on H (advice,
    // fixed and sigma columns, the grand products, the permuted lookup columns), so N_i on H is a linear combination of those vectors -- no
    // transform -- plus ONE transform for the members held as coefficients only (the folded quotient and the random polynomial, both in the set
    // {x}); h and the opening quotient stay in the evaluation domain and are committed over the Lagrange basis (deg < n: the same points).
    // 1 transform instead of 8.  Needs every opening point outside H (a point of H would be a root of Z_T on H): else the coset path below.
    bool on_h = !ctx->tune.no_shplonk_lagrange && pk->srs->d_g_lagrange != nullptr;
    for (const Fr &p : super_points) {
        Fr t = p;
        for (uint32_t i = 0; i < pk->k; i++) t = hfr::mul(t, t);
        if (t == hfr::ONE) on_h = false;
    }
    if (on_h) {
        std::map<int, const void *> lagr;
        for (const Query &q : queries)
            if (q.d_lagr) lagr[q.poly_id] = q.d_lagr;
        const std::vector<Fr> super(super_points.begin(), super_points.end());
        // Sharded proof: everything below except the one transform is elementwise over H, so rank r computes rows [r n / world, (r + 1) n / world)
        // of h and of the opening quotient; the slices meet in one all-gather each, right before their (window-sharded) commitments.
        const bool by_rows = comm && comm->world > 1 && comm->allgather_dev && n % ((size_t)comm->world * 32) == 0 && !ctx->tune.no_shplonk_shard;
        const size_t cnt = by_rows ? n / (size_t)comm->world : n, off = by_rows ? cnt * (size_t)comm->rank : 0;
        auto sl = [&](const void *p) -> const void * { return (const uint8_t *)p + off * 32; };
        auto slw = [&](void *p) -> void * { return (uint8_t *)p + off * 32; };
        auto gather = [&](void *d_buf) -> int32_t {
            if (!by_rows) return SB_OK;
            SB_CUDA_TRY(sync_stream(ctx, st));
            if (comm->allgather_dev(comm->user, d_buf, cnt * 32, (void *)st) != 0) { set_last_error("sb_comm.allgather_dev failed"); return SB_ERR_ARG; }
            return SB_OK;
        };
        uint8_t *d_m;
        SB_TRY(scratch_get(ctx, "sh_m", sets.size() * n * 32, (void **)&d_m));
        {
            std::vector<fr_t> sd;
            for (const Fr &p : super) sd.push_back(to_dev(p));
            SB_TRY(fr_vanish(ctx, sl(pk->omega_pows), sd, slw(d_invd), cnt, st));
            SB_TRY(fr_batch_invert(ctx, slw(d_invd), cnt, st));
        }
        std::vector<std::vector<std::vector<Fr>>> low(sets.size());
        auto complement = [&](const RSet &s) {
            std::vector<Fr> c;
            for (const Fr &p : super) {
                bool in = false;
                for (const Fr &q : s.pts) in = in || (q == p);
                if (!in) c.push_back(p);
            }
            return c;
        };
        Fr v_pow = hfr::ONE;
        for (size_t si = 0; si < sets.size(); si++) {
            RSet &s = sets[si];
            Fr y_pow = hfr::ONE;
            std::vector<Fr> head(s.pts.size(), hfr::ZERO);
            std::vector<const void *> lag_polys, coef_polys;
            std::vector<fr_t> lag_coeffs, coef_coeffs;
            const std::vector<std::vector<Fr>> basis = lagrange_basis(s.pts);
            for (size_t mi = 0; mi < s.members.size(); mi++) {
                const int pid = s.members[mi];
                std::vector<Fr> evs;
                for (const Fr &p : s.pts) evs.push_back(find_eval(pid, p));
                std::vector<Fr> r = lagrange_interpolate(basis, evs);
                low[si].push_back(r);
                for (size_t i = 0; i < r.size(); i++) head[i] = hfr::add(head[i], hfr::mul(r[i], y_pow));
                if (lagr.count(pid)) { lag_polys.push_back(sl(lagr[pid])); lag_coeffs.push_back(to_dev(y_pow)); }
                else { coef_polys.push_back(polys[pid]); coef_coeffs.push_back(to_dev(y_pow)); }
                y_pow = hfr::mul(y_pow, y);
            }
            std::vector<fr_t> head_d;
            for (const Fr &h : head) head_d.push_back(to_dev(h));
            SB_REQUIRE(head_d.size() <= 8, "shplonk: rotation sets of more than 8 points are not supported");
            void *d_mi = d_m + si * n * 32;   // M_i = sum_m y^m p_m on H (kept for the opening quotient)
            if (!coef_polys.empty()) {
                SB_TRY(fr_lincomb(ctx, d_nx, coef_polys, coef_coeffs, {}, n, false, st));
                SB_TRY(ntt_run(ctx, d_nx, (const uint8_t *)pk->dom->omega.v, pk->k, st));
                lag_polys.push_back(sl(d_nx));
                lag_coeffs.push_back(to_dev(hfr::ONE));
            }
            SB_TRY(fr_lincomb(ctx, slw(d_mi), lag_polys, lag_coeffs, {}, cnt, false, st));
            std::vector<fr_t> comp;
            for (const Fr &p : complement(s)) comp.push_back(to_dev(p));
            SB_TRY(fr_div_combine(ctx, slw(d_hx), sl(d_mi), sl(d_invd), sl(pk->omega_pows), comp, to_dev(v_pow), cnt, si == 0, st, &head_d));
            v_pow = hfr::mul(v_pow, v);
        }
        uint8_t pt[64];
        SB_TRY(gather(d_hx));
        SB_TRY(msm_commit(ctx, comm, pk->srs, 1, d_hx, n, pt, st));
        if (!tr.write_point(pt)) { set_last_error("shplonk: quotient commitment is the identity"); return SB_ERR_ARG; }
        const Fr u = tr.squeeze();
        {   // u in H would make X - u vanish on the evaluation domain: negligible, but not silently wrong
            Fr t = u;
            for (uint32_t i = 0; i < pk->k; i++) t = hfr::mul(t, t);
            SB_REQUIRE(!(t == hfr::ONE), "shplonk: the opening challenge lies in the evaluation domain");
        }
        std::vector<const void *> lc_polys;
        std::vector<fr_t> lc_coeffs;
        Fr const_term = hfr::ZERO, z0 = hfr::ONE;
        v_pow = hfr::ONE;
        for (size_t si = 0; si < sets.size(); si++) {
            Fr z_i = hfr::ONE;
            for (const Fr &p : complement(sets[si])) z_i = hfr::mul(z_i, hfr::sub(u, p));
            if (si == 0) z0 = z_i;
            const Fr scale = hfr::mul(z_i, v_pow);
            Fr y_pow = hfr::ONE;
            for (size_t mi = 0; mi < sets[si].members.size(); mi++) {
                const_term = hfr::add(const_term, hfr::mul(hfr::mul(scale, y_pow), eval_small(low[si][mi], u)));
                y_pow = hfr::mul(y_pow, y);
            }
            lc_polys.push_back(sl(d_m + si * n * 32));
            lc_coeffs.push_back(to_dev(scale));
            v_pow = hfr::mul(v_pow, v);
        }
        Fr zt = hfr::ONE;
        for (const Fr &p : super) zt = hfr::mul(zt, hfr::sub(u, p));
        lc_polys.push_back(sl(d_hx));
        lc_coeffs.push_back(to_dev(hfr::neg(zt)));
        SB_TRY(fr_lincomb(ctx, slw(d_lx), lc_polys, lc_coeffs, {}, cnt, false, st));
        SB_TRY(fr_vanish(ctx, sl(pk->omega_pows), {to_dev(u)}, slw(d_nx), cnt, st));
        SB_TRY(fr_batch_invert(ctx, slw(d_nx), cnt, st));
        SB_TRY(fr_open_quotient(ctx, slw(d_lx), sl(d_lx), sl(d_nx), to_dev(const_term), to_dev(hfr::inv(z0)), cnt, st));
        SB_TRY(gather(d_lx));
        SB_TRY(msm_commit(ctx, comm, pk->srs, 1, d_lx, n, pt, st));
        if (!tr.write_point(pt)) { set_last_error("shplonk: opening commitment is the identity"); return SB_ERR_ARG; }
        return SB_OK;
    }
    // h(X) = sum_i v^i N_i(X) / Z_{S_i}(X), all divisions exact.  Work on the coset g*H: E = sum_i v^i NTT(g^j N_i) * (1 / Z_{S_i}) is
    // accumulated in the evaluation domain, so the whole sum needs ONE inverse NTT, and 1 / Z_{S_i} = (1 / Z_T) * prod_{p in T \ S_i}(x - p)
    // with T the super set of opening points, so ONE batch inversion serves every rotation set.
    const std::vector<Fr> super(super_points.begin(), super_points.end());
    {
        std::vector<fr_t> sd;
        for (const Fr &p : super) sd.push_back(to_dev(p));
        SB_TRY(fr_vanish(ctx, pk->div_x, sd, d_invd, n, st));
        SB_TRY(fr_batch_invert(ctx, d_invd, n, st));
    }
    std::vector<std::vector<std::vector<Fr>>> low(sets.size());
    Fr v_pow = hfr::ONE;
    for (size_t si = 0; si < sets.size(); si++) {
        RSet &s = sets[si];
        Fr y_pow = hfr::ONE;
        std::vector<Fr> head(s.pts.size(), hfr::ZERO);
        std::vector<const void *> lc_polys;
        std::vector<fr_t> lc_coeffs;
        const std::vector<std::vector<Fr>> basis = lagrange_basis(s.pts);
        for (size_t mi = 0; mi < s.members.size(); mi++) {
            const int pid = s.members[mi];
            std::vector<Fr> evs;
            for (const Fr &p : s.pts) evs.push_back(find_eval(pid, p));
            std::vector<Fr> r = lagrange_interpolate(basis, evs);
            low[si].push_back(r);
            for (size_t i = 0; i < r.size(); i++) head[i] = hfr::add(head[i], hfr::mul(r[i], y_pow));
            lc_polys.push_back(polys[pid]);
            lc_coeffs.push_back(to_dev(y_pow));
            y_pow = hfr::mul(y_pow, y);
        }
        std::vector<fr_t> head_d;
        for (const Fr &h : head) head_d.push_back(to_dev(h));
        SB_REQUIRE(head_d.size() <= 4, "shplonk: rotation sets of more than 4 points are not supported");
        SB_TRY(fr_lincomb(ctx, d_nx, lc_polys, lc_coeffs, head_d, n, false, st));
        // N_i on the coset
        {
            NttFuse f;
            f.pre_vec = pk->div_g_pows;
            SB_TRY(ntt_run_fused(ctx, d_nx, d_nx, (const uint8_t *)pk->dom->omega.v, pk->k, &f, st));
        }
        std::vector<fr_t> comp;
        for (const Fr &p : super) {
            bool in = false;
            for (const Fr &q : s.pts) in = in || (q == p);
            if (!in) comp.push_back(to_dev(p));
        }
        SB_TRY(fr_div_combine(ctx, d_hx, d_nx, d_invd, pk->div_x, comp, to_dev(v_pow), n, si == 0, st));
        v_pow = hfr::mul(v_pow, v);
    }
    {
        NttFuse f;
        f.post_vec = pk->div_ginv_scaled;  // g^-i / n
        SB_TRY(ntt_run_fused(ctx, d_hx, d_hx, (const uint8_t *)pk->dom->omega_inv.v, pk->k, &f, st));
    }
    uint8_t pt[64];
    SB_TRY(msm_commit(ctx, comm, pk->srs, 0, d_hx, n, pt, st));
    if (!tr.write_point(pt)) { set_last_error("shplonk: quotient commitment is the identity"); return SB_ERR_ARG; }
    const Fr u = tr.squeeze();
    std::vector<Fr> z_diffs;
    v_pow = hfr::ONE;
    Fr const_term = hfr::ZERO;
    std::vector<const void *> lc_polys;
    std::vector<fr_t> lc_coeffs;
    for (size_t si = 0; si < sets.size(); si++) {
        RSet &s = sets[si];
        Fr z_i = hfr::ONE;
        for (const Fr &p : super) {
            bool in = false;
            for (const Fr &q : s.pts) in = in || (q == p);
            if (!in) z_i = hfr::mul(z_i, hfr::sub(u, p));
        }
        z_diffs.push_back(z_i);
        Fr y_pow = hfr::ONE;
        const Fr scale = hfr::mul(z_i, v_pow);
        for (size_t mi = 0; mi < s.members.size(); mi++) {
            const Fr coeff = hfr::mul(scale, y_pow);
            lc_polys.push_back(polys[s.members[mi]]);
            lc_coeffs.push_back(to_dev(coeff));
            const_term = hfr::add(const_term, hfr::mul(coeff, eval_small(low[si][mi], u)));
            y_pow = hfr::mul(y_pow, y);
        }
        v_pow = hfr::mul(v_pow, v);
    }
    Fr zt = hfr::ONE;
    for (const Fr &p : super) zt = hfr::mul(zt, hfr::sub(u, p));
    lc_polys.push_back(d_hx);
    lc_coeffs.push_back(to_dev(hfr::neg(zt)));
    SB_TRY(fr_lincomb(ctx, d_lx, lc_polys, lc_coeffs, {to_dev(const_term)}, n, false, st));
    SB_TRY(poly_div_by_roots(ctx, pk, d_lx, {u}, st));
    SB_TRY(fr_scale(ctx, d_lx, n, to_dev(hfr::inv(z_diffs[0])), st));
    SB_TRY(msm_commit(ctx, comm, pk->srs, 0, d_lx, n, pt, st));
    if (!tr.write_point(pt)) { set_last_error("shplonk: opening commitment is the identity"); return SB_ERR_ARG; }
    return SB_OK;
}


// ------------------------------------------------------------------ quotient numerator terms
// Column-table layout of evaluate_h (both the full extended domain and one coset of it):
//   advice | fixed | instance | sigma (P) | permutation Z (n_sets) | l_0, l_last, l_active, X | per lookup: Z, A', S'
struct HLayout {
    int A, F, E_SIGMA, E_PZ, E_L0, E_LLAST, E_LACT, E_X, E_LK, n_cols;
    HLayout(const ConstraintSystem &cs, int P, int n_sets, size_t n_lookups) {
        A = cs.A; F = cs.F;
        E_SIGMA = A + F + 1; E_PZ = E_SIGMA + P; E_L0 = E_PZ + n_sets; E_LLAST = E_L0 + 1; E_LACT = E_L0 + 2; E_X = E_L0 + 3; E_LK = E_L0 + 4;
        n_cols = E_LK + 3 * (int)n_lookups;
    }
};

// gate polynomials, then the permutation argument's terms, then every lookup's terms, in halo2's fold order
std::vector<ExprP> h_terms(const ConstraintSystem &cs, int P, const std::vector<std::pair<int, int>> &sets /* (first column, count) */, size_t n_lookups,
                           const Fr &theta, const Fr &beta, const Fr &gamma) {
    const int n_sets = (int)sets.size(), bf = cs.blinding;
    const HLayout lay(cs, P, n_sets, n_lookups);
    const int E_SIGMA = lay.E_SIGMA, E_PZ = lay.E_PZ, E_L0 = lay.E_L0, E_LLAST = lay.E_LLAST, E_LACT = lay.E_LACT, E_X = lay.E_X, E_LK = lay.E_LK;
    ColMap em;
    em.advice0 = 0; em.fixed0 = cs.A; em.instance0 = cs.A + cs.F;
    std::vector<ExprP> terms;
    for (auto &g : cs.gates) terms.push_back(bind_expr(*g, em));
    const ExprP one = e_const(fr_t::one());
    const ExprP l0 = e_col(E_L0, 0), llast = e_col(E_LLAST, 0), lact = e_col(E_LACT, 0);
    if (n_sets > 0) {
        auto Z = [&](int s, int r) { return e_col(E_PZ + s, r); };
        terms.push_back(e_mul(e_sub(one, Z(0, 0)), l0));
        terms.push_back(e_mul(e_sub(e_mul(Z(n_sets - 1, 0), Z(n_sets - 1, 0)), Z(n_sets - 1, 0)), llast));
        for (int s = 1; s < n_sets; s++) terms.push_back(e_mul(e_sub(Z(s, 0), Z(s - 1, -(bf + 1))), l0));
        Fr cur_delta = hfr::ONE;
        const Fr DELTA = fr_from_hex("0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2");
        for (int s = 0; s < n_sets; s++) {
            ExprP left = Z(s, 1), right = Z(s, 0);
            for (int j = 0; j < sets[s].second; j++) {
                const auto &pc = cs.perm_cols[sets[s].first + j];
                ExprP val = e_col(em.col(pc.first, pc.second), 0);
                left = e_mul(left, e_add(e_add(val, e_mul(ec(beta), e_col(E_SIGMA + sets[s].first + j, 0))), ec(gamma)));
                right = e_mul(right, e_add(e_add(val, e_mul(ec(hfr::mul(cur_delta, beta)), e_col(E_X, 0))), ec(gamma)));
                cur_delta = hfr::mul(cur_delta, DELTA);
            }
            terms.push_back(e_mul(e_sub(left, right), lact));
        }
    }
    for (size_t li = 0; li < n_lookups; li++) {
        ExprP z0 = e_col(E_LK + 3 * li, 0), z1 = e_col(E_LK + 3 * li, 1);
        ExprP a0 = e_col(E_LK + 3 * li + 1, 0), am1 = e_col(E_LK + 3 * li + 1, -1), s0 = e_col(E_LK + 3 * li + 2, 0);
        auto compress = [&](const std::vector<json::ValueP> &exprs) {
            ExprP acc = nullptr;
            for (auto &e : exprs) {
                ExprP b = bind_expr(*e, em);
                acc = acc ? e_add(e_mul(acc, ec(theta)), b) : b;
            }
            return acc;
        };
        ExprP cin = compress(cs.lookups[li].input), ctab = compress(cs.lookups[li].table);
        ExprP a_minus_s = e_sub(a0, s0);
        terms.push_back(e_mul(e_sub(one, z0), l0));
        terms.push_back(e_mul(e_sub(e_mul(z0, z0), z0), llast));
        terms.push_back(e_mul(e_sub(e_mul(z1, e_mul(e_add(a0, ec(beta)), e_add(s0, ec(gamma)))), e_mul(z0, e_mul(e_add(cin, ec(beta)), e_add(ctab, ec(gamma))))), lact));
        terms.push_back(e_mul(a_minus_s, l0));
        terms.push_back(e_mul(e_mul(a_minus_s, e_sub(a0, am1)), lact));
    }
    return terms;
}

// The quotient-numerator program of this key for the given challenges.  Structure (DAG, schedule, slots) depends only on the
// constraint system, so it is compiled once with placeholder challenges; afterwards the constant table is patched.
Program h_program_for(const sb_pk *pk, const std::vector<std::pair<int, int>> &sets, size_t n_lookups, const Fr &theta, const Fr &beta, const Fr &gamma, const Fr &y) {
    const ConstraintSystem &cs = pk->cs;
    const Fr DELTA = fr_from_hex("0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2");
    sb_pk::HProgramCache &hc = pk->hcache;
    std::lock_guard<std::mutex> lk(hc.mu);
    if (!hc.valid) {
        // placeholders: fixed, pairwise distinct, astronomically unlikely to collide with a constant of the circuit
        const Fr t0 = fr_from_hex("0x1b3c5d7e9fa1c3e5071929bb4d6f8192a3b5c7d9ebfd0f21334557697b8d9fb1");
        const Fr b0 = fr_from_hex("0x0a1c2e40526476889aacbed0e2f40618293b4d5f718395a7b9cbddef01132537");
        const Fr g0 = fr_from_hex("0x2f1d0bf9e7d5c3b1a08f7e6d5c4b3a291807f6e5d4c3b2a1908f7e6d5c4b3a29");
        const Fr y0 = fr_from_hex("0x13579bdf02468ace13579bdf02468ace13579bdf02468ace13579bdf02468acf");
        const std::vector<ExprP> terms = h_terms(cs, pk->P, sets, n_lookups, t0, b0, g0);
        fr_t yd = to_dev(y0);
        hc.prog = compile_terms(terms, &yd);
        std::vector<Fr> db(pk->P), yp(terms.size() + 1);
        Fr d = hfr::ONE;
        for (int j = 0; j < pk->P; j++) { db[j] = hfr::mul(d, b0); d = hfr::mul(d, DELTA); }
        yp[0] = hfr::ONE;
        for (size_t e = 1; e < yp.size(); e++) yp[e] = hfr::mul(yp[e - 1], y0);
        for (uint32_t ci = 0; ci < hc.prog.consts.size(); ci++) {
            const Fr c = to_host(hc.prog.consts[ci]);
            if (c == t0) { hc.theta.push_back(ci); continue; }
            if (c == b0) { hc.beta.push_back(ci); continue; }
            if (c == g0) { hc.gamma.push_back(ci); continue; }
            bool hit = false;
            for (int j = 1; j < pk->P && !hit; j++)
                if (c == db[j]) { hc.delta_beta.push_back({ci, j}); hit = true; }
            for (size_t e = 1; e < yp.size() && !hit; e++)
                if (c == yp[e]) { hc.ypow.push_back({ci, (int)e}); hit = true; }
        }
        hc.valid = true;
    }
    Program p = hc.prog;
    for (uint32_t ci : hc.theta) p.consts[ci] = to_dev(theta);
    for (uint32_t ci : hc.beta) p.consts[ci] = to_dev(beta);
    for (uint32_t ci : hc.gamma) p.consts[ci] = to_dev(gamma);
    if (!hc.delta_beta.empty()) {
        std::vector<Fr> db(pk->P);
        Fr d = hfr::ONE;
        for (int j = 0; j < pk->P; j++) { db[j] = hfr::mul(d, beta); d = hfr::mul(d, DELTA); }
        for (auto &e : hc.delta_beta) p.consts[e.first] = to_dev(db[e.second]);
    }
    if (!hc.ypow.empty()) {
        int emax = 0;
        for (auto &e : hc.ypow) emax = std::max(emax, e.second);
        std::vector<Fr> yp(emax + 1);
        yp[0] = hfr::ONE;
        for (int e = 1; e <= emax; e++) yp[e] = hfr::mul(yp[e - 1], y);
        for (auto &e : hc.ypow) p.consts[e.first] = to_dev(yp[e.second]);
    }
    return p;
}

// ------------------------------------------------------------------ sharded proving
// One process per GPU; every rank runs the whole transcript in lock step on replicated inputs and owns
//   * a contiguous base range of every MSM (partial commitments are all-gathered as 64-byte points and added on the host),
//   * 2^(ext_k-k) / world cosets of the extended domain: extended index i = (i >> rs_log) * 2^rs_log + j lies on the coset
//     zeta * ext_omega^j * H, rotations move inside a coset, so coset NTTs, evaluate_h and the division by t(X) need no exchange;
//     the quotient's coset-major values are all-gathered once (device buffer) before the extended inverse NTT.
// Table bases on a power-of-two number of ranks can also be sharded by BUCKET RESIDUE (rank r keeps the digits whose bucket index is r mod world):
// accumulation, sort atomics AND the bucket reduction divide by world exactly (15 windows do not divide by 8, and the bucket reduction of a window
// shard is as large as the single-GPU one), but every rank recodes every window of every scalar.  Measured (profiles/r02u, r02s, r02y): the residue
// split wins where the bucket sets are large (c >= 20, i.e. k >= 23: 8 GPUs 134.5 -> 117.6 ms, 2 GPUs 297 -> 284 ms) and on 2 GPUs at k = 20
// (38.8 -> 38.1 ms); at k = 20 on 8 GPUs the window split is faster (18.6 vs 19.5 ms) -- the replicated recoding outweighs a 0.2 ms bucket tree.
// Returns log2(world) for the residue split, 0 for the window split.  SB_SHARD_MSM_BY_WINDOW / SB_SHARD_MSM_BY_RESIDUE force one.
static uint32_t msm_residue_log(sb_ctx *ctx, const sb_comm *comm, const MsmTables *tabs, size_t n) {
    if (!comm || comm->world <= 1 || !tabs || ctx->tune.shard_msm_by_window || ctx->tune.msm_no_bucket_tree) return 0;
    const uint32_t w = (uint32_t)comm->world;
    if (w & (w - 1)) return 0;
    uint32_t lg = 0;
    while ((1u << lg) < w) lg++;
    if (tabs->c < lg + 9) return 0;   // at least 256 buckets per rank
    if (ctx->tune.shard_msm_by_residue) return lg;
    return (tabs->c >= 20 || (w == 2 && n >= ((size_t)1 << 19))) ? lg : 0;
}

int32_t msm_commit(sb_ctx *ctx, const sb_comm *comm, const sb_srs *srs, int basis, const void *d_scalars, size_t n, uint8_t out[64], cudaStream_t st) {
    if (!comm || comm->world <= 1) return srs_msm(ctx, srs, basis, d_scalars, n, out, st);
    const void *d_bases = basis == 0 ? srs->d_g : srs->d_g_lagrange;
    const MsmTables *tabs = srs->tab[basis].d_tables ? &srs->tab[basis] : nullptr;
    const uint32_t Wd = (uint32_t)comm->world, r = (uint32_t)comm->rank;
    uint32_t c, W;
    msm_window_shape(ctx, n, &c, &W);
    if (tabs) { c = tabs->c; W = tabs->W; }
    if (W >= Wd && !ctx->tune.shard_msm_by_range) {
        // by signed-digit window: rank r accumulates windows [r W / world, (r + 1) W / world) over ALL bases (level-1 additions, sort and bucket
        // reduction all divide by world).  Plain bases: the W window sums (128 B XYZZ each) meet on the host and are folded by Horner there.
        // Table bases: every rank's partial already carries its 2^(c w) factors, so the host adds world points.
        const uint32_t lo = r * W / Wd, hi = (r + 1) * W / Wd;
        if (tabs) {
            uint8_t mine[128];
            std::vector<uint8_t> all((size_t)Wd * 128);
            const uint32_t rl = msm_residue_log(ctx, comm, tabs, n);
            if (rl) SB_TRY(msm_run_tables_batch_windows(ctx, tabs, d_scalars, n, 1, 0, (int32_t)W, mine, st, nullptr, r, rl));
            else SB_TRY(msm_run_tables(ctx, tabs, d_scalars, n, (int32_t)lo, (int32_t)hi, mine, st));
            if (comm->allgather_host(comm->user, mine, all.data(), 128) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
            msm_fold_windows(all.data(), Wd, 0, out);
            return SB_OK;
        }
        const uint32_t per = (W + Wd - 1) / Wd;  // slots per rank in the gathered buffer (ragged tails stay zero = identity)
        std::vector<uint8_t> mine((size_t)per * 128, 0), all((size_t)Wd * per * 128), win((size_t)W * 128);
        SB_TRY(msm_run_windows(ctx, d_bases, d_scalars, n, lo, hi, mine.data(), st));
        if (comm->allgather_host(comm->user, mine.data(), all.data(), (size_t)per * 128) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
        for (uint32_t q = 0; q < Wd; q++) {
            const uint32_t qlo = q * W / Wd, qhi = (q + 1) * W / Wd;
            memcpy(win.data() + (size_t)qlo * 128, all.data() + (size_t)q * per * 128, (size_t)(qhi - qlo) * 128);
        }
        msm_fold_windows(win.data(), W, c, out);
        return SB_OK;
    }
    // by base range (north_star): 64-byte partial commitments added on the host
    const size_t per = n / Wd, lo = r * per, hi = (r + 1 == Wd) ? n : lo + per;
    uint8_t part[64];
    SB_TRY(msm_run(ctx, (const uint8_t *)d_bases + lo * 64, (const uint8_t *)d_scalars + lo * 32, hi - lo, part, st));
    std::vector<uint8_t> all((size_t)Wd * 64);
    if (comm->allgather_host(comm->user, part, all.data(), 64) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
    return sb_g1_sum_affine(all.data(), Wd, out);
}

// m commitments over contiguous scalar vectors: one batched launch set on a single GPU, the sharded path one by one
int32_t msm_commit_batch(sb_ctx *ctx, const sb_comm *comm, const sb_srs *srs, int basis, const void *d_scalars, size_t n, uint32_t m, uint8_t *out, cudaStream_t st) {
    if (!comm || comm->world <= 1) return srs_msm_batch(ctx, srs, basis, d_scalars, n, m, out, st);
    const MsmTables *tabs = srs->tab[basis].d_tables ? &srs->tab[basis] : nullptr;
    const uint32_t Wd = (uint32_t)comm->world, r = (uint32_t)comm->rank;
    if (tabs && tabs->W >= Wd && (uint64_t)tabs->W * n * m < (1ull << 32) - 8 && !ctx->tune.shard_msm_by_range) {
        // all m commitments, this rank's windows, ONE launch set and ONE exchange of m XYZZ partials per rank
        const uint32_t lo = r * tabs->W / Wd, hi = (r + 1) * tabs->W / Wd;
        std::vector<uint8_t> mine((size_t)m * 128), all((size_t)Wd * m * 128), col((size_t)Wd * 128);
        const uint32_t rl = msm_residue_log(ctx, comm, tabs, n);
        if (rl) SB_TRY(msm_run_tables_batch_windows(ctx, tabs, d_scalars, n, m, 0, (int32_t)tabs->W, mine.data(), st, nullptr, r, rl));
        else SB_TRY(msm_run_tables_batch_windows(ctx, tabs, d_scalars, n, m, (int32_t)lo, (int32_t)hi, mine.data(), st));
        if (comm->allgather_host(comm->user, mine.data(), all.data(), (size_t)m * 128) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
        for (uint32_t j = 0; j < m; j++) {
            for (uint32_t q = 0; q < Wd; q++) memcpy(col.data() + (size_t)q * 128, all.data() + ((size_t)q * m + j) * 128, 128);
            msm_fold_windows(col.data(), Wd, 0, out + (size_t)j * 64);
        }
        return SB_OK;
    }
    for (uint32_t j = 0; j < m; j++) SB_TRY(msm_commit(ctx, comm, srs, basis, (const uint8_t *)d_scalars + (size_t)j * n * 32, n, out + (size_t)j * 64, st));
    return SB_OK;
}

// mixed-basis batch (vector j over basis basis_of[j] of one table slab): one launch set on a single GPU, window-sharded with one exchange otherwise
int32_t msm_commit_batch_mixed(sb_ctx *ctx, const sb_comm *comm, const sb_srs *srs, const void *d_scalars, size_t n, uint32_t m, const uint8_t *basis_of, uint8_t *out,
                               cudaStream_t st) {
    const MsmTables *t0 = &srs->tab[0], *t1 = &srs->tab[1];
    if (!comm || comm->world <= 1) return msm_run_tables_batch_mixed(ctx, t0, t1, d_scalars, n, m, basis_of, out, st);
    const uint32_t Wd = (uint32_t)comm->world, r = (uint32_t)comm->rank;
    SB_REQUIRE(m <= 8 && t0->W >= Wd && (uint64_t)t0->W * n * m < (1ull << 32) - 8, "msm_commit_batch_mixed: shape");
    uint32_t off[8];
    for (uint32_t j = 0; j < m; j++) off[j] = basis_of[j] ? (uint32_t)((uint64_t)t0->W * t0->stride) : 0u;
    const uint32_t lo = r * t0->W / Wd, hi = (r + 1) * t0->W / Wd;
    std::vector<uint8_t> mine((size_t)m * 128), all((size_t)Wd * m * 128), col((size_t)Wd * 128);
    const uint32_t rl = msm_residue_log(ctx, comm, t0, n);
    if (rl) SB_TRY(msm_run_tables_batch_windows(ctx, t0, d_scalars, n, m, 0, (int32_t)t0->W, mine.data(), st, off, r, rl));
    else SB_TRY(msm_run_tables_batch_windows(ctx, t0, d_scalars, n, m, (int32_t)lo, (int32_t)hi, mine.data(), st, off));
    if (comm->allgather_host(comm->user, mine.data(), all.data(), (size_t)m * 128) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
    for (uint32_t j = 0; j < m; j++) {
        for (uint32_t q = 0; q < Wd; q++) memcpy(col.data() + (size_t)q * 128, all.data() + ((size_t)q * m + j) * 128, 128);
        msm_fold_windows(col.data(), Wd, 0, out + (size_t)j * 64);
    }
    return SB_OK;
}

// values of the polynomial `d_coeff` (n coefficients) on the coset slots s0, s0 + step, ... (count of them): NTT_n(coeff[m] * g_s^m), the scaling
// applied as the coefficients are loaded; one batched launch set, output b at d_out + b * out_stride elements
int32_t coset_values(sb_ctx *ctx, const sb_pk *pk, const void *d_coeff, uint32_t s0, uint32_t step, uint32_t count, void *d_out, size_t out_stride, cudaStream_t st) {
    if (count == 0) return SB_OK;
    NttFuse f;
    f.pre_vec = (const uint8_t *)pk->coset_pows + (size_t)s0 * pk->n * 32;
    f.batch = count;
    f.src_stride = 0; f.pre_stride = (uint64_t)step * pk->n; f.dst_stride = out_stride;
    return ntt_run_fused(ctx, d_coeff, d_out, (const uint8_t *)pk->dom->omega.v, pk->k, &f, st);
}

// where the assigned advice cells come from: dense host columns (A x n x 32 B), dense device columns (left untouched), or the non-zero cells only
struct Witness {
    const uint8_t *host = nullptr;
    const void *dev = nullptr;
    const uint32_t *cells = nullptr;   // (col, row) pairs
    const uint8_t *values = nullptr;   // 32 B each
    size_t n_cells = 0;
    bool sparse = false;
};

int32_t create_proof_impl(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_inst, const Witness &wit, ChaCha20Rng &rng,
                          Transcript &tr, cudaStream_t st) {
    const ConstraintSystem &cs = pk->cs;
    const sb_domain *d = pk->dom;
    const size_t n = pk->n;
    const int A = cs.A, F = cs.F, bf = cs.blinding, P = pk->P;
    const size_t usable = n - (size_t)(bf + 1);
    SB_REQUIRE(n_inst <= usable, "create_proof: too many instance values");
    const Fr omega = to_host(d->omega), omega_inv = to_host(d->omega_inv);
    uint8_t pt[64];
    // wall-clock per stage (every stage ends on a stream synchronisation: MSM results and challenges come back to the host)
    int stage = 0;
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&]() {
        sync_stream(ctx, st);
        auto now = std::chrono::steady_clock::now();
        if (stage < 12) ctx->last_proof_stage_ms[stage++] = std::chrono::duration<float, std::milli>(now - t_prev).count();
        t_prev = now;
    };
    for (int i = 0; i < 12; i++) ctx->last_proof_stage_ms[i] = 0;
    for (int i = 0; i < 5; i++) ctx->acc_msm_ms[i] = 0;
    ctx->acc_msm_digits = 0;
    ctx->acc_msm_sets = 0;
    ctx->acc_msm_d2h = 0;
    // Side stream: coeff_to_extended of the per-proof polynomials depends on no later challenge, so it is enqueued as soon as a
    // polynomial exists and runs under the latency-bound tails of the commitments on the main stream; joined before evaluate_h.
    const bool use_side = ctx->side_stream && st == ctx->stream && !ctx->tune.no_side_stream;
    cudaStream_t st2 = use_side ? ctx->side_stream : st;
    auto side_after_main = [&]() -> int32_t {  // everything enqueued on st so far happens before later work on st2
        if (!use_side) return SB_OK;
        SB_CUDA_TRY(cudaEventRecord(ctx->side_ev[0], st));
        SB_CUDA_TRY(cudaStreamWaitEvent(st2, ctx->side_ev[0], 0));
        return SB_OK;
    };
    auto main_after_side = [&]() -> int32_t {
        if (!use_side) return SB_OK;
        SB_CUDA_TRY(cudaEventRecord(ctx->side_ev[1], st2));
        SB_CUDA_TRY(cudaStreamWaitEvent(st, ctx->side_ev[1], 0));
        return SB_OK;
    };
    struct SideJoin {  // an early return must not leave side work in flight on buffers the next call reuses
        cudaStream_t s; bool on;
        ~SideJoin() { if (on) cudaStreamSynchronize(s); }
    } side_join{st2, use_side};
    // The quotient's cosets (sb_pk: n_cos = j - 1 cosets of the size-n subgroup), dealt round-robin to the ranks of a sharded proof (slot s belongs
    // to rank s mod world; a single GPU owns them all), and the slab that receives the per-proof polynomials' values on the owned cosets
    // (index [owned coset][polynomial][row]); polynomial order: advice | instance | permutation Z | per lookup (Z, A', S')
    const int n_sets_all = (P + (cs.degree - 2) - 1) / (cs.degree - 2);
    const uint32_t n_cos = (uint32_t)pk->n_cos;
    const uint32_t world = comm ? (uint32_t)comm->world : 1u, rank = comm ? (uint32_t)comm->rank : 0u;
    std::vector<uint32_t> own;  // coset slots of this rank
    for (uint32_t s2 = rank; s2 < n_cos; s2 += world) own.push_back(s2);
    const uint32_t co_per = (uint32_t)own.size();
    const size_t n_dyn = (size_t)A + 1 + (size_t)n_sets_all + 3 * cs.lookups.size();
    uint8_t *d_dyn = nullptr;
    SB_TRY(scratch_get(ctx, "pf_coset_dyn", (size_t)(co_per ? co_per : 1) * n_dyn * n * 32, (void **)&d_dyn));
    auto dyn_slot = [&](uint32_t jl, size_t q) -> void * { return d_dyn + ((size_t)jl * n_dyn + q) * n * 32; };
    auto side_cosets = [&](size_t q, const void *d_coeff) -> int32_t {  // values of one polynomial on every owned coset, on the side stream
        return coset_values(ctx, pk, d_coeff, rank, world, co_per, dyn_slot(0, q), n_dyn * n, st2);
    };

    // lagrange_to_coeff of a column every rank holds in full: a distributed four-step NTT at the largest k (each rank 1 / world of both passes, an
    // all-to-all between them, an all-gather after), the local fused transform otherwise.  The choice depends on (k, world) only: lock step.
    const bool dist_ntt = comm && world > 1 && comm->alltoall_dev && (world & (world - 1)) == 0 && pk->k >= (uint32_t)ctx->tune.dist_ntt_min_k && pk->k >= 16 && pk->k <= 24;
    auto l2c_repl = [&](void *d_poly, cudaStream_t s) -> int32_t {
        if (dist_ntt) return ntt_run_dist(ctx, comm, d_poly, (const uint8_t *)d->omega_inv.v, pk->k, &d->ifft_divisor, s);
        return dom_l2c(ctx, d, d_poly, s);
    };

    // ---- transcript preamble
    tr.common_scalar(pk->transcript_repr);
    std::vector<Fr> inst(n_inst);
    memcpy(inst.data(), instances, n_inst * 32);
    for (const Fr &v : inst) tr.common_scalar(v);

    // ---- device buffers of this proof
    void *d_inst, *d_inst_poly;
    SB_TRY(scratch_get(ctx, "pf_inst", n * 32, &d_inst));
    SB_TRY(scratch_get(ctx, "pf_inst_poly", n * 32, &d_inst_poly));
    SB_CUDA_TRY(cudaMemsetAsync(d_inst, 0, n * 32, st));
    SB_CUDA_TRY(cudaMemcpyAsync(d_inst, inst.data(), n_inst * 32, cudaMemcpyHostToDevice, st));
    // few instance values: their coset values come from rotations of the key's l_0 coset (prover_kernels.cu::instance_coset), no transform
    const bool inst_direct = n_inst <= 12 && !ctx->tune.no_inst_direct;
    if (!inst_direct) {
        SB_CUDA_TRY(cudaMemcpyAsync(d_inst_poly, d_inst, n * 32, cudaMemcpyDeviceToDevice, st));
        SB_TRY(dom_l2c(ctx, d, d_inst_poly, st));
    }
    std::vector<void *> adv(A), adv_poly(A);
    void *d_random_early = nullptr;
    {
        uint8_t *base_v, *base_p;
        SB_TRY(scratch_get(ctx, "pf_adv", (size_t)(A + 1) * n * 32, (void **)&base_v));  // + 1: the vanishing argument's random polynomial (early commitment)
        SB_TRY(scratch_get(ctx, "pf_adv_poly", (size_t)A * n * 32, (void **)&base_p));
        for (int c = 0; c < A; c++) {
            adv[c] = base_v + (size_t)c * n * 32;
            adv_poly[c] = base_p + (size_t)c * n * 32;
        }
        d_random_early = base_v + (size_t)A * n * 32;
        const size_t adv_bytes = (size_t)A * n * 32;
        if (wit.sparse) {
            // only the assigned cells cross PCIe (a few thousand for MstInclusionCircuit, whatever k is): zero the columns, upload, scatter
            SB_REQUIRE(A <= 32, "sparse witness: at most 32 advice columns");
            for (size_t i = 0; i < wit.n_cells; i++)
                SB_REQUIRE((int)wit.cells[2 * i] < A && wit.cells[2 * i + 1] < n, "sparse witness: cell out of range");
            void *d_cells, *d_vals;
            SB_TRY(scratch_get(ctx, "pf_wit_cells", wit.n_cells * 8 + 16, &d_cells));
            SB_TRY(scratch_get(ctx, "pf_wit_vals", wit.n_cells * 32 + 32, &d_vals));
            SB_CUDA_TRY(cudaMemsetAsync(base_v, 0, adv_bytes, st));
            if (wit.n_cells) {
                SB_CUDA_TRY(cudaMemcpyAsync(d_cells, wit.cells, wit.n_cells * 8, cudaMemcpyHostToDevice, st));
                SB_CUDA_TRY(cudaMemcpyAsync(d_vals, wit.values, wit.n_cells * 32, cudaMemcpyHostToDevice, st));
                SB_TRY(scatter_cells(ctx, adv.data(), (uint32_t)A, d_cells, d_vals, wit.n_cells, st));
            }
        } else if (wit.dev) {
            SB_CUDA_TRY(cudaMemcpyAsync(base_v, wit.dev, adv_bytes, cudaMemcpyDeviceToDevice, st));  // the blinding rows are written into the copy
        } else if (comm && comm->world > 1 && adv_bytes % ((size_t)comm->world * 256) == 0) {
            // every rank holds the same host witness: upload 1 / world of it over PCIe and gather the rest over NVLink
            const size_t per = adv_bytes / (size_t)comm->world, off = per * (size_t)comm->rank;
            SB_CUDA_TRY(cudaMemcpyAsync(base_v + off, wit.host + off, per, cudaMemcpyHostToDevice, st));
            SB_CUDA_TRY(sync_stream(ctx, st));
            if (comm->allgather_dev(comm->user, base_v, per, (void *)st) != 0) { set_last_error("sb_comm.allgather_dev failed"); return SB_ERR_ARG; }
        } else {
            SB_CUDA_TRY(cudaMemcpyAsync(base_v, wit.host, adv_bytes, cudaMemcpyHostToDevice, st));
        }
    }
    // blinding rows, blinds (drawn; KZG ignores them), commitments
    for (int c = 0; c < A; c++) {
        std::vector<Fr> blind(n - usable);
        for (Fr &b : blind) b = rng.next_fr();
        SB_TRY(upload_frs(ctx, (uint8_t *)adv[c] + usable * 32, blind, st));
    }
    for (int c = 0; c < A; c++) (void)rng.next_fr();
    // The commitment below is over the Lagrange basis, so the coefficient forms (needed by the coset transforms and the evaluations only) are
    // made on the side stream too: in a sharded proof the commitment is latency-bound and these replicated transforms hide under it.
    SB_TRY(side_after_main());  // advice coefficient forms, advice / instance cosets on the side stream, under the commitment below
    if (dist_ntt) {
        for (int c = 0; c < A; c++) {
            SB_CUDA_TRY(cudaMemcpyAsync(adv_poly[c], adv[c], n * 32, cudaMemcpyDeviceToDevice, st2));
            SB_TRY(l2c_repl(adv_poly[c], st2));
        }
    } else {   // lagrange_to_coeff of the A advice columns: one batched out-of-place inverse transform (n^-1 folded in)
        NttFuse f;
        f.has_scale = true;
        f.scale = d->ifft_divisor;
        f.batch = (uint32_t)A;
        f.src_stride = n; f.dst_stride = n;
        SB_TRY(ntt_run_fused(ctx, adv[0], adv_poly[0], (const uint8_t *)d->omega_inv.v, pk->k, &f, st2));
    }
    for (int c = 0; c < A; c++) SB_TRY(side_cosets((size_t)c, adv_poly[c]));
    if (inst_direct) {
        for (uint32_t jl = 0; jl < co_per; jl++)
            SB_TRY(instance_coset(ctx, (const uint8_t *)pk->l0 + (size_t)own[jl] * n * 32, d_inst, (uint32_t)n_inst, n, dyn_slot(jl, (size_t)A), st2));
    } else {
        SB_TRY(side_cosets((size_t)A, d_inst_poly));
    }
    // The vanishing argument's random polynomial depends on no challenge, only on the RNG stream: a CLONE of the RNG is advanced past every
    // draw that precedes its seed (all data-independent counts), so the polynomial can be generated now and committed in the SAME launch set
    // as the advice columns (mixed-basis batch: advice over the Lagrange tables, the random polynomial over the monomial tables).  The main
    // RNG still makes its draws at the usual place, so the proof bytes do not change.
    uint8_t random_commitment[64];
    bool random_early = false;
    {
        const sb_srs *srs = pk->srs;
        const bool shard_ok = !comm || comm->world <= 1 || (srs->tab[0].W >= (uint32_t)comm->world && !ctx->tune.shard_msm_by_range);
        const bool mixed_ok = shard_ok && A + 1 <= 8 && srs->tab_slab && srs->tab[0].d_tables && srs->tab[1].d_tables && !ctx->tune.no_early_random;
        std::vector<uint8_t> pts((size_t)(A + 1) * 64);
        if (mixed_ok) {
            ChaCha20Rng ahead = rng;
            const size_t n_z_cols = (size_t)((P + (cs.degree - 2) - 1) / (cs.degree - 2)) + cs.lookups.size();
            const size_t draws = cs.lookups.size() * (2 * (n - usable) + 2) + n_z_cols * ((size_t)bf + 1);
            for (size_t q = 0; q < draws; q++) (void)ahead.next_fr();
            uint8_t seed[32];
            ahead.fill_bytes(seed, 32);
            ChaCha20Rng child;
            child.seed(seed);
            SB_TRY(chacha_fr_fill(ctx, child.key, 0, d_random_early, n, st));
            std::vector<uint8_t> basis_of((size_t)A + 1, 1);
            basis_of[A] = 0;
            SB_TRY(msm_commit_batch_mixed(ctx, comm, srs, adv[0], n, (uint32_t)A + 1, basis_of.data(), pts.data(), st));
            memcpy(random_commitment, pts.data() + (size_t)A * 64, 64);
            random_early = true;
        } else {
            SB_TRY(msm_commit_batch(ctx, comm, pk->srs, 1, adv[0], n, (uint32_t)A, pts.data(), st));  // the A columns are contiguous
        }
        for (int c = 0; c < A; c++)
            if (!tr.write_point(pts.data() + (size_t)c * 64)) { set_last_error("advice commitment is the identity"); return SB_ERR_ARG; }
    }
    const Fr theta = tr.squeeze();
    mark();  // [0] instance / advice upload, blinding, lagrange_to_coeff, 3 advice commitments

    // column tables.  Lagrange: advice | fixed | instance | sigma | omega^i | lookup scratch (4)
    ColMap lm;
    lm.advice0 = 0; lm.fixed0 = A; lm.instance0 = A + F;
    const int L_SIGMA = A + F + 1, L_OMEGA = L_SIGMA + P, L_LK = L_OMEGA + 1;
    std::vector<const void *> lcols(L_LK + 4, nullptr);
    for (int c = 0; c < A; c++) lcols[c] = adv[c];
    for (int c = 0; c < F; c++) lcols[A + c] = pk->fixed_values[c];
    lcols[A + F] = d_inst;
    for (int j = 0; j < P; j++) lcols[L_SIGMA + j] = pk->sigma_values[j];
    lcols[L_OMEGA] = pk->omega_pows;

    // ---- lookups: compress, permute, commit
    struct LookupState { void *c_in, *c_tab, *p_in, *p_tab, *in_poly, *tab_poly, *z_poly, *z_vals = nullptr; };
    std::vector<LookupState> lks(cs.lookups.size());
    for (size_t li = 0; li < cs.lookups.size(); li++) {
        LookupState &L = lks[li];
        uint8_t *b;
        SB_TRY(scratch_get(ctx, ("pf_lk" + std::to_string(li)).c_str(), 7 * n * 32, (void **)&b));
        L.c_in = b; L.c_tab = b + n * 32; L.p_in = b + 2 * n * 32; L.p_tab = b + 3 * n * 32;
        L.in_poly = b + 4 * n * 32; L.tab_poly = b + 5 * n * 32; L.z_poly = b + 6 * n * 32;
        std::vector<ExprP> in_terms, tab_terms;
        for (auto &e : cs.lookups[li].input) in_terms.push_back(bind_expr(*e, lm));
        for (auto &e : cs.lookups[li].table) tab_terms.push_back(bind_expr(*e, lm));
        fr_t th = to_dev(theta);
        SB_TRY(expr_eval(ctx, compile_terms(in_terms, &th), lcols, pk->k, 0, L.c_in, st));
        SB_TRY(expr_eval(ctx, compile_terms(tab_terms, &th), lcols, pk->k, 0, L.c_tab, st));
        int32_t rc = lookup_permute(ctx, L.c_in, L.c_tab, n, usable, L.p_in, L.p_tab, st);
        if (rc != SB_OK) return rc;
        std::vector<Fr> blind(n - usable);
        for (Fr &x : blind) x = rng.next_fr();
        SB_TRY(upload_frs(ctx, (uint8_t *)L.p_in + usable * 32, blind, st));
        for (Fr &x : blind) x = rng.next_fr();
        SB_TRY(upload_frs(ctx, (uint8_t *)L.p_tab + usable * 32, blind, st));
        SB_TRY(side_after_main());   // coefficient and coset forms on the side stream, under the (Lagrange-basis) commitment below
        SB_CUDA_TRY(cudaMemcpyAsync(L.in_poly, L.p_in, n * 32, cudaMemcpyDeviceToDevice, st2));
        SB_TRY(l2c_repl(L.in_poly, st2));
        (void)rng.next_fr();
        SB_CUDA_TRY(cudaMemcpyAsync(L.tab_poly, L.p_tab, n * 32, cudaMemcpyDeviceToDevice, st2));
        SB_TRY(l2c_repl(L.tab_poly, st2));
        (void)rng.next_fr();
        SB_TRY(side_cosets((size_t)A + 1 + n_sets_all + 3 * li + 1, L.in_poly));
        SB_TRY(side_cosets((size_t)A + 1 + n_sets_all + 3 * li + 2, L.tab_poly));
        uint8_t pin_tab[128];
        SB_TRY(msm_commit_batch(ctx, comm, pk->srs, 1, L.p_in, n, 2, pin_tab, st));  // p_in and p_tab are adjacent in the lookup scratch block
        if (!tr.write_point(pin_tab) || !tr.write_point(pin_tab + 64)) { set_last_error("lookup commitment is the identity"); return SB_ERR_ARG; }
    }
    const Fr beta = tr.squeeze();
    const Fr gamma = tr.squeeze();
    mark();  // [1] lookup: compress, sort / permute, 2 iNTT, 2 commitments

    // ---- permutation argument and lookup products: all grand-product columns Z are built side by side in one
    //      buffer and committed by ONE batched MSM.  A permutation set's product starts from the previous set's last value; each set is
    //      scanned from 1 and scaled afterwards by the running boundary value (read back once), which is the same field element.
    const int chunk = cs.degree - 2;
    const int n_sets = (P + chunk - 1) / chunk;
    struct PermSet { void *z_poly; int first, count; void *z_vals = nullptr; };
    std::vector<PermSet> psets(n_sets);
    if (n_sets + (int)lks.size() > 0) {  // a circuit with neither copy constraints nor lookups has no grand product to commit
        const int n_z = n_sets + (int)lks.size();
        uint8_t *d_den, *d_num, *d_zall;
        SB_TRY(scratch_get(ctx, "pf_perm_den", (size_t)n_z * n * 32, (void **)&d_den));
        SB_TRY(scratch_get(ctx, "pf_perm_num", (size_t)n_z * n * 32, (void **)&d_num));
        SB_TRY(scratch_get(ctx, "pf_z_all", (size_t)n_z * n * 32, (void **)&d_zall));
        uint8_t *zb;
        SB_TRY(scratch_get(ctx, "pf_perm_polys", (size_t)n_sets * n * 32, (void **)&zb));
        Fr delta_pow = hfr::ONE;
        const Fr DELTA = fr_from_hex("0x09226b6e22c6f0ca64ec26aad4c86e715b5f898e5e963f25870e56bbe533e9a2");
        // numerators and denominators of every grand product side by side, then ONE batch inversion and ONE product pass for all of them
        // Sharded proof: the ratios, their inversion and the running products are computed for rows [r n / world, (r + 1) n / world) only (the
        // expressions read every column at rotation 0); the slices' total products meet on the host, every rank scales its slice by the product
        // of the slices before it, and one in-place all-gather per Z column replicates the result for the (window / residue sharded) commitment.
        uint32_t gp_log = 0;
        while (world > 1 && (1u << gp_log) < world) gp_log++;
        // (one more host exchange and an all-gather per column.  Measured: 2 GPUs k = 23 284.8 -> 281.5 ms but k = 20 37.6 -> 37.8; 8 GPUs k = 23 117.6 ->
        // 111.6, k = 20 18.18 -> 18.03, k = 17 6.9 -> 7.3: on from k = 22, and from k = 20 on four or more ranks)
        const bool gp_rows = world > 1 && (1u << gp_log) == world && comm->allgather_dev && pk->k >= gp_log + 12 && !ctx->tune.no_grand_shard &&
                             (pk->k >= 22 || (world >= 4 && pk->k >= 20) || ctx->tune.grand_shard);
        const size_t gp_cnt = gp_rows ? n >> gp_log : n, gp_off = gp_rows ? gp_cnt * rank : 0;
        std::vector<const void *> lcols_sl;
        auto ratio_terms = [&](ExprP den, ExprP num, int z) -> int32_t {
            if (gp_rows) {
                lcols_sl.resize(lcols.size());
                for (size_t c = 0; c < lcols.size(); c++) lcols_sl[c] = lcols[c] ? (const uint8_t *)lcols[c] + gp_off * 32 : nullptr;
                SB_TRY(expr_eval(ctx, compile_terms({den}, nullptr), lcols_sl, pk->k - gp_log, 0, d_den + ((size_t)z * n + gp_off) * 32, st));
                return expr_eval(ctx, compile_terms({num}, nullptr), lcols_sl, pk->k - gp_log, 0, d_num + ((size_t)z * n + gp_off) * 32, st);
            }
            SB_TRY(expr_eval(ctx, compile_terms({den}, nullptr), lcols, pk->k, 0, d_den + (size_t)z * n * 32, st));
            return expr_eval(ctx, compile_terms({num}, nullptr), lcols, pk->k, 0, d_num + (size_t)z * n * 32, st);
        };
        for (int s = 0; s < n_sets; s++) {
            PermSet &S = psets[s];
            S.first = s * chunk;
            S.count = std::min(chunk, P - S.first);
            S.z_poly = zb + (size_t)s * n * 32;
            ExprP den = nullptr, num = nullptr;
            for (int j = 0; j < S.count; j++) {
                const auto &pc = cs.perm_cols[S.first + j];
                ExprP val = e_col(lm.col(pc.first, pc.second), 0);
                ExprP dterm = e_add(e_add(e_mul(ec(beta), e_col(L_SIGMA + S.first + j, 0)), ec(gamma)), val);
                ExprP nterm = e_add(e_add(e_mul(ec(hfr::mul(delta_pow, beta)), e_col(L_OMEGA, 0)), ec(gamma)), val);
                den = den ? e_mul(den, dterm) : dterm;
                num = num ? e_mul(num, nterm) : nterm;
                delta_pow = hfr::mul(delta_pow, DELTA);
            }
            SB_TRY(ratio_terms(den, num, s));
        }
        for (size_t li = 0; li < lks.size(); li++) {
            LookupState &L = lks[li];
            lcols[L_LK + 0] = L.c_in; lcols[L_LK + 1] = L.c_tab; lcols[L_LK + 2] = L.p_in; lcols[L_LK + 3] = L.p_tab;
            ExprP den = e_mul(e_add(e_col(L_LK + 2, 0), ec(beta)), e_add(e_col(L_LK + 3, 0), ec(gamma)));
            ExprP num = e_mul(e_add(e_col(L_LK + 0, 0), ec(beta)), e_add(e_col(L_LK + 1, 0), ec(gamma)));
            SB_TRY(ratio_terms(den, num, n_sets + (int)li));
        }
        std::vector<Fr> local_last(n_sets);
        if (gp_rows) {
            // per column: this rank's rows.  Host record per rank: n_z slice totals, then (last rank) the un-chained boundary values z_s[n - bf - 1]
            std::vector<Fr> rec(2 * (size_t)n_z, hfr::ZERO), last_z(n_z), last_r(n_z);
            const size_t b_row = n - (size_t)bf - 1;   // boundary row: lies in the last rank's slice (bf + 1 < n / world)
            for (int z = 0; z < n_z; z++) {
                uint8_t *den_sl = d_den + ((size_t)z * n + gp_off) * 32, *z_sl = d_zall + ((size_t)z * n + gp_off) * 32;
                SB_TRY(fr_batch_invert(ctx, den_sl, gp_cnt, st));
                SB_TRY(fp_vec_op(ctx, 0, 0, den_sl, d_num + ((size_t)z * n + gp_off) * 32, den_sl, gp_cnt, st));
                SB_TRY(fr_running_product(ctx, den_sl, gp_cnt, fr_t::one(), z_sl, gp_cnt, st));
                SB_CUDA_TRY(cudaMemcpyAsync(&last_z[z], z_sl + (gp_cnt - 1) * 32, 32, cudaMemcpyDeviceToHost, st));
                SB_CUDA_TRY(cudaMemcpyAsync(&last_r[z], den_sl + (gp_cnt - 1) * 32, 32, cudaMemcpyDeviceToHost, st));
                if (rank + 1 == world) SB_CUDA_TRY(cudaMemcpyAsync(&rec[(size_t)n_z + z], d_zall + ((size_t)z * n + b_row) * 32, 32, cudaMemcpyDeviceToHost, st));
            }
            SB_CUDA_TRY(sync_stream(ctx, st));
            for (int z = 0; z < n_z; z++) rec[z] = hfr::mul(last_z[z], last_r[z]);
            std::vector<Fr> all(2 * (size_t)n_z * world);
            if (comm->allgather_host(comm->user, rec.data(), all.data(), rec.size() * 32) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
            Fr carry = hfr::ONE;   // chains the permutation sets: set s starts from the (chained) boundary value of set s - 1
            for (int z = 0; z < n_z; z++) {
                Fr before = hfr::ONE, before_last = hfr::ONE;   // product of the slices before this rank / before the last rank
                for (uint32_t q = 0; q + 1 < world; q++) {
                    if (q < rank) before = hfr::mul(before, all[(size_t)q * 2 * n_z + z]);
                    before_last = hfr::mul(before_last, all[(size_t)q * 2 * n_z + z]);
                }
                const bool perm = z < n_sets;
                const Fr scale = perm ? hfr::mul(before, carry) : before;
                if (!(scale == hfr::ONE)) SB_TRY(fr_scale(ctx, d_zall + ((size_t)z * n + gp_off) * 32, gp_cnt, to_dev(scale), st));
                if (perm) {
                    local_last[z] = hfr::mul(before_last, all[(size_t)(world - 1) * 2 * n_z + n_z + z]);   // un-chained z_s[n - bf - 1]
                    carry = hfr::mul(carry, local_last[z]);
                }
            }
        } else {
        SB_TRY(fr_batch_invert(ctx, d_den, (size_t)n_z * n, st));
        SB_TRY(fp_vec_op(ctx, 0, 0, d_den, d_num, d_den, (size_t)n_z * n, st));
        for (int z = 0; z < n_z; z++) SB_TRY(fr_running_product(ctx, d_den + (size_t)z * n * 32, n, fr_t::one(), d_zall + (size_t)z * n * 32, n, st));
        // boundary values of the un-chained permutation products, one read-back
        for (int s = 0; s + 1 < n_sets; s++)
            SB_CUDA_TRY(cudaMemcpyAsync(&local_last[s], d_zall + ((size_t)s * n + (n - (size_t)bf - 1)) * 32, 32, cudaMemcpyDeviceToHost, st));
        if (n_sets > 1) SB_CUDA_TRY(sync_stream(ctx, st));
        Fr carry = hfr::ONE;
        for (int s = 1; s < n_sets; s++) {
            carry = hfr::mul(carry, local_last[s - 1]);
            SB_TRY(fr_scale(ctx, d_zall + (size_t)s * n * 32, n, to_dev(carry), st));
        }
        }
        // blinding rows and blinds in halo2's draw order: every permutation set, then every lookup product
        for (int z = 0; z < n_z; z++) {
            std::vector<Fr> blind(bf);
            for (Fr &x : blind) x = rng.next_fr();
            SB_TRY(upload_frs(ctx, d_zall + ((size_t)z * n + (n - bf)) * 32, blind, st));
            (void)rng.next_fr();
        }
        if (gp_rows) {   // the blinding rows above landed in every rank's copy of the last slice; the gather brings the last rank's (identical) ones
            SB_CUDA_TRY(sync_stream(ctx, st));
            for (int z = 0; z < n_z; z++)
                if (comm->allgather_dev(comm->user, d_zall + (size_t)z * n * 32, gp_cnt * 32, (void *)st) != 0) { set_last_error("sb_comm.allgather_dev failed"); return SB_ERR_ARG; }
        }
        // coefficient and coset forms of every Z on the side stream (sharded proving: coefficient form only), under the batched commitment
        SB_TRY(side_after_main());
        for (int s = 0; s < n_sets; s++) {
            PermSet &S = psets[s];
            S.z_vals = d_zall + (size_t)s * n * 32;   // the values on the subgroup stay valid to the end of the proof (SHPLONK reads them)
            SB_CUDA_TRY(cudaMemcpyAsync(S.z_poly, d_zall + (size_t)s * n * 32, n * 32, cudaMemcpyDeviceToDevice, st2));
            SB_TRY(l2c_repl(S.z_poly, st2));
            SB_TRY(side_cosets((size_t)A + 1 + s, S.z_poly));
        }
        for (size_t li = 0; li < lks.size(); li++) {
            LookupState &L = lks[li];
            L.z_vals = d_zall + (size_t)(n_sets + (int)li) * n * 32;
            SB_CUDA_TRY(cudaMemcpyAsync(L.z_poly, d_zall + (size_t)(n_sets + (int)li) * n * 32, n * 32, cudaMemcpyDeviceToDevice, st2));
            SB_TRY(l2c_repl(L.z_poly, st2));
            SB_TRY(side_cosets((size_t)A + 1 + n_sets_all + 3 * li, L.z_poly));
        }
        std::vector<uint8_t> pts((size_t)n_z * 64);
        SB_TRY(msm_commit_batch(ctx, comm, pk->srs, 1, d_zall, n, (uint32_t)n_z, pts.data(), st));
        for (int s = 0; s < n_sets; s++)
            if (!tr.write_point(pts.data() + (size_t)s * 64)) { set_last_error("permutation product commitment is the identity"); return SB_ERR_ARG; }
        mark();  // [2] permutation + lookup products: ratios, batch inversion, scans, ONE batched commitment (iNTTs / coset NTTs on the side stream)
        for (size_t li = 0; li < lks.size(); li++)
            if (!tr.write_point(pts.data() + (size_t)(n_sets + (int)li) * 64)) { set_last_error("lookup product commitment is the identity"); return SB_ERR_ARG; }
    }

    mark();  // [3] lookup product
    // ---- vanishing argument: random polynomial from a child ChaCha20 stream (1-thread case of the fork)
    void *d_random;
    SB_TRY(scratch_get(ctx, "pf_random", n * 32, &d_random));
    {
        uint8_t seed[32];
        rng.fill_bytes(seed, 32);
        (void)rng.next_fr();
        if (random_early) {
            d_random = d_random_early;  // generated and committed with the advice columns (same seed: the clone made the same draws)
            memcpy(pt, random_commitment, 64);
        } else {
            ChaCha20Rng child;
            child.seed(seed);
            SB_TRY(chacha_fr_fill(ctx, child.key, 0, d_random, n, st));  // coefficient i = keystream block i
            SB_TRY(msm_commit(ctx, comm, pk->srs, 0, d_random, n, pt, st));
        }
        if (!tr.write_point(pt)) { set_last_error("random polynomial commitment is the identity"); return SB_ERR_ARG; }
    }
    const Fr yy = tr.squeeze();
    mark();  // [4] random polynomial (ChaCha20 on the device) + commitment

    // ---- evaluate_h: one fused program over the extended coset; first join the side stream (all coset forms)
    SB_TRY(main_after_side());
    mark();  // [5] coeff_to_extended of advice / instance / lookup polynomials (sharded: done per owned coset in stage 6)
    const int E_SIGMA = A + F + 1, E_PZ = E_SIGMA + P, E_L0 = E_PZ + n_sets, E_LLAST = E_L0 + 1, E_LACT = E_L0 + 2, E_X = E_L0 + 3, E_LK = E_L0 + 4;
    const size_t n_ecols = (size_t)E_LK + 3 * lks.size();
    std::vector<std::pair<int, int>> set_ranges;
    for (int s2 = 0; s2 < n_sets; s2++) set_ranges.push_back({psets[s2].first, psets[s2].count});
    // per-coset data of the quotient: slot (r, jl) of the exchange buffer holds coset slot jl * world + r; `per` slots per rank
    const uint32_t per = (n_cos + world - 1) / world;
    uint8_t *d_hcm;
    void *d_h;
    SB_TRY(scratch_get(ctx, "pf_h_cm", (size_t)per * world * n * 32, (void **)&d_hcm));
    SB_TRY(scratch_get(ctx, "pf_h", (size_t)n_cos * n * 32, &d_h));
    auto hcm_slot = [&](uint32_t s2) -> uint8_t * { return d_hcm + ((size_t)(s2 % world) * per + s2 / world) * n * 32; };
    {
        Program hp;
        if (ctx->tune.no_hprog_cache) {
            fr_t yd = to_dev(yy);
            hp = compile_terms(h_terms(cs, P, set_ranges, lks.size(), theta, beta, gamma), &yd);
        } else {
            hp = h_program_for(pk, set_ranges, lks.size(), theta, beta, gamma, yy);
        }
        ctx->last_h_program[0] = (uint32_t)(hp.code.size() / 3);
        ctx->last_h_program[1] = hp.n_mul;
        ctx->last_h_program[2] = hp.n_addsub;
        ctx->last_h_program[3] = hp.n_slots;
        ctx->last_h_rows = (uint64_t)co_per * n;
        if (!ctx->h_ev[0]) {  // created once per context: an early SB_TRY return below leaks nothing
            SB_CUDA_TRY(cudaEventCreate(&ctx->h_ev[0]));
            SB_CUDA_TRY(cudaEventCreate(&ctx->h_ev[1]));
        }
        cudaEvent_t e0 = ctx->h_ev[0], e1 = ctx->h_ev[1];
        SB_CUDA_TRY(cudaEventRecord(e0, st));
        // one launch of the fused program per owned coset: the key's columns are coset-major slabs, the per-proof polynomials' values were
        // computed on the side stream as the polynomials appeared (slab d_dyn); rotations move inside a coset
        std::vector<int> dyn_col;  // column-table index of per-proof polynomial q
        for (int c = 0; c < A; c++) dyn_col.push_back(c);
        dyn_col.push_back(A + F);
        for (int s2 = 0; s2 < n_sets; s2++) dyn_col.push_back(E_PZ + s2);
        for (size_t li = 0; li < lks.size(); li++) {
            dyn_col.push_back(E_LK + 3 * (int)li);
            dyn_col.push_back(E_LK + 3 * (int)li + 1);
            dyn_col.push_back(E_LK + 3 * (int)li + 2);
        }
        // the key's program as compiled straight-line code when NVRTC is there (same DAG, same schedule: same values); only its constants are per proof
        const ExprJit *jit = nullptr;
        void *d_jit_consts = nullptr;
        if (!ctx->tune.no_jit && !ctx->tune.no_hprog_cache) {
            sb_pk::HProgramCache &hc = pk->hcache;
            std::lock_guard<std::mutex> lk(hc.mu);
            if (!hc.jit_tried) {
                hc.jit_tried = true;
                hc.jit = expr_jit_compile(hc.prog, &hc.jit_why);
            }
            jit = hc.jit;
        }
        ctx->last_h_jit = jit != nullptr;
        if (jit) {
            static thread_local uint32_t ring = 0;
            uint8_t *slab;
            const size_t slice = 16 << 10;
            SB_REQUIRE(hp.consts.size() * 32 <= slice, "evaluate_h: constant table too large");
            SB_TRY(scratch_get(ctx, "hjit_consts", 4 * slice, (void **)&slab));
            d_jit_consts = slab + (size_t)(ring++ % 4) * slice;
            SB_TRY(h2d_staged(ctx, d_jit_consts, hp.consts.data(), hp.consts.size() * 32, st));
        }
        std::vector<const void *> ccols(n_ecols, nullptr);
        for (uint32_t jl = 0; jl < co_per; jl++) {
            const size_t off = (size_t)own[jl] * n * 32;
            for (int c = 0; c < F; c++) ccols[A + c] = (const uint8_t *)pk->fixed_cosets[c] + off;
            for (int j = 0; j < P; j++) ccols[E_SIGMA + j] = (const uint8_t *)pk->sigma_cosets[j] + off;
            ccols[E_L0] = (const uint8_t *)pk->l0 + off; ccols[E_LLAST] = (const uint8_t *)pk->l_last + off;
            ccols[E_LACT] = (const uint8_t *)pk->l_active + off; ccols[E_X] = (const uint8_t *)pk->x_coset + off;
            for (size_t q = 0; q < dyn_col.size(); q++) ccols[dyn_col[q]] = dyn_slot(jl, q);
            if (jit) SB_TRY(expr_jit_run(ctx, jit, d_jit_consts, ccols, pk->k, 0, hcm_slot(own[jl]), st));
            else SB_TRY(expr_eval(ctx, hp, ccols, pk->k, 0, hcm_slot(own[jl]), st));
        }
        SB_CUDA_TRY(cudaEventRecord(e1, st));
        SB_CUDA_TRY(sync_stream(ctx, st));
        cudaEventElapsedTime(&ctx->last_h_ms, e0, e1);
    }
    mark();  // [6] evaluate_h (fused program, one launch per owned coset)
    // ---- quotient: per coset, back to the coefficients d_s(X) = sum_q h_q(X) g_s^(qn) of the numerator restricted to g_s H (size-n inverse NTT, n^-1
    //      and the un-scaling g_s^-i fused); the cosets meet (sharded: one all-gather); a constant matrix gives h's pieces (and divides by t(X))
    if (co_per) {  // the owned slots are adjacent in the exchange buffer: one batched launch set
        NttFuse f;
        f.has_scale = true;
        f.scale = d->ifft_divisor;
        f.post_vec = (const uint8_t *)pk->coset_pows_inv + (size_t)rank * n * 32;
        f.batch = co_per;
        f.src_stride = n; f.dst_stride = n; f.post_stride = (uint64_t)world * n;
        SB_TRY(ntt_run_fused(ctx, hcm_slot(rank), hcm_slot(rank), (const uint8_t *)d->omega_inv.v, pk->k, &f, st));
    }
    if (world > 1) {
        SB_CUDA_TRY(sync_stream(ctx, st));
        if (comm->allgather_dev(comm->user, d_hcm, (size_t)per * n * 32, (void *)st) != 0) { set_last_error("sb_comm.allgather_dev failed"); return SB_ERR_ARG; }
    }
    {
        std::vector<const void *> slots(n_cos);
        for (uint32_t s2 = 0; s2 < n_cos; s2++) slots[s2] = hcm_slot(s2);
        SB_TRY(fr_coset_combine(ctx, slots, pk->combine, d_h, n, st));
    }
    const int n_pieces = cs.degree - 1;
    for (int i = 0; i < n_pieces; i++) (void)rng.next_fr();
    {
        std::vector<uint8_t> pts((size_t)n_pieces * 64);
        SB_TRY(msm_commit_batch(ctx, comm, pk->srs, 0, d_h, n, (uint32_t)n_pieces, pts.data(), st));
        for (int i = 0; i < n_pieces; i++)
            if (!tr.write_point(pts.data() + (size_t)i * 64)) { set_last_error("quotient piece commitment is the identity"); return SB_ERR_ARG; }
    }
    const Fr x = tr.squeeze();
    mark();  // [7] per-coset inverse NTTs, coset combine (includes the division by t(X)), quotient piece commitments
    const Fr xn = fpow(x, (uint64_t)n);
    // h(X) folded at x^n
    void *d_hfold;
    SB_TRY(scratch_get(ctx, "pf_hfold", n * 32, &d_hfold));
    {
        Fr p = hfr::ONE;
        for (int i = 0; i < n_pieces; i++) {
            SB_TRY(fr_axpy(ctx, d_hfold, (uint8_t *)d_h + (size_t)i * n * 32, to_dev(p), n, i == 0, st));
            p = hfr::mul(p, xn);
        }
    }

    // ---- evaluations (SURVEY A.5 order), one batched kernel
    const Fr x_next = rotate(x, omega, omega_inv, 1), x_prev = rotate(x, omega, omega_inv, -1), x_last = rotate(x, omega, omega_inv, -(bf + 1));
    std::vector<const void *> ev_polys;
    std::vector<fr_t> ev_pts;
    auto want = [&](const void *p, const Fr &at) { ev_polys.push_back(p); ev_pts.push_back(to_dev(at)); return (int)ev_polys.size() - 1; };
    std::vector<int> i_adv, i_fix, i_sig;
    for (auto &q : cs.advice_q) i_adv.push_back(want(adv_poly[q.first], rotate(x, omega, omega_inv, q.second)));
    for (auto &q : cs.fixed_q) i_fix.push_back(want(pk->fixed_polys[q.first], rotate(x, omega, omega_inv, q.second)));
    const int i_rand = want(d_random, x);
    for (int j = 0; j < P; j++) i_sig.push_back(want(pk->sigma_polys[j], x));
    std::vector<std::vector<int>> i_perm(n_sets);
    for (int s = 0; s < n_sets; s++) {
        i_perm[s].push_back(want(psets[s].z_poly, x));
        i_perm[s].push_back(want(psets[s].z_poly, x_next));
        if (s != n_sets - 1) i_perm[s].push_back(want(psets[s].z_poly, x_last));
    }
    std::vector<std::vector<int>> i_lk(lks.size());
    for (size_t li = 0; li < lks.size(); li++) {
        i_lk[li] = {want(lks[li].z_poly, x), want(lks[li].z_poly, x_next), want(lks[li].in_poly, x), want(lks[li].in_poly, x_prev), want(lks[li].tab_poly, x)};
    }
    const int i_h = want(d_hfold, x);
    std::vector<fr_t> evd;
    if (world > 1 && !ctx->tune.no_shplonk_shard) {
        // sharded proof: the (polynomial, point) pairs are dealt to the ranks in contiguous runs; the 32-byte values meet on the host
        const size_t m_ev = ev_polys.size(), per_ev = (m_ev + world - 1) / world;
        const size_t lo = std::min(m_ev, per_ev * rank), hi = std::min(m_ev, lo + per_ev);
        std::vector<fr_t> mine_v;
        if (hi > lo) {
            std::vector<const void *> sub_p(ev_polys.begin() + lo, ev_polys.begin() + hi);
            std::vector<fr_t> sub_x(ev_pts.begin() + lo, ev_pts.begin() + hi);
            SB_TRY(fr_eval_polys(ctx, sub_p, sub_x, n, mine_v, st));
        }
        std::vector<fr_t> mine(per_ev, fr_t::zero()), all(per_ev * world);
        for (size_t i = 0; i < mine_v.size(); i++) mine[i] = mine_v[i];
        if (comm->allgather_host(comm->user, mine.data(), all.data(), per_ev * 32) != 0) { set_last_error("sb_comm.allgather_host failed"); return SB_ERR_ARG; }
        evd.resize(m_ev);
        for (size_t i = 0; i < m_ev; i++) evd[i] = all[(i / per_ev) * per_ev + (i % per_ev)];
    } else {
        SB_TRY(fr_eval_polys(ctx, ev_polys, ev_pts, n, evd, st));
    }
    std::vector<Fr> ev(evd.size());
    for (size_t i = 0; i < evd.size(); i++) ev[i] = to_host(evd[i]);
    for (int i : i_adv) tr.write_scalar(ev[i]);
    for (int i : i_fix) tr.write_scalar(ev[i]);
    tr.write_scalar(ev[i_rand]);
    for (int i : i_sig) tr.write_scalar(ev[i]);
    for (int s = 0; s < n_sets; s++)
        for (int i : i_perm[s]) tr.write_scalar(ev[i]);
    for (size_t li = 0; li < lks.size(); li++)
        for (int i : i_lk[li]) tr.write_scalar(ev[i]);

    mark();  // [8] h fold + the batched evaluations
    // ---- multi-open queries in halo2's order
    std::vector<Query> q;
    int next_id = 0;
    std::map<const void *, int> ids;
    auto pid = [&](const void *p) { auto it = ids.find(p); if (it != ids.end()) return it->second; return ids[p] = next_id++; };
    for (size_t i = 0; i < cs.advice_q.size(); i++) {
        const void *p = adv_poly[cs.advice_q[i].first];
        q.push_back({pid(p), rotate(x, omega, omega_inv, cs.advice_q[i].second), p, ev[i_adv[i]], adv[cs.advice_q[i].first]});
    }
    for (int s = 0; s < n_sets; s++) {
        q.push_back({pid(psets[s].z_poly), x, psets[s].z_poly, ev[i_perm[s][0]], psets[s].z_vals});
        q.push_back({pid(psets[s].z_poly), x_next, psets[s].z_poly, ev[i_perm[s][1]], psets[s].z_vals});
    }
    for (int s = n_sets - 2; s >= 0; s--) q.push_back({pid(psets[s].z_poly), x_last, psets[s].z_poly, ev[i_perm[s][2]], psets[s].z_vals});
    for (size_t li = 0; li < lks.size(); li++) {
        const LookupState &L = lks[li];
        q.push_back({pid(L.z_poly), x, L.z_poly, ev[i_lk[li][0]], L.z_vals});
        q.push_back({pid(L.in_poly), x, L.in_poly, ev[i_lk[li][2]], L.p_in});
        q.push_back({pid(L.tab_poly), x, L.tab_poly, ev[i_lk[li][4]], L.p_tab});
        q.push_back({pid(L.in_poly), x_prev, L.in_poly, ev[i_lk[li][3]], L.p_in});
        q.push_back({pid(L.z_poly), x_next, L.z_poly, ev[i_lk[li][1]], L.z_vals});
    }
    for (size_t i = 0; i < cs.fixed_q.size(); i++) {
        const void *p = pk->fixed_polys[cs.fixed_q[i].first];
        q.push_back({pid(p), rotate(x, omega, omega_inv, cs.fixed_q[i].second), p, ev[i_fix[i]], pk->fixed_values[cs.fixed_q[i].first]});
    }
    for (int j = 0; j < P; j++) q.push_back({pid(pk->sigma_polys[j]), x, pk->sigma_polys[j], ev[i_sig[j]], pk->sigma_values[j]});
    q.push_back({pid(d_hfold), x, d_hfold, ev[i_h], nullptr});
    q.push_back({pid(d_random), x, d_random, ev[i_rand], nullptr});
    int32_t rc_sh = shplonk(ctx, pk, comm, tr, q, st);
    mark();  // [9] SHPLONK
    return rc_sh;
}

}  // namespace

extern "C" {

static void pk_free(sb_pk *pk) {
    expr_jit_free(pk->hcache.jit);
    for (void *p : pk->owned) cudaFree(p);
    if (pk->dom) sb_domain_destroy(pk->dom);
    if (pk->srs) sb_srs_destroy(const_cast<sb_srs *>(pk->srs));  // drops the key's reference
    delete pk;
}

// every column / rotation an expression tree touches must exist (a bad index would be an out-of-bounds host read later)
static void validate_expr(const json::Value &e, const ConstraintSystem &cs) {
    const std::string &k = e[0].as_str();
    if (k == "const") return;
    if (k == "advice" || k == "fixed" || k == "instance") {
        const int c = (int)e[1].as_int(), r = (int)e[2].as_int();
        const int lim = k == "advice" ? cs.A : k == "fixed" ? cs.F : cs.I;
        if (c < 0 || c >= lim) throw std::runtime_error("expression: " + k + " column index out of range");
        if (r < -64 || r > 64) throw std::runtime_error("expression: rotation out of range");
        return;
    }
    if (k == "neg") { validate_expr(e[1], cs); return; }
    if (k == "add" || k == "sub" || k == "mul") { validate_expr(e[1], cs); validate_expr(e[2], cs); return; }
    throw std::runtime_error("unknown expression node " + k);
}
static void validate_cs(const ConstraintSystem &cs, size_t n) {
    if (cs.degree < 3) throw std::runtime_error("constraint system: degree must be >= 3 (halo2 clamps cs.degree() to the permutation argument's 3)");
    if (cs.A < 1 || cs.A > 64 || cs.F < 0 || cs.F > 256) throw std::runtime_error("constraint system: column counts out of range");
    if (cs.blinding < 0 || (size_t)cs.blinding + 1 >= n) throw std::runtime_error("constraint system: blinding_factors + 1 must be < 2^k");
    if (cs.n_instances < 0 || (size_t)cs.n_instances > n - (size_t)cs.blinding - 1) throw std::runtime_error("constraint system: num_instances exceeds the usable rows");
    for (auto &q : cs.advice_q)
        if (q.first < 0 || q.first >= cs.A || q.second < -64 || q.second > 64) throw std::runtime_error("advice query out of range");
    for (auto &q : cs.fixed_q)
        if (q.first < 0 || q.first >= cs.F || q.second < -64 || q.second > 64) throw std::runtime_error("fixed query out of range");
    for (auto &pc : cs.perm_cols) {
        const int lim = pc.first == "advice" ? cs.A : pc.first == "fixed" ? cs.F : pc.first == "instance" ? cs.I : -1;
        if (lim < 0 || pc.second < 0 || pc.second >= lim) throw std::runtime_error("permutation column out of range");
    }
    for (auto &g : cs.gates) validate_expr(*g, cs);
    for (auto &l : cs.lookups) {
        for (auto &e : l.input) validate_expr(*e, cs);
        for (auto &e : l.table) validate_expr(*e, cs);
    }
}

static int32_t pk_create_common(sb_ctx *ctx, const sb_srs *srs, const char *cs_json, uint32_t k, const uint8_t *fixed_values, const uint8_t *sigma_values,
                                const SparseAssignment *sparse, const uint8_t transcript_repr[32], sb_pk **out_pk) {
    SB_REQUIRE(srs->k == k, "sb_pk_create: SRS size does not match k (downsize first)");
    SB_REQUIRE(k >= 7 && k <= 25, "sb_pk_create: k must be in [7, 25]");
    CtxGuard g(ctx);
    sb_pk *pk = new sb_pk();
    try {
        pk->cs = parse_cs(cs_json);
        validate_cs(pk->cs, (size_t)1 << k);
    } catch (const std::exception &e) {
        set_last_error("sb_pk_create: %s", e.what());
        pk_free(pk);
        return SB_ERR_ARG;
    }
    pk->srs = srs;
    const_cast<sb_srs *>(srs)->refs.fetch_add(1, std::memory_order_relaxed);
    pk->k = k;
    pk->n = (size_t)1 << k;
    pk->P = (int)pk->cs.perm_cols.size();
    memcpy(pk->transcript_repr.v, transcript_repr, 32);
    int32_t rc = sb_domain_create(ctx, (uint32_t)pk->cs.degree, k, &pk->dom);
    if (rc != SB_OK) { pk_free(pk); return rc; }
    pk->ext_k = pk->dom->ext_k;
    pk->ext_n = (size_t)1 << pk->ext_k;
    // fixed-base window tables for both bases of this key's SRS (the handle is shared and logically const: tables change speed, not results)
    if (!ctx->tune.no_tables) {
        rc = srs_precompute_impl(ctx, const_cast<sb_srs *>(srs), 3, 0);
        if (rc != SB_OK) { pk_free(pk); return rc; }
    }
    try {
        rc = pk_build(ctx, pk, fixed_values, sigma_values, sparse, ctx->stream);
    } catch (const std::exception &e) {
        set_last_error("sb_pk_create: %s", e.what());
        rc = SB_ERR_ARG;
    }
    if (rc != SB_OK) {
        cudaStreamSynchronize(ctx->stream);
        pk_free(pk);
        return rc;
    }
    *out_pk = pk;
    return SB_OK;
}

int32_t sb_pk_create(sb_ctx *ctx, const sb_srs *srs, const char *cs_json, uint32_t k, const uint8_t *fixed_values, const uint8_t *sigma_values,
                     const uint8_t transcript_repr[32], sb_pk **out_pk) {
    if (!ctx || !srs || !cs_json || !fixed_values || !sigma_values || !transcript_repr || !out_pk) return SB_ERR_ARG;
    return pk_create_common(ctx, srs, cs_json, k, fixed_values, sigma_values, nullptr, transcript_repr, out_pk);
}

int32_t sb_pk_create_sparse(sb_ctx *ctx, const sb_srs *srs, const char *cs_json, uint32_t k, const uint32_t *fixed_cells, const uint8_t *fixed_cell_values,
                            size_t n_fixed, const uint32_t *perm_cells, size_t n_perm, const uint8_t transcript_repr[32], sb_pk **out_pk) {
    if (!ctx || !srs || !cs_json || !transcript_repr || !out_pk || (n_fixed && (!fixed_cells || !fixed_cell_values)) || (n_perm && !perm_cells)) return SB_ERR_ARG;
    SparseAssignment sp;
    sp.fixed_cells = fixed_cells; sp.fixed_values = fixed_cell_values; sp.n_fixed = n_fixed;
    sp.perm_cells = perm_cells; sp.n_perm = n_perm;
    return pk_create_common(ctx, srs, cs_json, k, nullptr, nullptr, &sp, transcript_repr, out_pk);
}

int32_t sb_pk_destroy(sb_pk *pk) {
    if (!pk) return SB_OK;
    pk_free(pk);
    return SB_OK;
}

int32_t sb_pk_commitments(const sb_pk *pk, uint8_t *fixed_comms, uint8_t *sigma_comms) {
    if (!pk || !fixed_comms || !sigma_comms) return SB_ERR_ARG;
    memcpy(fixed_comms, pk->fixed_comms.data(), pk->fixed_comms.size());
    memcpy(sigma_comms, pk->sigma_comms.data(), pk->sigma_comms.size());
    return SB_OK;
}

static int32_t create_proof_entry(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const Witness &wit,
                                  const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len) {
    if (!ctx || !pk || !rng_seed || !proof_out || !proof_len || (n_instances && !instances)) return SB_ERR_ARG;
    if (wit.sparse ? (wit.n_cells && (!wit.cells || !wit.values)) : (!wit.host && !wit.dev)) return SB_ERR_ARG;
    SB_REQUIRE(transcript_kind == 0 || transcript_kind == 1, "transcript_kind must be 0 (Blake2b) or 1 (Keccak256/EVM)");
    if (comm) {
        const uint32_t n_cosets = 1u << (pk->ext_k - pk->k);
        SB_REQUIRE(comm->world >= 1 && comm->rank >= 0 && comm->rank < comm->world, "sb_comm: rank / world out of range");
        SB_REQUIRE((uint32_t)comm->world <= n_cosets, "sb_comm: world must not exceed the 2^(extended_k - k) cosets of the extended domain");
        SB_REQUIRE(comm->world == 1 || (comm->allgather_host && comm->allgather_dev), "sb_comm: callbacks missing");
    }
    CtxGuard g(ctx);
    ChaCha20Rng rng;
    rng.seed(rng_seed);
    KeccakTranscript kt;
    Blake2bTranscript bt;
    Transcript &tr = transcript_kind == 1 ? (Transcript &)kt : (Transcript &)bt;
    int32_t rc;
    try {
        rc = create_proof_impl(ctx, pk, comm, instances, n_instances, wit, rng, tr, ctx->stream);
    } catch (const std::exception &e) {
        set_last_error("sb_create_proof: %s", e.what());
        return SB_ERR_ARG;
    }
    if (rc != SB_OK) return rc;
    *proof_len = tr.proof.size();
    SB_REQUIRE(tr.proof.size() <= proof_cap, "sb_create_proof: output buffer too small");
    memcpy(proof_out, tr.proof.data(), tr.proof.size());
    return SB_OK;
}
int32_t sb_create_proof(sb_ctx *ctx, const sb_pk *pk, const uint8_t *instances, size_t n_instances, const uint8_t *advice, const uint8_t rng_seed[32],
                        int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len) {
    if (!advice) return SB_ERR_ARG;
    Witness w;
    w.host = advice;
    return create_proof_entry(ctx, pk, nullptr, instances, n_instances, w, rng_seed, transcript_kind, proof_out, proof_cap, proof_len);
}
int32_t sb_create_proof_dev(sb_ctx *ctx, const sb_pk *pk, const uint8_t *instances, size_t n_instances, const void *d_advice, const uint8_t rng_seed[32],
                            int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len) {
    if (!d_advice) return SB_ERR_ARG;
    Witness w;
    w.dev = d_advice;
    return create_proof_entry(ctx, pk, nullptr, instances, n_instances, w, rng_seed, transcript_kind, proof_out, proof_cap, proof_len);
}
int32_t sb_create_proof_sparse(sb_ctx *ctx, const sb_pk *pk, const uint8_t *instances, size_t n_instances, const uint32_t *advice_cells, const uint8_t *advice_cell_values,
                               size_t n_cells, const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len) {
    Witness w;
    w.sparse = true;
    w.cells = advice_cells; w.values = advice_cell_values; w.n_cells = n_cells;
    return create_proof_entry(ctx, pk, nullptr, instances, n_instances, w, rng_seed, transcript_kind, proof_out, proof_cap, proof_len);
}
int32_t sb_create_proof_sharded(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const uint8_t *advice,
                                const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len) {
    if (!comm || !advice) return SB_ERR_ARG;
    Witness w;
    w.host = advice;
    return create_proof_entry(ctx, pk, comm, instances, n_instances, w, rng_seed, transcript_kind, proof_out, proof_cap, proof_len);
}
int32_t sb_create_proof_sharded_dev(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const void *d_advice,
                                    const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out, size_t proof_cap, size_t *proof_len) {
    if (!comm || !d_advice) return SB_ERR_ARG;
    Witness w;
    w.dev = d_advice;
    return create_proof_entry(ctx, pk, comm, instances, n_instances, w, rng_seed, transcript_kind, proof_out, proof_cap, proof_len);
}
int32_t sb_create_proof_sharded_sparse(sb_ctx *ctx, const sb_pk *pk, const sb_comm *comm, const uint8_t *instances, size_t n_instances, const uint32_t *advice_cells,
                                       const uint8_t *advice_cell_values, size_t n_cells, const uint8_t rng_seed[32], int32_t transcript_kind, uint8_t *proof_out,
                                       size_t proof_cap, size_t *proof_len) {
    if (!comm) return SB_ERR_ARG;
    Witness w;
    w.sparse = true;
    w.cells = advice_cells; w.values = advice_cell_values; w.n_cells = n_cells;
    return create_proof_entry(ctx, pk, comm, instances, n_instances, w, rng_seed, transcript_kind, proof_out, proof_cap, proof_len);
}

// ---- building blocks with host buffers (sb_batch_invert, sb_grand_product, sb_sort_fr, sb_eval_poly, sb_kate_div)
int32_t sb_fr_batch_invert(sb_ctx *ctx, uint8_t *a, size_t n) {
    if (!ctx || (n && !a)) return SB_ERR_ARG;
    CtxGuard g(ctx);
    void *d;
    SB_TRY(scratch_get(ctx, "bb_a", n * 32, &d));
    SB_CUDA_TRY(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(fr_batch_invert(ctx, d, n, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, d, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_fr_running_product(sb_ctx *ctx, const uint8_t *a, size_t n_a, const uint8_t init[32], uint8_t *z, size_t n_z) {
    if (!ctx || !init || (n_a && !a) || (n_z && !z)) return SB_ERR_ARG;
    CtxGuard g(ctx);
    void *da, *dz;
    SB_TRY(scratch_get(ctx, "bb_a", (n_a + 1) * 32, &da));
    SB_TRY(scratch_get(ctx, "bb_b", (n_z + 1) * 32, &dz));
    SB_CUDA_TRY(cudaMemcpyAsync(da, a, n_a * 32, cudaMemcpyHostToDevice, ctx->stream));
    fr_t i0;
    memcpy(i0.v, init, 32);
    SB_TRY(fr_running_product(ctx, da, n_a, i0, dz, n_z, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(z, dz, n_z * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
int32_t sb_fr_eval_polynomial(sb_ctx *ctx, const uint8_t *coeffs, size_t n, const uint8_t *points, size_t n_points, uint8_t *out) {
    if (!ctx || !coeffs || !points || !out) return SB_ERR_ARG;
    CtxGuard g(ctx);
    void *d;
    SB_TRY(scratch_get(ctx, "bb_a", n * 32, &d));
    SB_CUDA_TRY(cudaMemcpyAsync(d, coeffs, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<const void *> polys(n_points, d);
    std::vector<fr_t> xs(n_points), res;
    memcpy(xs.data(), points, n_points * 32);
    SB_TRY(fr_eval_polys(ctx, polys, xs, n, res, ctx->stream));
    memcpy(out, res.data(), n_points * 32);
    return SB_OK;
}
int32_t sb_fr_sort(sb_ctx *ctx, uint8_t *a, size_t n) {
    if (!ctx || (n && !a)) return SB_ERR_ARG;
    CtxGuard g(ctx);
    size_t N = 1;
    while (N < n) N <<= 1;
    void *d, *d2;
    SB_TRY(scratch_get(ctx, "bb_a", N * 32, &d));
    SB_TRY(scratch_get(ctx, "bb_b", N * 32, &d2));
    SB_CUDA_TRY(cudaMemcpyAsync(d, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    // canonical form for the comparison (halo2curves `Ord`), back to Montgomery afterwards: x * 1 and x * R^2
    fr_t one_c = fr_t::zero();
    one_c.v[0] = 1;
    SB_TRY(fr_scale(ctx, d, n, one_c, ctx->stream));
    SB_TRY(sort_u256(ctx, d, n, N, ctx->stream));
    SB_TRY(fr_scale(ctx, d, n, fr_t::r2(), ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(a, d, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    (void)d2;
    return SB_OK;
}
int32_t sb_lookup_permute(sb_ctx *ctx, const uint8_t *input, const uint8_t *table, size_t n, size_t usable, uint8_t *permuted_input, uint8_t *permuted_table) {
    if (!ctx || !input || !table || !permuted_input || !permuted_table || usable > n) return SB_ERR_ARG;
    CtxGuard g(ctx);
    void *di, *dt, *dpi, *dpt;
    SB_TRY(scratch_get(ctx, "bb_a", n * 32, &di));
    SB_TRY(scratch_get(ctx, "bb_b", n * 32, &dt));
    SB_TRY(scratch_get(ctx, "bb_c", n * 32, &dpi));
    SB_TRY(scratch_get(ctx, "bb_d", n * 32, &dpt));
    SB_CUDA_TRY(cudaMemcpyAsync(di, input, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(dt, table, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SB_TRY(lookup_permute(ctx, di, dt, n, usable, dpi, dpt, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(permuted_input, dpi, usable * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(permuted_table, dpt, usable * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}
// halo2 `kate_division(a, b)`: q = (a - a(b)) / (X - b); n = 2^log_n coefficients in, n - 1 out (out has n slots, top = 0)
int32_t sb_kate_division(sb_ctx *ctx, const uint8_t *a, uint32_t log_n, const uint8_t b[32], uint8_t *q) {
    if (!ctx || !a || !b || !q) return SB_ERR_ARG;
    SB_REQUIRE(log_n >= 7 && log_n <= 28, "sb_kate_division: log_n must be in [7, 28]");
    CtxGuard g(ctx);
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)1 << log_n;
    void *dp, *dg, *dx, *dgi;
    SB_TRY(scratch_get(ctx, "bb_a", n * 32, &dp));
    SB_TRY(scratch_get(ctx, "bb_b", n * 32, &dg));
    SB_TRY(scratch_get(ctx, "bb_c", n * 32, &dx));
    SB_TRY(scratch_get(ctx, "bb_d", n * 32, &dgi));
    SB_CUDA_TRY(cudaMemcpyAsync(dp, a, n * 32, cudaMemcpyHostToDevice, st));
    sb_domain *dom = nullptr;
    SB_TRY(sb_domain_create(ctx, 2, log_n, &dom));
    fr_t bd;
    memcpy(bd.v, b, 32);
    std::vector<const void *> polys(1, dp);
    std::vector<fr_t> xs(1, bd), res;
    SB_TRY(fr_eval_polys(ctx, polys, xs, n, res, st));
    SB_TRY(fr_sub_head(ctx, dp, res.data(), 1, st));
    SB_TRY(fr_gen_powers(ctx, dg, to_dev(DIV_G), n, st));
    SB_TRY(fr_gen_powers(ctx, dx, dom->omega, n, st));
    SB_TRY(fr_scale(ctx, dx, n, to_dev(DIV_G), st));
    SB_TRY(fr_gen_powers(ctx, dgi, to_dev(hfr::inv(DIV_G)), n, st));
    SB_TRY(fr_scale(ctx, dgi, n, dom->ifft_divisor, st));
    int32_t rc = poly_div_by_roots_k(ctx, log_n, dom->omega, dom->omega_inv, dg, dx, dgi, dp, {to_host(bd)}, st);
    sb_domain_destroy(dom);
    if (rc != SB_OK) return rc;
    SB_CUDA_TRY(cudaMemcpyAsync(q, dp, n * 32, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(sync_stream(ctx, st));
    return SB_OK;
}

// ---- halo2 `Evaluator::evaluate_h` as a standalone entry: the quotient numerator of a constraint system over caller-supplied columns ----
// Column order (HLayout): advice (A) | fixed (F) | instance (1) | sigma (P) | permutation Z (ceil(P / (degree - 2))) | l_0, l_last, l_active, X |
// per lookup: Z, A', S'.  Every column holds 2^log_rows values of its polynomial on ONE common domain; a row rotation r moves by r << rot_scale_log
// entries cyclically (halo2's extended domain in natural order: rot_scale_log = extended_k - k; one coset of the size-n subgroup: 0).
static int32_t evaluate_h_common(sb_ctx *ctx, const char *cs_json, const std::vector<const void *> &cols, uint32_t log_rows, uint32_t rot_scale_log, const uint8_t theta[32],
                                 const uint8_t beta[32], const uint8_t gamma[32], const uint8_t y[32], void *d_out, cudaStream_t st) {
    ConstraintSystem cs;
    try {
        cs = parse_cs(cs_json);
        validate_cs(cs, (size_t)1 << 28);
    } catch (const std::exception &e) {
        set_last_error("sb_evaluate_h: %s", e.what());
        return SB_ERR_ARG;
    }
    const int P = (int)cs.perm_cols.size(), chunk = cs.degree - 2;
    std::vector<std::pair<int, int>> sets;
    for (int f = 0; f < P; f += chunk) sets.push_back({f, std::min(chunk, P - f)});
    const HLayout lay(cs, P, (int)sets.size(), cs.lookups.size());
    SB_REQUIRE((int)cols.size() == lay.n_cols, "sb_evaluate_h: column count does not match the constraint system (advice | fixed | instance | sigma | Z | l0 l_last l_active X | lookups)");
    Fr th, be, ga, yy;
    memcpy(th.v, theta, 32); memcpy(be.v, beta, 32); memcpy(ga.v, gamma, 32); memcpy(yy.v, y, 32);
    try {
        fr_t yd = to_dev(yy);
        const Program hp = compile_terms(h_terms(cs, P, sets, cs.lookups.size(), th, be, ga), &yd);
        ctx->last_h_program[0] = (uint32_t)(hp.code.size() / 3); ctx->last_h_program[1] = hp.n_mul; ctx->last_h_program[2] = hp.n_addsub; ctx->last_h_program[3] = hp.n_slots;
        return expr_eval(ctx, hp, cols, log_rows, rot_scale_log, d_out, st);
    } catch (const std::exception &e) {
        set_last_error("sb_evaluate_h: %s", e.what());
        return SB_ERR_ARG;
    }
}
int32_t sb_evaluate_h_dev(sb_ctx *ctx, const char *cs_json, const void *const *d_columns, size_t n_columns, uint32_t log_rows, uint32_t rot_scale_log, const uint8_t theta[32],
                          const uint8_t beta[32], const uint8_t gamma[32], const uint8_t y[32], void *d_out, void *stream) {
    if (!ctx || !cs_json || !d_columns || !theta || !beta || !gamma || !y || !d_out) return SB_ERR_ARG;
    SB_REQUIRE(log_rows >= 1 && log_rows <= 28 && rot_scale_log < log_rows, "sb_evaluate_h: log_rows must be in [1, 28] and rot_scale_log below it");
    for (size_t i = 0; i < n_columns; i++) SB_REQUIRE(d_columns[i] != nullptr, "sb_evaluate_h: null column");
    CtxGuard g(ctx);
    return evaluate_h_common(ctx, cs_json, std::vector<const void *>(d_columns, d_columns + n_columns), log_rows, rot_scale_log, theta, beta, gamma, y, d_out, pick_stream(ctx, stream));
}
int32_t sb_evaluate_h(sb_ctx *ctx, const char *cs_json, const uint8_t *const *columns, size_t n_columns, uint32_t log_rows, uint32_t rot_scale_log, const uint8_t theta[32],
                      const uint8_t beta[32], const uint8_t gamma[32], const uint8_t y[32], uint8_t *out) {
    if (!ctx || !cs_json || !columns || !theta || !beta || !gamma || !y || !out) return SB_ERR_ARG;
    SB_REQUIRE(log_rows >= 1 && log_rows <= 28 && rot_scale_log < log_rows, "sb_evaluate_h: log_rows must be in [1, 28] and rot_scale_log below it");
    CtxGuard g(ctx);
    const size_t bytes = (size_t)32 << log_rows;
    uint8_t *slab;
    void *d_out;
    SB_TRY(scratch_get(ctx, "eh_cols", bytes * n_columns, (void **)&slab));
    SB_TRY(scratch_get(ctx, "eh_out", bytes, &d_out));
    std::vector<const void *> cols(n_columns);
    for (size_t i = 0; i < n_columns; i++) {
        SB_REQUIRE(columns[i] != nullptr, "sb_evaluate_h: null column");
        cols[i] = slab + i * bytes;
        SB_CUDA_TRY(cudaMemcpyAsync(slab + i * bytes, columns[i], bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    SB_TRY(evaluate_h_common(ctx, cs_json, cols, log_rows, rot_scale_log, theta, beta, gamma, y, d_out, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// ---- ParamsKZG::commit / commit_lagrange for m polynomials at once (one launch set over the fixed-base tables) ----
int32_t sb_msm_g1_batch(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const uint8_t *scalars, size_t n, size_t m, uint8_t *out_affine) {
    if (!ctx || !srs || !out_affine || (n && m && !scalars)) return SB_ERR_ARG;
    SB_REQUIRE(basis == SB_BASIS_MONOMIAL || basis == SB_BASIS_LAGRANGE, "basis must be 0 or 1");
    SB_REQUIRE(n <= ((size_t)1 << srs->k), "msm: more scalars than SRS bases");
    SB_REQUIRE(m >= 1 && m <= 64, "sb_msm_g1_batch: 1..64 vectors");
    CtxGuard g(ctx);
    void *ds;
    SB_TRY(scratch_get(ctx, "mx_scalars", n * m * 32 + 32, &ds));
    SB_CUDA_TRY(cudaMemcpyAsync(ds, scalars, n * m * 32, cudaMemcpyHostToDevice, ctx->stream));
    // launch sets of at most 8 vectors (the bucket arena of a set grows with its width)
    for (size_t j0 = 0; j0 < m; j0 += 8) {
        const uint32_t cnt = (uint32_t)std::min<size_t>(8, m - j0);
        SB_TRY(srs_msm_batch(ctx, srs, basis, (const uint8_t *)ds + j0 * n * 32, n, cnt, out_affine + j0 * 64, ctx->stream));
    }
    return SB_OK;
}
int32_t sb_msm_g1_batch_dev(sb_ctx *ctx, const sb_srs *srs, int32_t basis, const void *d_scalars, size_t n, size_t m, uint8_t *out_affine, void *stream) {
    if (!ctx || !srs || !out_affine || (n && m && !d_scalars)) return SB_ERR_ARG;
    SB_REQUIRE(basis == SB_BASIS_MONOMIAL || basis == SB_BASIS_LAGRANGE, "basis must be 0 or 1");
    SB_REQUIRE(n <= ((size_t)1 << srs->k), "msm: more scalars than SRS bases");
    SB_REQUIRE(m >= 1 && m <= 64, "sb_msm_g1_batch: 1..64 vectors");
    CtxGuard g(ctx);
    for (size_t j0 = 0; j0 < m; j0 += 8) {
        const uint32_t cnt = (uint32_t)std::min<size_t>(8, m - j0);
        SB_TRY(srs_msm_batch(ctx, srs, basis, (const uint8_t *)d_scalars + j0 * n * 32, n, cnt, out_affine + j0 * 64, pick_stream(ctx, stream)));
    }
    return SB_OK;
}

// ---- the grand-product column of the permutation / lookup arguments (halo2 permutation::Argument::commit inner loop):
//      z[0] = init, z[i + 1] = z[i] * num[i] / den[i]; n_z values are written (n_z <= n + 1).  One batch inversion, one product pass, one scan. ----
int32_t sb_grand_product(sb_ctx *ctx, const uint8_t *numerators, const uint8_t *denominators, size_t n, const uint8_t init[32], uint8_t *z, size_t n_z) {
    if (!ctx || !init || (n && (!numerators || !denominators)) || (n_z && !z)) return SB_ERR_ARG;
    SB_REQUIRE(n_z <= n + 1, "sb_grand_product: n_z must be <= n + 1");
    CtxGuard g(ctx);
    cudaStream_t st = ctx->stream;
    void *dn, *dd, *dz;
    SB_TRY(scratch_get(ctx, "bb_a", (n + 1) * 32, &dn));
    SB_TRY(scratch_get(ctx, "bb_b", (n + 1) * 32, &dd));
    SB_TRY(scratch_get(ctx, "bb_c", (n_z + 1) * 32, &dz));
    SB_CUDA_TRY(cudaMemcpyAsync(dn, numerators, n * 32, cudaMemcpyHostToDevice, st));
    SB_CUDA_TRY(cudaMemcpyAsync(dd, denominators, n * 32, cudaMemcpyHostToDevice, st));
    SB_TRY(fr_batch_invert(ctx, dd, n, st));
    SB_TRY(fp_vec_op(ctx, 0, 0, dd, dn, dd, n, st));
    fr_t i0;
    memcpy(i0.v, init, 32);
    SB_TRY(fr_running_product(ctx, dd, n, i0, dz, n_z, st));
    SB_CUDA_TRY(cudaMemcpyAsync(z, dz, n_z * 32, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(sync_stream(ctx, st));
    return SB_OK;
}

// `ParamsKZG::setup(k, rng)` (utils.rs:70) with an explicit secret: g[i] = [tau^i] G, g_lagrange[i] = [L_i(tau)] G, all on the device.
// UNSAFE by construction (the caller knows tau): test / benchmark SRS only.
int32_t sb_srs_setup_unsafe(sb_ctx *ctx, uint32_t k, const uint8_t tau_mont[32], sb_srs **out_srs) {
    if (!ctx || !tau_mont || !out_srs) return SB_ERR_ARG;
    SB_REQUIRE(k >= 7 && k <= 26, "sb_srs_setup_unsafe: k must be in [7, 26]");
    CtxGuard g(ctx);
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)1 << k;
    Fr tau;
    memcpy(tau.v, tau_mont, 32);
    sb_domain *dom = nullptr;
    SB_TRY(sb_domain_create(ctx, 2, k, &dom));
    sb_srs *srs = new sb_srs();
    srs->k = k;
    if (cudaMalloc(&srs->d_g, n * 64) != cudaSuccess || cudaMalloc(&srs->d_g_lagrange, n * 64) != cudaSuccess) {
        set_last_error("sb_srs_setup_unsafe: cudaMalloc(2 x %zu) failed", n * 64);
        if (srs->d_g) cudaFree(srs->d_g);
        delete srs;
        sb_domain_destroy(dom);
        return SB_ERR_ALLOC;
    }
    void *d_s, *d_w;
    int32_t rc = scratch_get(ctx, "setup_s", n * 32, &d_s);
    if (rc == SB_OK) rc = scratch_get(ctx, "setup_w", n * 32, &d_w);
    if (rc == SB_OK) rc = fr_gen_powers(ctx, d_s, to_dev(tau), n, st);
    if (rc == SB_OK) rc = g1_fixed_base_mul(ctx, d_s, n, srs->d_g, st);
    // L_i(tau) = omega^i * (tau^n - 1) / (n * (tau - omega^i))
    if (rc == SB_OK) rc = fr_gen_powers(ctx, d_w, dom->omega, n, st);
    if (rc == SB_OK) rc = expr_eval(ctx, compile_terms({e_sub(ec(tau), e_col(0, 0))}, nullptr), {d_w}, k, 0, d_s, st);
    if (rc == SB_OK) rc = fr_batch_invert(ctx, d_s, n, st);
    if (rc == SB_OK) {
        const Fr c = hfr::mul(hfr::sub(hfr::pow_u64(tau, (uint64_t)n), hfr::ONE), to_host(dom->ifft_divisor));
        rc = expr_eval(ctx, compile_terms({e_mul(e_mul(ec(c), e_col(0, 0)), e_col(1, 0))}, nullptr), {d_w, d_s}, k, 0, d_s, st);
    }
    if (rc == SB_OK) rc = g1_fixed_base_mul(ctx, d_s, n, srs->d_g_lagrange, st);
    if (rc == SB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = SB_ERR_CUDA;
    sb_domain_destroy(dom);
    if (rc != SB_OK) {
        cudaFree(srs->d_g);
        cudaFree(srs->d_g_lagrange);
        delete srs;
        return rc;
    }
    *out_srs = srs;
    return SB_OK;
}
int32_t sb_srs_downsize(sb_ctx *ctx, const sb_srs *srs, uint32_t new_k, sb_srs **out_srs) {
    if (!ctx || !srs || !out_srs) return SB_ERR_ARG;
    SB_REQUIRE(new_k <= srs->k, "sb_srs_downsize: the new k must not exceed the SRS's k (halo2 asserts k <= self.k)");
    CtxGuard g(ctx);
    cudaStream_t st = ctx->stream;
    const size_t n = (size_t)1 << new_k;
    sb_domain *dom = nullptr;
    SB_TRY(sb_domain_create(ctx, 2, new_k, &dom));
    sb_srs *out = new sb_srs();
    out->k = new_k;
    if (cudaMalloc(&out->d_g, n * 64) != cudaSuccess || cudaMalloc(&out->d_g_lagrange, n * 64) != cudaSuccess) {
        set_last_error("sb_srs_downsize: cudaMalloc(2 x %zu) failed", n * 64);
        if (out->d_g) cudaFree(out->d_g);
        delete out;
        sb_domain_destroy(dom);
        return SB_ERR_ALLOC;
    }
    int32_t rc = SB_OK;
    if (cudaMemcpyAsync(out->d_g, srs->d_g, n * 64, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = SB_ERR_CUDA;
    if (rc == SB_OK) {
        if (new_k == srs->k) {
            if (cudaMemcpyAsync(out->d_g_lagrange, srs->d_g_lagrange, n * 64, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = SB_ERR_CUDA;
        } else {
            rc = g1_to_lagrange(ctx, out->d_g, new_k, dom->omega_inv, dom->ifft_divisor, out->d_g_lagrange, st);
        }
    }
    if (rc == SB_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = SB_ERR_CUDA;
    sb_domain_destroy(dom);
    if (rc != SB_OK) {
        cudaFree(out->d_g);
        cudaFree(out->d_g_lagrange);
        delete out;
        return rc;
    }
    *out_srs = out;
    return SB_OK;
}
int32_t sb_srs_download(sb_ctx *ctx, const sb_srs *srs, uint8_t *g_out, uint8_t *g_lagrange_out) {
    if (!ctx || !srs || !g_out || !g_lagrange_out) return SB_ERR_ARG;
    CtxGuard g(ctx);
    const size_t bytes = (size_t)64 << srs->k;
    SB_CUDA_TRY(cudaMemcpyAsync(g_out, srs->d_g, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaMemcpyAsync(g_lagrange_out, srs->d_g_lagrange, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

// CPU-only check of the evaluate_h compiler: builds the quotient-numerator program of a constraint system (challenges theta, beta,
// gamma, y derived from `seed`), evaluates it for one row of pseudo-random column values with the host interpreter AND by walking
// the expression trees directly (no CSE, no slots), and returns both values plus the program shape.
static hfr::Fr direct_eval(const ExprP &e, const std::map<std::pair<int, int>, hfr::Fr> &vals) {
    switch (e->kind) {
        case Expr::CONST: return to_host(e->c);
        case Expr::COL: return vals.at({e->col, e->rot});
        case Expr::NEG: return hfr::neg(direct_eval(e->a, vals));
        case Expr::ADD: return hfr::add(direct_eval(e->a, vals), direct_eval(e->b, vals));
        case Expr::SUB: return hfr::sub(direct_eval(e->a, vals), direct_eval(e->b, vals));
        default: return hfr::mul(direct_eval(e->a, vals), direct_eval(e->b, vals));
    }
}
static void collect_inputs(const ExprP &e, std::map<std::pair<int, int>, hfr::Fr> &vals, ChaCha20Rng &rng) {
    if (!e) return;
    if (e->kind == Expr::COL) {
        if (!vals.count({e->col, e->rot})) vals[{e->col, e->rot}] = rng.next_fr();
        return;
    }
    collect_inputs(e->a, vals, rng);
    collect_inputs(e->b, vals, rng);
}
int32_t sb_test_h_program(const char *cs_json, uint64_t seed, uint8_t out_program_value[32], uint8_t out_direct_value[32], uint32_t out_shape[4]) {
    if (!cs_json || !out_program_value || !out_direct_value || !out_shape) return SB_ERR_ARG;
    try {
        ConstraintSystem cs = parse_cs(cs_json);
        ChaCha20Rng rng;
        rng.seed_from_u64(seed);
        const Fr theta = rng.next_fr(), beta = rng.next_fr(), gamma = rng.next_fr(), y = rng.next_fr();
        const int P = (int)cs.perm_cols.size(), chunk = cs.degree - 2;
        std::vector<std::pair<int, int>> sets;
        for (int f = 0; f < P; f += chunk) sets.push_back({f, std::min(chunk, P - f)});
        const std::vector<ExprP> terms = h_terms(cs, P, sets, cs.lookups.size(), theta, beta, gamma);
        fr_t yd = to_dev(y);
        Program prog = compile_terms(terms, &yd);
        std::map<std::pair<int, int>, hfr::Fr> vals;
        for (auto &t : terms) collect_inputs(t, vals, rng);
        std::vector<fr_t> inputs(prog.inputs.size() / 2);
        for (size_t i = 0; i < inputs.size(); i++) inputs[i] = to_dev(vals.at({prog.inputs[2 * i], prog.inputs[2 * i + 1]}));
        fr_t got = program_eval_host(prog, inputs);
        hfr::Fr acc = hfr::ZERO;
        for (size_t i = 0; i < terms.size(); i++) acc = hfr::add(hfr::mul(acc, y), direct_eval(terms[i], vals));
        // the per-key cached program (compiled with placeholder challenges, constants patched) must give the same value, twice
        {
            sb_pk fake;
            fake.cs = cs;
            fake.P = P;
            for (int rep = 0; rep < 2; rep++) {
                Program cached = h_program_for(&fake, sets, cs.lookups.size(), theta, beta, gamma, y);
                std::vector<fr_t> in2(cached.inputs.size() / 2);
                for (size_t i = 0; i < in2.size(); i++) in2[i] = to_dev(vals.at({cached.inputs[2 * i], cached.inputs[2 * i + 1]}));
                if (!(to_host(program_eval_host(cached, in2)) == acc)) {
                    set_last_error("sb_test_h_program: cached program (constants patched) disagrees with the direct evaluation");
                    return SB_ERR_ARG;
                }
            }
        }
        memcpy(out_program_value, got.v, 32);
        memcpy(out_direct_value, acc.v, 32);
        out_shape[0] = (uint32_t)(prog.code.size() / 3);
        out_shape[1] = prog.n_mul;
        out_shape[2] = prog.n_addsub;
        out_shape[3] = prog.n_slots;
    } catch (const std::exception &e) {
        set_last_error("sb_test_h_program: %s", e.what());
        return SB_ERR_ARG;
    }
    return SB_OK;
}

// CPU-side check of the evaluate_h code generator: the CUDA source the JIT would compile for this constraint system
int32_t sb_test_h_jit_source(const char *cs_json, char *out, size_t cap, size_t *out_len) {
    if (!cs_json || !out_len) return SB_ERR_ARG;
    try {
        ConstraintSystem cs = parse_cs(cs_json);
        const int P = (int)cs.perm_cols.size(), chunk = cs.degree - 2;
        std::vector<std::pair<int, int>> sets;
        for (int f = 0; f < P; f += chunk) sets.push_back({f, std::min(chunk, P - f)});
        sb_pk fake;
        fake.cs = cs;
        fake.P = P;
        const Program prog = h_program_for(&fake, sets, cs.lookups.size(), hfr::ONE, hfr::ONE, hfr::ONE, hfr::ONE);
        const std::string src = expr_jit_source(prog);
        *out_len = src.size();
        if (out && cap >= src.size()) memcpy(out, src.data(), src.size());
    } catch (const std::exception &e) {
        set_last_error("sb_test_h_jit_source: %s", e.what());
        return SB_ERR_ARG;
    }
    return SB_OK;
}
int32_t sb_last_proof_msm(const sb_ctx *ctx, float out_ms[5], uint64_t *out_digits, uint32_t *out_launch_sets) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    for (int i = 0; i < 5; i++) out_ms[i] = ctx->acc_msm_ms[i];
    if (out_digits) *out_digits = ctx->acc_msm_digits;
    if (out_launch_sets) *out_launch_sets = ctx->acc_msm_sets;
    return SB_OK;
}
int32_t sb_last_proof_d2h(const sb_ctx *ctx, uint64_t *out_bytes) {
    if (!ctx || !out_bytes) return SB_ERR_ARG;
    *out_bytes = ctx->acc_msm_d2h;
    return SB_OK;
}
int32_t sb_last_proof_stages(const sb_ctx *ctx, float out_ms[12]) {
    if (!ctx || !out_ms) return SB_ERR_ARG;
    for (int i = 0; i < 12; i++) out_ms[i] = ctx->last_proof_stage_ms[i];
    return SB_OK;
}
int32_t sb_last_h_rows(const sb_ctx *ctx, uint64_t *out_rows) {
    if (!ctx || !out_rows) return SB_ERR_ARG;
    *out_rows = ctx->last_h_rows;
    return SB_OK;
}
int32_t sb_last_h_jit(const sb_ctx *ctx, int32_t *out_used) {
    if (!ctx || !out_used) return SB_ERR_ARG;
    *out_used = ctx->last_h_jit ? 1 : 0;
    return SB_OK;
}
int32_t sb_last_h_profile(const sb_ctx *ctx, float *out_ms, uint32_t out_program[4]) {
    if (!ctx || !out_ms || !out_program) return SB_ERR_ARG;
    *out_ms = ctx->last_h_ms;
    for (int i = 0; i < 4; i++) out_program[i] = ctx->last_h_program[i];
    return SB_OK;
}

}  // extern "C"
