/* CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the CPU algorithms the reference runs for the hot path.  The code
 * itself lives in crates that are NOT vendored under /root/reference:
 *   halo2curves 0.1.0 (crates.io, Cargo.lock:2272-2276)          bn256::{Fr,Fq,G1Affine,G1}
 *   halo2_proofs 0.2.0 @ summa-dev/halo2#8386d6e (Cargo.lock:2239-2255)
 *       arithmetic::{best_multiexp, multiexp_serial, best_fft, recursive_butterfly_arithmetic}
 *       poly::EvaluationDomain::{lagrange_to_coeff, coeff_to_extended, extended_to_coeff,
 *                                divide_by_vanishing_poly}
 * Reference call sites: zk_prover/src/circuits/utils.rs:55,64,70,75,76,94-102,171-178.
 * The published algorithms are restated from SURVEY.md Appendix A.1-A.4.
 *
 * Pinning: oracle/bn254.py (python big ints) is checked against the reference's golden vectors
 * (tests/test_oracle_golden.py: verifier-contract constants, SRS file, fixed_comms[4] MSM KAT);
 * this C twin is checked against bn254.py on random inputs and on the same KATs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Data layout = halo2curves memory layout: 4 x u64 LE limbs, Montgomery form.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

typedef struct { u64 m[4]; u64 inv; u64 r[4]; u64 r2[4]; } field_t;

static field_t FR, FQ;
static int g_init = 0;

/* ------------------------------------------------------------------ field arithmetic */
static inline int geq(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
static inline void sub_nc(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (u64)d; br = (u64)(d >> 64) & 1; }
}
static inline void f_add(const field_t *F, u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 c = 0, t[4];
    for (int i = 0; i < 4; i++) { u128 s = (u128)a[i] + b[i] + c; t[i] = (u64)s; c = (u64)(s >> 64); }
    if (c || geq(t, F->m)) sub_nc(r, t, F->m); else memcpy(r, t, 32);
}
static inline void f_sub(const field_t *F, u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 br = 0, t[4];
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - br; t[i] = (u64)d; br = (u64)(d >> 64) & 1; }
    if (br) { u64 c = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)t[i] + F->m[i] + c; t[i] = (u64)s; c = (u64)(s >> 64); } }
    memcpy(r, t, 32);
}
static inline int f_is_zero(const u64 a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int f_eq(const u64 a[4], const u64 b[4]) { return a[0]==b[0] && a[1]==b[1] && a[2]==b[2] && a[3]==b[3]; }

/* Montgomery product a*b*2^-256 mod m (CIOS, 4x64 limbs) */
static inline void f_mul(const field_t *F, u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u64 carry = 0; u128 acc;
        for (int j = 0; j < 4; j++) { acc = (u128)a[j] * b[i] + t[j] + carry; t[j] = (u64)acc; carry = (u64)(acc >> 64); }
        acc = (u128)t[4] + carry; t[4] = (u64)acc; t[5] = (u64)(acc >> 64);
        u64 mm = t[0] * F->inv;
        acc = (u128)mm * F->m[0] + t[0]; carry = (u64)(acc >> 64);
        for (int j = 1; j < 4; j++) { acc = (u128)mm * F->m[j] + t[j] + carry; t[j - 1] = (u64)acc; carry = (u64)(acc >> 64); }
        acc = (u128)t[4] + carry; t[3] = (u64)acc; t[4] = t[5] + (u64)(acc >> 64);
    }
    if (t[4] || geq(t, F->m)) sub_nc(r, t, F->m); else memcpy(r, t, 32);
}
static inline void f_sqr(const field_t *F, u64 r[4], const u64 a[4]) { f_mul(F, r, a, a); }
static inline void f_dbl(const field_t *F, u64 r[4], const u64 a[4]) { f_add(F, r, a, a); }
static void f_pow(const field_t *F, u64 r[4], const u64 a[4], const u64 e[4]) {
    u64 acc[4]; memcpy(acc, F->r, 32);
    for (int i = 255; i >= 0; i--) {
        f_sqr(F, acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, acc, acc, a);
    }
    memcpy(r, acc, 32);
}
static void f_inv(const field_t *F, u64 r[4], const u64 a[4]) {
    u64 e[4] = {2, 0, 0, 0}, pm2[4];
    sub_nc(pm2, F->m, e);
    f_pow(F, r, a, pm2);
}
static void f_from_canon(const field_t *F, u64 r[4], const u64 a[4]) { f_mul(F, r, a, F->r2); }
static void f_to_canon(const field_t *F, u64 r[4], const u64 a[4]) { u64 one[4] = {1, 0, 0, 0}; f_mul(F, r, a, one); }

static void field_setup(field_t *F, const u64 m[4]) {
    memcpy(F->m, m, 32);
    u64 x = 1; /* Newton: x = m^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) x *= 2 - m[0] * x;
    F->inv = (u64)0 - x;
    /* R = 2^256 mod m, R2 = 2^512 mod m by repeated doubling */
    u64 t[4] = {1, 0, 0, 0};
    for (int i = 0; i < 256; i++) f_add(F, t, t, t);
    memcpy(F->r, t, 32);
    for (int i = 0; i < 256; i++) f_add(F, t, t, t);
    memcpy(F->r2, t, 32);
}

void oracle_init(void) {
    if (g_init) return;
    /* contracts/src/InclusionVerifier.sol:209-210 */
    static const u64 R_MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static const u64 Q_MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    field_setup(&FR, R_MOD);
    field_setup(&FQ, Q_MOD);
    g_init = 1;
}

/* vector helpers exported for tests (all Montgomery in/out unless stated) */
void oracle_fr_mul(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_mul(&FR, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_fr_add(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_add(&FR, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_fr_sub(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_sub(&FR, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_fq_mul(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_mul(&FQ, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_fq_add(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_add(&FQ, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_fq_sub(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_sub(&FQ, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_fr_inv(u64 *r, const u64 *a, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_inv(&FR, r + 4 * i, a + 4 * i); }
void oracle_fr_from_canon(u64 *r, const u64 *a, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_from_canon(&FR, r + 4 * i, a + 4 * i); }
void oracle_fr_to_canon(u64 *r, const u64 *a, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_to_canon(&FR, r + 4 * i, a + 4 * i); }
void oracle_fq_from_canon(u64 *r, const u64 *a, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_from_canon(&FQ, r + 4 * i, a + 4 * i); }
void oracle_fq_to_canon(u64 *r, const u64 *a, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_to_canon(&FQ, r + 4 * i, a + 4 * i); }

/* ------------------------------------------------------------------ G1 (Jacobian, a = 0, b = 3) */
typedef struct { u64 x[4], y[4], z[4]; } jac_t;   /* identity: z == 0 */
typedef struct { u64 x[4], y[4]; } aff_t;         /* identity: (0, 0) (halo2curves G1Affine) */

static inline int aff_is_id(const aff_t *p) { return f_is_zero(p->x) && f_is_zero(p->y); }
static inline void jac_set_id(jac_t *p) { memset(p, 0, sizeof *p); memcpy(p->y, FQ.r, 32); }
static inline int jac_is_id(const jac_t *p) { return f_is_zero(p->z); }

static void jac_double(jac_t *r, const jac_t *p) {
    if (jac_is_id(p)) { *r = *p; return; }
    const field_t *F = &FQ;
    u64 a[4], b[4], c[4], d[4], e[4], f[4], t[4], x3[4], y3[4], z3[4];
    f_sqr(F, a, p->x); f_sqr(F, b, p->y); f_sqr(F, c, b);
    f_add(F, d, p->x, b); f_sqr(F, d, d); f_sub(F, d, d, a); f_sub(F, d, d, c); f_dbl(F, d, d);
    f_dbl(F, e, a); f_add(F, e, e, a);
    f_sqr(F, f, e);
    f_mul(F, z3, p->y, p->z); f_dbl(F, z3, z3);
    f_dbl(F, t, d); f_sub(F, x3, f, t);
    f_sub(F, t, d, x3); f_mul(F, y3, e, t);
    f_dbl(F, c, c); f_dbl(F, c, c); f_dbl(F, c, c); f_sub(F, y3, y3, c);
    memcpy(r->x, x3, 32); memcpy(r->y, y3, 32); memcpy(r->z, z3, 32);
}
static void jac_add(jac_t *r, const jac_t *p, const jac_t *q) {
    if (jac_is_id(p)) { *r = *q; return; }
    if (jac_is_id(q)) { *r = *p; return; }
    const field_t *F = &FQ;
    u64 z1z1[4], z2z2[4], u1[4], u2[4], s1[4], s2[4], h[4], rr[4], hh[4], hhh[4], v[4], t[4], x3[4], y3[4], z3[4];
    f_sqr(F, z1z1, p->z); f_sqr(F, z2z2, q->z);
    f_mul(F, u1, p->x, z2z2); f_mul(F, u2, q->x, z1z1);
    f_mul(F, s1, p->y, q->z); f_mul(F, s1, s1, z2z2);
    f_mul(F, s2, q->y, p->z); f_mul(F, s2, s2, z1z1);
    if (f_eq(u1, u2)) {
        if (f_eq(s1, s2)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    f_sub(F, h, u2, u1); f_sub(F, rr, s2, s1);
    f_sqr(F, hh, h); f_mul(F, hhh, h, hh); f_mul(F, v, u1, hh);
    f_sqr(F, x3, rr); f_sub(F, x3, x3, hhh); f_dbl(F, t, v); f_sub(F, x3, x3, t);
    f_sub(F, t, v, x3); f_mul(F, y3, rr, t); f_mul(F, t, s1, hhh); f_sub(F, y3, y3, t);
    f_mul(F, z3, p->z, q->z); f_mul(F, z3, z3, h);
    memcpy(r->x, x3, 32); memcpy(r->y, y3, 32); memcpy(r->z, z3, 32);
}
static void jac_add_affine(jac_t *r, const jac_t *p, const aff_t *q) {
    if (aff_is_id(q)) { *r = *p; return; }
    if (jac_is_id(p)) { memcpy(r->x, q->x, 32); memcpy(r->y, q->y, 32); memcpy(r->z, FQ.r, 32); return; }
    const field_t *F = &FQ;
    u64 z1z1[4], u2[4], s2[4], h[4], rr[4], hh[4], hhh[4], v[4], t[4], x3[4], y3[4], z3[4];
    f_sqr(F, z1z1, p->z);
    f_mul(F, u2, q->x, z1z1);
    f_mul(F, s2, q->y, p->z); f_mul(F, s2, s2, z1z1);
    if (f_eq(p->x, u2)) {
        if (f_eq(p->y, s2)) { jac_double(r, p); return; }
        jac_set_id(r); return;
    }
    f_sub(F, h, u2, p->x); f_sub(F, rr, s2, p->y);
    f_sqr(F, hh, h); f_mul(F, hhh, h, hh); f_mul(F, v, p->x, hh);
    f_sqr(F, x3, rr); f_sub(F, x3, x3, hhh); f_dbl(F, t, v); f_sub(F, x3, x3, t);
    f_sub(F, t, v, x3); f_mul(F, y3, rr, t); f_mul(F, t, p->y, hhh); f_sub(F, y3, y3, t);
    f_mul(F, z3, p->z, h);
    memcpy(r->x, x3, 32); memcpy(r->y, y3, 32); memcpy(r->z, z3, 32);
}
static void jac_to_affine(aff_t *r, const jac_t *p) {
    if (jac_is_id(p)) { memset(r, 0, sizeof *r); return; }
    const field_t *F = &FQ;
    u64 zi[4], zi2[4], zi3[4];
    f_inv(F, zi, p->z); f_sqr(F, zi2, zi); f_mul(F, zi3, zi2, zi);
    f_mul(F, r->x, p->x, zi2); f_mul(F, r->y, p->y, zi3);
}

/* affine a + b (out affine), for tests */
void oracle_g1_add_affine(u64 out[8], const u64 a[8], const u64 b[8]) {
    oracle_init();
    jac_t acc; jac_set_id(&acc);
    jac_add_affine(&acc, &acc, (const aff_t *)a);
    jac_add_affine(&acc, &acc, (const aff_t *)b);
    jac_to_affine((aff_t *)out, &acc);
}
/* affine scalar multiplication k*P, k given as a Montgomery Fr (like Rust `P * k`) */
void oracle_g1_mul(u64 out[8], const u64 p[8], const u64 k_mont[4]) {
    oracle_init();
    u64 k[4]; f_to_canon(&FR, k, k_mont);
    jac_t acc; jac_set_id(&acc);
    for (int i = 255; i >= 0; i--) {
        jac_double(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) jac_add_affine(&acc, &acc, (const aff_t *)p);
    }
    jac_to_affine((aff_t *)out, &acc);
}

/* ------------------------------------------------------------------ best_multiexp */
/* halo2 `multiexp_serial`: c = 1 (n<4) | 3 (n<32) | ceil(ln n); segments = 256/c + 1, processed
 * MSB -> LSB with c doublings each; buckets [2^c - 1]; running-sum reduction. */
static unsigned get_at(unsigned segment, unsigned c, const uint8_t bytes[32]) {
    unsigned skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    for (unsigned i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
    u64 tmp; memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    tmp %= (1ULL << c);
    return (unsigned)tmp;
}
static void multiexp_serial(const u64 *coeffs, const aff_t *bases, size_t n, jac_t *acc) {
    unsigned c;
    if (n < 4) c = 1; else if (n < 32) c = 3; else c = (unsigned)ceil(log((double)n));
    uint8_t *repr = (uint8_t *)malloc(32 * (n ? n : 1));
    for (size_t i = 0; i < n; i++) { u64 t[4]; f_to_canon(&FR, t, coeffs + 4 * i); memcpy(repr + 32 * i, t, 32); }
    unsigned segments = 256 / c + 1;
    size_t nb = ((size_t)1 << c) - 1;
    jac_t *buckets = (jac_t *)malloc(sizeof(jac_t) * nb);
    for (int seg = (int)segments - 1; seg >= 0; seg--) {
        for (unsigned k = 0; k < c; k++) jac_double(acc, acc);
        for (size_t b = 0; b < nb; b++) jac_set_id(&buckets[b]);
        for (size_t i = 0; i < n; i++) {
            unsigned d = get_at((unsigned)seg, c, repr + 32 * i);
            if (d) jac_add_affine(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        }
        jac_t run; jac_set_id(&run);
        for (size_t b = nb; b-- > 0;) { jac_add(&run, &run, &buckets[b]); jac_add(acc, acc, &run); }
    }
    free(buckets); free(repr);
}
typedef struct { const u64 *coeffs; const aff_t *bases; size_t n; jac_t acc; } msm_job_t;
static void *msm_worker(void *p) { msm_job_t *j = (msm_job_t *)p; jac_set_id(&j->acc); multiexp_serial(j->coeffs, j->bases, j->n, &j->acc); return NULL; }

/* `best_multiexp`: n > threads -> contiguous chunks of n/threads, partial sums folded. out = affine. */
void oracle_best_multiexp(u64 out_affine[8], const u64 *coeffs, const u64 *bases, size_t n, int threads) {
    oracle_init();
    jac_t total; jac_set_id(&total);
    if (threads < 1) threads = 1;
    if (n > (size_t)threads && threads > 1) {
        size_t chunk = n / (size_t)threads;
        size_t njobs = (n + chunk - 1) / chunk;
        msm_job_t *jobs = (msm_job_t *)calloc(njobs, sizeof(msm_job_t));
        pthread_t *tids = (pthread_t *)calloc(njobs, sizeof(pthread_t));
        for (size_t j = 0; j < njobs; j++) {
            size_t lo = j * chunk, hi = lo + chunk > n ? n : lo + chunk;
            jobs[j].coeffs = coeffs + 4 * lo; jobs[j].bases = (const aff_t *)bases + lo; jobs[j].n = hi - lo;
            pthread_create(&tids[j], NULL, msm_worker, &jobs[j]);
        }
        for (size_t j = 0; j < njobs; j++) { pthread_join(tids[j], NULL); jac_add(&total, &total, &jobs[j].acc); }
        free(jobs); free(tids);
    } else {
        multiexp_serial(coeffs, (const aff_t *)bases, n, &total);
    }
    jac_to_affine((aff_t *)out_affine, &total);
}

/* ------------------------------------------------------------------ best_fft */
static inline size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0; for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r;
}
typedef struct { u64 *a; size_t n; size_t chunk; const u64 *tw; int depth; } fft_job_t;
static void recursive_butterfly(u64 *a, size_t n, size_t twiddle_chunk, const u64 *tw, int spawn_depth);
static void *fft_worker(void *p) { fft_job_t *j = (fft_job_t *)p; recursive_butterfly(j->a, j->n, j->chunk, j->tw, j->depth); return NULL; }
/* halo2 `recursive_butterfly_arithmetic`: halves via join, then one combine layer */
static void recursive_butterfly(u64 *a, size_t n, size_t twiddle_chunk, const u64 *tw, int spawn_depth) {
    const field_t *F = &FR;
    if (n == 2) {
        u64 t[4]; memcpy(t, a + 4, 32);
        memcpy(a + 4, a, 32);
        f_add(F, a, a, t); f_sub(F, a + 4, a + 4, t);
        return;
    }
    u64 *left = a, *right = a + 4 * (n / 2);
    if (spawn_depth > 0) {
        fft_job_t job = {left, n / 2, twiddle_chunk * 2, tw, spawn_depth - 1};
        pthread_t tid; pthread_create(&tid, NULL, fft_worker, &job);
        recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, spawn_depth - 1);
        pthread_join(tid, NULL);
    } else {
        recursive_butterfly(left, n / 2, twiddle_chunk * 2, tw, 0);
        recursive_butterfly(right, n / 2, twiddle_chunk * 2, tw, 0);
    }
    for (size_t i = 0; i < n / 2; i++) {
        u64 t[4];
        if (i == 0) memcpy(t, right, 32); else f_mul(F, t, right + 4 * i, tw + 4 * (i * twiddle_chunk));
        u64 u[4]; memcpy(u, left + 4 * i, 32);
        f_add(F, left + 4 * i, u, t); f_sub(F, right + 4 * i, u, t);
    }
}
void oracle_best_fft(u64 *a, const u64 omega[4], uint32_t log_n, int threads) {
    oracle_init();
    size_t n = (size_t)1 << log_n;
    if (n == 1) return;
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, log_n);
        if (i < j) { u64 t[4]; memcpy(t, a + 4 * i, 32); memcpy(a + 4 * i, a + 4 * j, 32); memcpy(a + 4 * j, t, 32); }
    }
    size_t half = n / 2;
    u64 *tw = (u64 *)malloc(32 * half);
    memcpy(tw, FR.r, 32);
    for (size_t i = 1; i < half; i++) f_mul(&FR, tw + 4 * i, tw + 4 * (i - 1), omega);
    int depth = 0; while ((1 << (depth + 1)) <= threads && (size_t)4 << depth <= n) depth++;
    if (threads <= 1) depth = 0;
    recursive_butterfly(a, n, 1, tw, depth);
    free(tw);
}

/* ------------------------------------------------------------------ EvaluationDomain pieces */
/* a[i] *= s  */
void oracle_fr_scale(u64 *a, const u64 s[4], size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_mul(&FR, a + 4 * i, a + 4 * i, s); }
/* a[i] *= pat[i % m]   (coset zeta pattern, t_inv pattern) */
void oracle_fr_scale_pattern(u64 *a, const u64 *pat, size_t m, size_t n) { oracle_init(); for (size_t i = 0; i < n; i++) f_mul(&FR, a + 4 * i, a + 4 * i, pat + 4 * (i % m)); }

/* ------------------------------------------------------------------ synthetic bases (test / bench inputs) */
/* out[i] = (s + i*t) * G for i < n, affine Montgomery: distinct valid curve points in O(n) field
 * work (one mixed addition each + batched normalisation).  NOT an SRS: throughput/parity input only. */
typedef struct { u64 *out; size_t lo, hi; u64 s[4], t[4]; } gen_job_t;
static void fr_add_small_mul(u64 r[4], const u64 s[4], const u64 t[4], u64 i) {
    /* r = s + i*t mod r (all Montgomery) */
    u64 ic[4] = {i, 0, 0, 0}, im[4], it[4];
    f_from_canon(&FR, im, ic); f_mul(&FR, it, im, t); f_add(&FR, r, s, it);
}
static void *gen_worker(void *p) {
    gen_job_t *j = (gen_job_t *)p;
    const size_t BATCH = 1024;
    aff_t g; memset(&g, 0, sizeof g);
    u64 one_c[4] = {1, 0, 0, 0}, two_c[4] = {2, 0, 0, 0};
    f_from_canon(&FQ, g.x, one_c); f_from_canon(&FQ, g.y, two_c);
    aff_t d, start; u64 k0[4];
    oracle_g1_mul((u64 *)&d, (const u64 *)&g, j->t);
    fr_add_small_mul(k0, j->s, j->t, (u64)j->lo);
    oracle_g1_mul((u64 *)&start, (const u64 *)&g, k0);
    jac_t cur; memcpy(cur.x, start.x, 32); memcpy(cur.y, start.y, 32); memcpy(cur.z, FQ.r, 32);
    if (aff_is_id(&start)) jac_set_id(&cur);
    jac_t *buf = (jac_t *)malloc(sizeof(jac_t) * BATCH);
    u64 *pref = (u64 *)malloc(32 * BATCH);
    for (size_t base = j->lo; base < j->hi; base += BATCH) {
        size_t m = j->hi - base < BATCH ? j->hi - base : BATCH;
        for (size_t i = 0; i < m; i++) { buf[i] = cur; jac_add_affine(&cur, &cur, &d); }
        /* batch inversion of z (identity cannot occur for the parameters the tests use) */
        u64 acc[4]; memcpy(acc, FQ.r, 32);
        for (size_t i = 0; i < m; i++) { memcpy(pref + 4 * i, acc, 32); f_mul(&FQ, acc, acc, buf[i].z); }
        u64 inv[4]; f_inv(&FQ, inv, acc);
        for (size_t i = m; i-- > 0;) {
            u64 zi[4], zi2[4], zi3[4];
            f_mul(&FQ, zi, inv, pref + 4 * i); f_mul(&FQ, inv, inv, buf[i].z);
            f_sqr(&FQ, zi2, zi); f_mul(&FQ, zi3, zi2, zi);
            u64 *o = j->out + 8 * (base + i);
            f_mul(&FQ, o, buf[i].x, zi2); f_mul(&FQ, o + 4, buf[i].y, zi3);
        }
    }
    free(buf); free(pref);
    return NULL;
}
void oracle_g1_gen_bases(u64 *out, size_t n, const u64 s_mont[4], const u64 t_mont[4], int threads) {
    oracle_init();
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = n ? (int)n : 1;
    gen_job_t *jobs = (gen_job_t *)calloc((size_t)threads, sizeof(gen_job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    size_t per = (n + (size_t)threads - 1) / (size_t)threads;
    int used = 0;
    for (int t = 0; t < threads; t++) {
        size_t lo = (size_t)t * per, hi = lo + per > n ? n : lo + per;
        if (lo >= hi) break;
        jobs[t].out = out; jobs[t].lo = lo; jobs[t].hi = hi;
        memcpy(jobs[t].s, s_mont, 32); memcpy(jobs[t].t, t_mont, 32);
        pthread_create(&tids[t], NULL, gen_worker, &jobs[t]); used++;
    }
    for (int t = 0; t < used; t++) pthread_join(tids[t], NULL);
    free(jobs); free(tids);
}

/* ------------------------------------------------------------------ O(n) prover helpers */
/* halo2 `batch_invert`: zeros stay zero */
void oracle_fr_batch_invert(u64 *a, size_t n) {
    oracle_init();
    u64 *pref = (u64 *)malloc(32 * (n ? n : 1));
    u64 acc[4]; memcpy(acc, FR.r, 32);
    for (size_t i = 0; i < n; i++) { memcpy(pref + 4 * i, acc, 32); if (!f_is_zero(a + 4 * i)) f_mul(&FR, acc, acc, a + 4 * i); }
    u64 inv[4]; f_inv(&FR, inv, acc);
    for (size_t i = n; i-- > 0;) {
        if (f_is_zero(a + 4 * i)) continue;
        u64 t[4]; f_mul(&FR, t, inv, pref + 4 * i); f_mul(&FR, inv, inv, a + 4 * i); memcpy(a + 4 * i, t, 32);
    }
    free(pref);
}
/* out[0] = init, out[i] = out[i-1] * a[i-1]  for i < n  (grand-product column Z; a has >= n-1 entries) */
void oracle_fr_running_product(u64 *out, const u64 *a, const u64 init[4], size_t n) {
    oracle_init();
    if (!n) return;
    memcpy(out, init, 32);
    for (size_t i = 1; i < n; i++) f_mul(&FR, out + 4 * i, out + 4 * (i - 1), a + 4 * (i - 1));
}
/* halo2 `eval_polynomial`: Horner */
void oracle_fr_eval_poly(u64 out[4], const u64 *coeffs, size_t n, const u64 x[4]) {
    oracle_init();
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t i = n; i-- > 0;) { f_mul(&FR, acc, acc, x); f_add(&FR, acc, acc, coeffs + 4 * i); }
    memcpy(out, acc, 32);
}
/* halo2 `kate_division`: q = (a - a(b)) / (X - b), n-1 coefficients */
void oracle_fr_kate_division(u64 *q, const u64 *a, size_t n, const u64 b[4]) {
    oracle_init();
    if (n < 2) return;
    u64 tmp[4] = {0, 0, 0, 0};
    for (size_t i = n - 1; i-- > 0;) {
        u64 lead[4];
        f_add(&FR, lead, a + 4 * (i + 1), tmp);      /* r - tmp with b negated == r + b * prev */
        memcpy(q + 4 * i, lead, 32);
        f_mul(&FR, tmp, lead, b);
    }
}
/* r[i] = a[i] * s + b[i] * t   (polynomial linear combinations) */
void oracle_fr_axpby(u64 *r, const u64 *a, const u64 s[4], const u64 *b, const u64 t[4], size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) { u64 x[4], y[4]; f_mul(&FR, x, a + 4 * i, s); f_mul(&FR, y, b + 4 * i, t); f_add(&FR, r + 4 * i, x, y); }
}

/* ====================================================================================================
 * Second half: what the restated `create_proof` (oracle/halo2_prover.py) needs to run at k = 17 .. 20 in
 * seconds on all host cores, the way halo2 runs it with rayon: element-wise passes split over threads
 * (`parallelize`), `Evaluator::evaluate_h` as a row-parallel interpreter of a flattened calculation list
 * (halo2 `GraphEvaluator`), chunk-parallel `eval_polynomial` / `batch_invert`, the lookup
 * argument's `permute_expression_pair`, ChaCha20 `Fr::random` streams, and the Merkle sum tree's
 * Keccak-256 + Poseidon hashing (zk_prover/src/merkle_sum_tree/{entry.rs:15-27,node.rs:16-85},
 * utils/build_tree.rs:5-78).  Threads: OpenMP, `oracle_set_threads`.
 * ==================================================================================================== */
#include <omp.h>

static int g_threads = 1;
void oracle_set_threads(int t) { g_threads = t < 1 ? 1 : t; }
int oracle_get_threads(void) { return g_threads; }

#define PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(g_threads)")

void oracle_par_fr_mul(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); PAR_FOR for (size_t i = 0; i < n; i++) f_mul(&FR, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_par_fr_add(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); PAR_FOR for (size_t i = 0; i < n; i++) f_add(&FR, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_par_fr_sub(u64 *r, const u64 *a, const u64 *b, size_t n) { oracle_init(); PAR_FOR for (size_t i = 0; i < n; i++) f_sub(&FR, r + 4 * i, a + 4 * i, b + 4 * i); }
void oracle_par_fr_scale(u64 *a, const u64 s[4], size_t n) { oracle_init(); PAR_FOR for (size_t i = 0; i < n; i++) f_mul(&FR, a + 4 * i, a + 4 * i, s); }
void oracle_par_fr_scale_pattern(u64 *a, const u64 *pat, size_t m, size_t n) { oracle_init(); PAR_FOR for (size_t i = 0; i < n; i++) f_mul(&FR, a + 4 * i, a + 4 * i, pat + 4 * (i % m)); }
void oracle_par_fr_add_const(u64 *r, const u64 *a, const u64 c[4], size_t n) { oracle_init(); PAR_FOR for (size_t i = 0; i < n; i++) f_add(&FR, r + 4 * i, a + 4 * i, c); }
void oracle_par_fr_axpby(u64 *r, const u64 *a, const u64 s[4], const u64 *b, const u64 t[4], size_t n) {
    oracle_init();
    PAR_FOR for (size_t i = 0; i < n; i++) { u64 x[4], y[4]; f_mul(&FR, x, a + 4 * i, s); f_mul(&FR, y, b + 4 * i, t); f_add(&FR, r + 4 * i, x, y); }
}
/* out[i] = base^i (Montgomery), i < n: each thread starts its chunk with one exponentiation */
void oracle_fr_powers(u64 *out, const u64 base[4], size_t n) {
    oracle_init();
    const size_t CH = 4096;
    const size_t nch = (n + CH - 1) / CH;
    PAR_FOR for (size_t c = 0; c < nch; c++) {
        size_t lo = c * CH, hi = lo + CH > n ? n : lo + CH;
        u64 e[4] = {lo, 0, 0, 0}, cur[4];
        f_pow(&FR, cur, base, e);
        for (size_t i = lo; i < hi; i++) { memcpy(out + 4 * i, cur, 32); f_mul(&FR, cur, cur, base); }
    }
}
/* halo2 `batch_invert` under `parallelize`: one Montgomery-trick chain per chunk (zeros stay zero) */
void oracle_par_fr_batch_invert(u64 *a, size_t n) {
    oracle_init();
    const size_t CH = 1 << 14;
    const size_t nch = (n + CH - 1) / CH;
    PAR_FOR for (size_t c = 0; c < nch; c++) {
        size_t lo = c * CH, hi = lo + CH > n ? n : lo + CH;
        oracle_fr_batch_invert(a + 4 * lo, hi - lo);
    }
}
/* halo2 `eval_polynomial`: chunks evaluated by Horner and recombined with x^offset */
void oracle_par_fr_eval_poly(u64 out[4], const u64 *coeffs, size_t n, const u64 x[4]) {
    oracle_init();
    const size_t CH = 1 << 14;
    const size_t nch = (n + CH - 1) / CH;
    if (nch <= 1) { oracle_fr_eval_poly(out, coeffs, n, x); return; }
    u64 *part = (u64 *)malloc(32 * nch);
    PAR_FOR for (size_t c = 0; c < nch; c++) {
        size_t lo = c * CH, hi = lo + CH > n ? n : lo + CH;
        u64 v[4], e[4] = {lo, 0, 0, 0}, xp[4];
        oracle_fr_eval_poly(v, coeffs + 4 * lo, hi - lo, x);
        f_pow(&FR, xp, x, e);
        f_mul(&FR, part + 4 * c, v, xp);
    }
    u64 acc[4] = {0, 0, 0, 0};
    for (size_t c = 0; c < nch; c++) f_add(&FR, acc, acc, part + 4 * c);
    memcpy(out, acc, 32);
    free(part);
}

/* ---- Evaluator::evaluate_h as a row-parallel register program (halo2 GraphEvaluator) ----
 * code: n_ins x 4 int32 (op, dst, a, b).  ops: 0 CONST dst <- consts[a]; 1 COL dst <- cols[a][(row + b * rot_scale) mod n];
 * 2 ADD; 3 SUB; 4 MUL; 5 NEG.  Registers are per-thread (halo2's `intermediates`).  out[row] = reg[out_reg]. */
void oracle_expr_eval(u64 *out, size_t n_rows, const u64 *const *cols, const int32_t *code, size_t n_ins, const u64 *consts, int32_t n_regs, int32_t out_reg,
                      int64_t rot_scale) {
    oracle_init();
#pragma omp parallel num_threads(g_threads)
    {
        u64 *reg = (u64 *)calloc((size_t)(n_regs > 0 ? n_regs : 1), 32);
#pragma omp for schedule(static)
        for (size_t row = 0; row < n_rows; row++) {
            for (size_t i = 0; i < n_ins; i++) {
                const int32_t *ins = code + 4 * i;
                u64 *d = reg + 4 * (size_t)ins[1];
                switch (ins[0]) {
                    case 0: memcpy(d, consts + 4 * (size_t)ins[2], 32); break;
                    case 1: {
                        int64_t r = ((int64_t)row + (int64_t)ins[3] * rot_scale) % (int64_t)n_rows;
                        if (r < 0) r += (int64_t)n_rows;
                        memcpy(d, cols[ins[2]] + 4 * (size_t)r, 32);
                        break;
                    }
                    case 2: f_add(&FR, d, reg + 4 * (size_t)ins[2], reg + 4 * (size_t)ins[3]); break;
                    case 3: f_sub(&FR, d, reg + 4 * (size_t)ins[2], reg + 4 * (size_t)ins[3]); break;
                    case 4: f_mul(&FR, d, reg + 4 * (size_t)ins[2], reg + 4 * (size_t)ins[3]); break;
                    default: { u64 z[4] = {0, 0, 0, 0}; f_sub(&FR, d, z, reg + 4 * (size_t)ins[2]); break; }
                }
            }
            memcpy(out + 4 * row, reg + 4 * (size_t)out_reg, 32);
        }
        free(reg);
    }
}

/* ---- lookup argument: halo2 `permute_expression_pair`.  Inputs Montgomery; returns 0, or 1 when an input value is not in the table. */
static int cmp256(const void *pa, const void *pb) {
    const u64 *a = (const u64 *)pa, *b = (const u64 *)pb;
    for (int i = 3; i >= 0; i--) { if (a[i] < b[i]) return -1; if (a[i] > b[i]) return 1; }
    return 0;
}
int oracle_permute_expression_pair(u64 *p_in, u64 *p_tab, const u64 *in, const u64 *tab, size_t usable) {
    oracle_init();
    u64 *a = (u64 *)malloc(32 * (usable ? usable : 1)), *t = (u64 *)malloc(32 * (usable ? usable : 1));
    PAR_FOR for (size_t i = 0; i < usable; i++) { f_to_canon(&FR, a + 4 * i, in + 4 * i); f_to_canon(&FR, t + 4 * i, tab + 4 * i); }
    qsort(a, usable, 32, cmp256);   /* permuted input = sorted input (Fr `Ord` = canonical value) */
    qsort(t, usable, 32, cmp256);   /* the BTreeMap of leftover table values, iterated in ascending order */
    /* first occurrence of each input value takes that value from the table multiset; repeated rows are filled afterwards from the
     * leftover table values in ascending order, popping rows from the END of the repeated-row list */
    size_t *repeated = (size_t *)malloc(sizeof(size_t) * (usable ? usable : 1));
    uint8_t *taken = (uint8_t *)calloc(usable ? usable : 1, 1);
    size_t n_rep = 0, tp = 0;
    int bad = 0;
    for (size_t row = 0; row < usable && !bad; row++) {
        if (row == 0 || cmp256(a + 4 * row, a + 4 * (row - 1)) != 0) {
            while (tp < usable && cmp256(t + 4 * tp, a + 4 * row) < 0) tp++;
            if (tp >= usable || cmp256(t + 4 * tp, a + 4 * row) != 0) { bad = 1; break; }
            taken[tp] = 1;
            memcpy(p_tab + 4 * row, a + 4 * row, 32);
            tp++;
        } else {
            repeated[n_rep++] = row;
        }
    }
    if (!bad) {
        for (size_t i = 0; i < usable; i++) {
            if (taken[i]) continue;
            if (!n_rep) { bad = 2; break; }
            memcpy(p_tab + 4 * repeated[--n_rep], t + 4 * i, 32);
        }
        if (n_rep) bad = 2;
    }
    if (!bad) {
        memcpy(p_in, a, 32 * usable);
        PAR_FOR for (size_t i = 0; i < usable; i++) { f_from_canon(&FR, p_in + 4 * i, p_in + 4 * i); f_from_canon(&FR, p_tab + 4 * i, p_tab + 4 * i); }
    }
    free(a); free(t); free(repeated); free(taken);
    return bad;
}

/* ---- rand_chacha 0.3.1 ChaCha20Rng + halo2curves `Fr::random` (oracle/chacha.py is the readable twin) ---- */
#define ROTL32(x, n) (((x) << (n)) | ((x) >> (32 - (n))))
#define CQR(a, b, c, d) a += b; d ^= a; d = ROTL32(d, 16); c += d; b ^= c; b = ROTL32(b, 12); a += b; d ^= a; d = ROTL32(d, 8); c += d; b ^= c; b = ROTL32(b, 7);
static void chacha20_block(const uint32_t key[8], u64 counter, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32), 0, 0};
    uint32_t s[16];
    memcpy(s, in, 64);
    for (int i = 0; i < 10; i++) {
        CQR(s[0], s[4], s[8], s[12]) CQR(s[1], s[5], s[9], s[13]) CQR(s[2], s[6], s[10], s[14]) CQR(s[3], s[7], s[11], s[15])
        CQR(s[0], s[5], s[10], s[15]) CQR(s[1], s[6], s[11], s[12]) CQR(s[2], s[7], s[8], s[13]) CQR(s[3], s[4], s[9], s[14])
    }
    for (int i = 0; i < 16; i++) out[i] = s[i] + in[i];
}
/* 512-bit little-endian integer (8 x u64) mod r, to Montgomery: lo * R^2 * R^-1 + hi * R^3 * R^-1 ... = lo*R + hi*2^256*R */
static void fr_from_u512(u64 r[4], const u64 w[8]) {
    u64 lo[4], hi[4], r3[4];
    /* w_lo, w_hi may exceed the modulus: f_mul tolerates any 256-bit inputs whose product / R stays < 2^256 * something; reduce first by
     * conditional subtraction (values < 2^256 < 6r) */
    memcpy(lo, w, 32); memcpy(hi, w + 4, 32);
    while (geq(lo, FR.m)) sub_nc(lo, lo, FR.m);
    while (geq(hi, FR.m)) sub_nc(hi, hi, FR.m);
    f_mul(&FR, r3, FR.r2, FR.r2);          /* R^3 */
    f_mul(&FR, lo, lo, FR.r2);             /* lo * R */
    f_mul(&FR, hi, hi, r3);                /* hi * R^2 = (hi * 2^256) * R */
    f_add(&FR, r, lo, hi);
}
/* out[i] = Fr::random drawn from keystream blocks starting at `block0` (one Fr = 8 x next_u64 = exactly one 64-byte block) */
void oracle_chacha_fr_fill(u64 *out, const uint32_t key[8], u64 block0, size_t n) {
    oracle_init();
    PAR_FOR for (size_t i = 0; i < n; i++) {
        uint32_t b[16]; u64 w[8];
        chacha20_block(key, block0 + i, b);
        for (int j = 0; j < 8; j++) w[j] = (u64)b[2 * j] | ((u64)b[2 * j + 1] << 32);
        fr_from_u512(out + 4 * i, w);
    }
}

/* ---- Keccak-256 (Ethereum padding), for Entry::new's hashed username (entry.rs:21) ---- */
static const u64 KRC[24] = {0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
    0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
    0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static void keccak_f(u64 st[25]) {
    static const int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    static const int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    for (int round = 0; round < 24; round++) {
        u64 bc[5], t;
        for (int i = 0; i < 5; i++) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; i++) { t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63)); for (int j = 0; j < 25; j += 5) st[j + i] ^= t; }
        t = st[1];
        for (int i = 0; i < 24; i++) { int j = piln[i]; u64 b = st[j]; st[j] = (t << rotc[i]) | (t >> (64 - rotc[i])); t = b; }
        for (int j = 0; j < 25; j += 5) { for (int i = 0; i < 5; i++) bc[i] = st[j + i]; for (int i = 0; i < 5; i++) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5]; }
        st[0] ^= KRC[round];
    }
}
void oracle_keccak256(uint8_t out[32], const uint8_t *data, size_t len) {
    u64 st[25]; memset(st, 0, sizeof st);
    const size_t rate = 136;
    while (len >= rate) { for (size_t i = 0; i < rate / 8; i++) { u64 w; memcpy(&w, data + 8 * i, 8); st[i] ^= w; } keccak_f(st); data += rate; len -= rate; }
    uint8_t blk[136]; memset(blk, 0, sizeof blk); memcpy(blk, data, len);
    blk[len] ^= 0x01; blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; i++) { u64 w; memcpy(&w, blk + 8 * i, 8); st[i] ^= w; }
    keccak_f(st);
    memcpy(out, st, 32);
}

/* ---- Poseidon (WIDTH 2, RATE 1, R_F 8, R_P 56, x^5) and the Merkle sum tree ---- */
static u64 P_RC[64][2][4], P_MDS[2][2][4];
static int g_poseidon = 0;
/* rc: 64 x 2 field elements, mds: 2 x 2, all Montgomery (tests/golden/poseidon_params.json, i.e. chips/poseidon/poseidon_params.rs) */
void oracle_poseidon_set_params(const u64 *rc, const u64 *mds) { oracle_init(); memcpy(P_RC, rc, sizeof P_RC); memcpy(P_MDS, mds, sizeof P_MDS); g_poseidon = 1; }
static inline void pow5(u64 r[4], const u64 a[4]) { u64 a2[4], a4[4]; f_sqr(&FR, a2, a); f_sqr(&FR, a4, a2); f_mul(&FR, r, a4, a); }
static void poseidon_permute(u64 s[2][4]) {
    for (int r = 0; r < 64; r++) {
        u64 t0[4], t1[4], a[4], b[4];
        f_add(&FR, t0, s[0], P_RC[r][0]); f_add(&FR, t1, s[1], P_RC[r][1]);
        pow5(t0, t0);
        if (r < 4 || r >= 60) pow5(t1, t1);
        f_mul(&FR, a, P_MDS[0][0], t0); f_mul(&FR, b, P_MDS[0][1], t1); f_add(&FR, s[0], a, b);
        f_mul(&FR, a, P_MDS[1][0], t0); f_mul(&FR, b, P_MDS[1][1], t1); f_add(&FR, s[1], a, b);
    }
}
/* ConstantLength<L>: state [0, L * 2^64], absorb one element per permutation, output state[0] */
static void poseidon_hash(u64 out[4], const u64 *inputs, size_t L) {
    u64 s[2][4], c[4] = {0, (u64)L, 0, 0};
    memset(s[0], 0, 32);
    f_from_canon(&FR, s[1], c);
    for (size_t i = 0; i < L; i++) { f_add(&FR, s[0], s[0], inputs + 4 * i); poseidon_permute(s); }
    memcpy(out, s[0], 32);
}
void oracle_poseidon_hash(u64 out[4], const u64 *inputs, size_t L) { oracle_init(); poseidon_hash(out, inputs, L); }

/* Entry::new + compute_leaf for every entry (usernames concatenated, offsets[n + 1]; balances n x n_cur u64), zero entries pad to 2^depth
 * (mst.rs:106-120); then build_merkle_tree_from_leaves (build_tree.rs:5-78).  Layout of the result: hashes[level][index] flat with level
 * offsets 0, 2^depth, 2^depth + 2^(depth-1), ...; balances likewise x n_cur; unames: 2^depth hashed usernames (mod r). */
void oracle_mst_build(u64 *hashes, u64 *balances, u64 *unames, const uint8_t *names, const uint32_t *offsets, const u64 *bal64, size_t n_entries, uint32_t n_cur, uint32_t depth) {
    oracle_init();
    const size_t leaves = (size_t)1 << depth;
    PAR_FOR for (size_t i = 0; i < leaves; i++) {
        u64 pre[4 * 34];
        memset(pre, 0, 32 * (size_t)(n_cur + 1));
        if (i < n_entries) {
            uint8_t h[32], le[32];
            oracle_keccak256(h, names + offsets[i], offsets[i + 1] - offsets[i]);
            for (int b = 0; b < 32; b++) le[b] = h[31 - b];   /* BigUint::from_bytes_be */
            u64 w[4]; memcpy(w, le, 32);
            while (geq(w, FR.m)) sub_nc(w, w, FR.m);
            f_from_canon(&FR, pre, w);
            for (uint32_t c = 0; c < n_cur; c++) { u64 v[4] = {bal64[i * n_cur + c], 0, 0, 0}; f_from_canon(&FR, pre + 4 * (c + 1), v); }
        }
        memcpy(unames + 4 * i, pre, 32);
        for (uint32_t c = 0; c < n_cur; c++) memcpy(balances + 4 * (i * n_cur + c), pre + 4 * (c + 1), 32);
        poseidon_hash(hashes + 4 * i, pre, n_cur + 1);
    }
    size_t off = 0;
    for (uint32_t level = 1; level <= depth; level++) {
        const size_t cnt = (size_t)1 << (depth - level), prev = off;
        off += (size_t)1 << (depth - level + 1);
        PAR_FOR for (size_t i = 0; i < cnt; i++) {
            u64 pre[4 * 35];
            for (uint32_t c = 0; c < n_cur; c++) {
                f_add(&FR, pre + 4 * c, balances + 4 * ((prev + 2 * i) * n_cur + c), balances + 4 * ((prev + 2 * i + 1) * n_cur + c));
                memcpy(balances + 4 * ((off + i) * n_cur + c), pre + 4 * c, 32);
            }
            memcpy(pre + 4 * n_cur, hashes + 4 * (prev + 2 * i), 32);
            memcpy(pre + 4 * (n_cur + 1), hashes + 4 * (prev + 2 * i + 1), 32);
            poseidon_hash(hashes + 4 * (off + i), pre, n_cur + 2);
        }
    }
}

/* ---- `ParamsKZG::setup(k, rng)` core (utils.rs:70): out[i] = scalars[i] * G, affine.  8-bit fixed windows over a table of
 * j * 2^(8w) * G (32 x 255 affine points), one mixed addition per non-zero byte, batched normalisation. ---- */
void oracle_g1_fixed_base_mul(u64 *out, const u64 *scalars_mont, size_t n) {
    oracle_init();
    aff_t *table = (aff_t *)malloc(sizeof(aff_t) * 32 * 255);
    {
        aff_t g; memset(&g, 0, sizeof g);
        u64 one_c[4] = {1, 0, 0, 0}, two_c[4] = {2, 0, 0, 0};
        f_from_canon(&FQ, g.x, one_c); f_from_canon(&FQ, g.y, two_c);
        jac_t base; memcpy(base.x, g.x, 32); memcpy(base.y, g.y, 32); memcpy(base.z, FQ.r, 32);
        for (int w = 0; w < 32; w++) {
            jac_t acc = base;
            aff_t base_aff; jac_to_affine(&base_aff, &base);
            for (int j = 1; j <= 255; j++) {
                jac_to_affine(&table[w * 255 + j - 1], &acc);
                jac_add_affine(&acc, &acc, &base_aff);
            }
            base = acc;  /* 256 * previous base */
        }
    }
    const size_t BATCH = 1024;
    const size_t nb = (n + BATCH - 1) / BATCH;
    PAR_FOR for (size_t bi = 0; bi < nb; bi++) {
        size_t lo = bi * BATCH, m = lo + BATCH > n ? n - lo : BATCH;
        jac_t *buf = (jac_t *)malloc(sizeof(jac_t) * BATCH);
        u64 *pref = (u64 *)malloc(32 * BATCH);
        for (size_t i = 0; i < m; i++) {
            u64 c[4]; f_to_canon(&FR, c, scalars_mont + 4 * (lo + i));
            const uint8_t *bytes = (const uint8_t *)c;
            jac_t acc; jac_set_id(&acc);
            for (int w = 0; w < 32; w++) if (bytes[w]) jac_add_affine(&acc, &acc, &table[w * 255 + bytes[w] - 1]);
            buf[i] = acc;
        }
        u64 acc[4]; memcpy(acc, FQ.r, 32);
        for (size_t i = 0; i < m; i++) { memcpy(pref + 4 * i, acc, 32); if (!jac_is_id(&buf[i])) f_mul(&FQ, acc, acc, buf[i].z); }
        u64 inv[4]; f_inv(&FQ, inv, acc);
        for (size_t i = m; i-- > 0;) {
            u64 *o = out + 8 * (lo + i);
            if (jac_is_id(&buf[i])) { memset(o, 0, 64); continue; }
            u64 zi[4], zi2[4], zi3[4];
            f_mul(&FQ, zi, inv, pref + 4 * i); f_mul(&FQ, inv, inv, buf[i].z);
            f_sqr(&FQ, zi2, zi); f_mul(&FQ, zi3, zi2, zi);
            f_mul(&FQ, o, buf[i].x, zi2); f_mul(&FQ, o + 4, buf[i].y, zi3);
        }
        free(buf); free(pref);
    }
    free(table);
}
