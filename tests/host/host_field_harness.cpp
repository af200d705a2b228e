// Host-side harness: compiles the *device* field/curve headers (fp.cuh, ec.cuh) with g++ using the
// bit-exact host emulation of the PTX carry primitives, and exports C entry points for ctypes so
// tests/test_host_field.py can compare them with the CPU oracle without a GPU.
#include "../../circuits_halo2_b200/csrc/ec.cuh"
#include <cstddef>
#include <cstring>
using namespace sb;
template <class F> static void binop(uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n, int op) {
    for (size_t i = 0; i < n; i++) {
        F x, y, z;
        memcpy(x.v, a + 8 * i, 32); memcpy(y.v, b + 8 * i, 32);
        z = op == 0 ? mul(x, y) : op == 1 ? add(x, y) : op == 2 ? sub(x, y) : op == 3 ? inv(x) : op == 4 ? to_mont(x) : op == 6 ? inv_fermat(x) : op == 7 ? sqr(x) : from_mont(x);
        memcpy(r + 8 * i, z.v, 32);
    }
}
extern "C" {
void ht_fr_op(uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n, int op) { binop<fr_t>(r, a, b, n, op); }
void ht_fq_op(uint32_t *r, const uint32_t *a, const uint32_t *b, size_t n, int op) { binop<fq_t>(r, a, b, n, op); }
// out(affine) = sum_i (+/-) pts[i] using madd into one accumulator, then general add of a second
// accumulator built from the second half, to exercise madd / add / dbl / to_affine.
void ht_sum_points(uint32_t *out, const uint32_t *pts, const uint8_t *negs, size_t n) {
    xyzz_t a = xyzz_t::identity(), b = xyzz_t::identity();
    for (size_t i = 0; i < n; i++) {
        affine_t p; memcpy(&p, pts + 16 * i, 64);
        if (i < n / 2) madd(a, p, negs[i] != 0); else madd(b, p, negs[i] != 0);
    }
    add(a, b);
    affine_t r = to_affine(a);
    memcpy(out, &r, 64);
}
void ht_double_xyzz(uint32_t *out, const uint32_t *pt, int times) {
    affine_t p; memcpy(&p, pt, 64);
    xyzz_t a = xyzz_t::from_affine(p);
    for (int i = 0; i < times; i++) a = dbl(a);
    affine_t r = to_affine(a);
    memcpy(out, &r, 64);
}
void ht_add_xyzz(uint32_t *out, const uint32_t *p1, const uint32_t *p2) {
    affine_t p, q; memcpy(&p, p1, 64); memcpy(&q, p2, 64);
    xyzz_t a = xyzz_t::from_affine(p), b = xyzz_t::from_affine(q);
    a = dbl(a); a = dbl(a);            // make ZZ != 1 on both sides
    b = dbl(b);
    xyzz_t a4 = a, b2 = b;
    add(a4, b2);                       // 4p + 2q
    affine_t r = to_affine(a4);
    memcpy(out, &r, 64);
}
}
