// Register-resident radix-8 NTT tile for BN254 Fr (sm_100a): the per-thread phases of `ntt_pass_kernel` (ntt.cu).
//
// Replaces halo2_proofs::arithmetic::best_fft and the scaling passes EvaluationDomain wraps around it.
//
// A pass transforms one digit (r bits, R = 2^r) of the index for a tile of 2^t elements (t = r + g: R rows x G = 2^g independent columns).
// Every thread keeps EIGHT elements in registers and runs up to three decimation-in-frequency levels (a radix-8 butterfly) on them without
// touching memory; between such stages the tile is exchanged through shared memory (XOR-folded 16-byte bank swizzle: conflict-free for every
// power-of-two stride).  So an element crosses shared memory ceil(r / 3) - 1 times per pass instead of 2 r times, and there are two barriers per
// exchange instead of one per level.  Tile index idx (t bits) of element (j, gg):  strided passes idx = j * G + gg (gg fastest: G x 32 B
// contiguous in HBM);  the last pass of a multi-pass plan idx = gg * R + j (its rows are contiguous in HBM and arrive through cp.async).
// In stage s the thread's eight elements differ in the 3 idx bits [p, p + 3), p = max(jshift, jshift + r - 3 (s + 1)); levels whose j-bit was
// already processed by the previous stage (r not a multiple of 3) are skipped.
//
// Everything here is __host__ __device__ and free of CUDA built-ins, so tests/host/host_ntt_harness.cpp runs the very same code for every
// thread of a tile on the CPU (phase by phase, barriers = loop boundaries) and tests/test_host_logic.py compares it with the oracle's best_fft.
#pragma once
#include "fp.cuh"

namespace sb {

static const int NTT_MAX_PASS = 8;

enum NttKind { NTT_STRIDED = 0, NTT_LAST = 1, NTT_SINGLE = 2 };

struct NttPassArgs {
    const uint4 *src;
    uint4 *dst;
    const uint4 *w;        // butterfly twiddles omega_R^e, e < R / 2 (global, natural order)
    const uint4 *tw_full;  // inter-pass twiddles of this pass laid out like its output inside one a-block: [(k << log_c) + c] = s * omega^(A c k); null: two-level
    const uint4 *t_lo;     // two-level fallback: omega^i, i < 2^log_tlo
    const uint4 *t_hi;     //                     omega^(i << log_tlo)
    const uint4 *pre_vec;  // optional, first pass: input element i is multiplied by pre_vec[i]
    const uint4 *post_vec; // optional, last pass: output element i is multiplied by post_vec[i]
    uint32_t log_n, r, g, log_a, log_c, log_r1, log_tlo;
    uint32_t kind, npass, n_mid;
    uint32_t eb;           // log2 of the elements a thread keeps in registers: 3 (radix-8 stages) or 2 (radix-4 stages, twice the threads per tile)
    uint32_t tile0;        // first tile of this launch (a rank of a distributed transform runs a sub-range of a pass' tiles)
    uint32_t mid_bits[NTT_MAX_PASS];  // radices of passes 2 .. P-1 (for the last pass' digit reversal)
    uint32_t pre_m;    // optional, first pass: input element i is multiplied by pre_pat[i % pre_m] (pre_m <= 8; 0 = off)
    uint32_t post_m;   // optional, last pass: output element i is multiplied by post_pat[i % post_m] (post_m <= 8; 0 = off)
    uint64_t n_in;     // first pass: input elements at index >= n_in read as zero (zero-padded polynomial)
    uint64_t n_out;    // last pass: output elements at index >= n_out are not stored (truncation)
    // batched launch (gridDim.y transforms at once): element strides between consecutive transforms of src / dst / pre_vec / post_vec
    uint64_t src_bstride, dst_bstride, pre_bstride, post_bstride;
    uint32_t has_tw_scale;  // two-level fallback, first pass: the plan's scale (n^-1 of an inverse transform) multiplies the twiddle
    Fp<FrParams> tw_scale;
    Fp<FrParams> pre_pat[8], post_pat[8];
};

// 16-byte bank-group swizzle: low three bits XOR-folded with every higher 3-bit group (conflict-free for consecutive and for 2^s-strided accesses)
SB_HD uint32_t ntt_swz(uint32_t i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9) ^ (i >> 12)) & 7u); }

SB_HD uint32_t ntt_brev(uint32_t x, uint32_t bits) {
#ifdef __CUDA_ARCH__
    return bits ? (__brev(x) >> (32 - bits)) : 0u;
#endif
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// geometry of one pass, derived once per thread
struct NttGeom {
    uint32_t r, g, t, jshift;  // j sits at idx bits [jshift, jshift + r); gg at [0, g) (strided) or [r, r + g) (last)
    uint32_t n_stages, eb;
    SB_HD NttGeom(const NttPassArgs &p) {
        r = p.r; g = p.g; t = r + g;
        eb = p.eb ? p.eb : 3;
        jshift = (p.kind == NTT_STRIDED) ? g : 0;
        n_stages = (r + eb - 1) / eb;
    }
    // lowest idx bit of the register window of stage s
    SB_HD uint32_t window(uint32_t s) const {
        const int w = (int)jshift + (int)r - (int)eb * ((int)s + 1);
        return (uint32_t)(w < (int)jshift ? (int)jshift : w);
    }
    // tile index of register b of thread tid in a stage whose window starts at p
    SB_HD uint32_t idx(uint32_t tid, uint32_t p, uint32_t b) const { return ((tid >> p) << (p + eb)) | (b << p) | (tid & ((1u << p) - 1u)); }
    SB_HD uint32_t j_of(uint32_t idx_) const { return (idx_ >> jshift) & ((1u << r) - 1u); }
    SB_HD uint32_t gg_of(uint32_t idx_) const { return jshift ? (idx_ & ((1u << g) - 1u)) : (idx_ >> r); }
};

// global element index of tile element (j, gg) on the INPUT side of the pass
struct NttTileCoord {
    uint64_t base;   // strided / single
    uint32_t c0;     // first column (strided)
    uint64_t k1_0, rest, rev;
    uint32_t log_rest;
    SB_HD NttTileCoord(const NttPassArgs &p, uint64_t tile_id) {
        base = 0; c0 = 0; k1_0 = 0; rest = 0; rev = 0; log_rest = 0;
        if (p.kind == NTT_STRIDED) {
            const uint32_t log_cg = p.log_c - p.g;
            const uint64_t a_idx = tile_id >> log_cg;
            c0 = (uint32_t)(tile_id & ((1ull << log_cg) - 1)) << p.g;
            base = (a_idx << (p.r + p.log_c)) + c0;
        } else if (p.kind == NTT_LAST) {
            log_rest = p.log_a - p.log_r1;
            rest = tile_id & ((1ull << log_rest) - 1);
            k1_0 = (tile_id >> log_rest) << p.g;
            // digit-reverse rest = (k_2 .. k_{P-1}), most significant first -> k_2 + R_2 k_3 + ...
            uint64_t tmp = rest;
            for (int m = (int)p.n_mid - 1; m >= 0; m--) {
                const uint32_t bits = p.mid_bits[m];
                const uint64_t d = tmp & ((1ull << bits) - 1);
                tmp >>= bits;
                uint32_t shift = 0;
                for (int q = 0; q < m; q++) shift += p.mid_bits[q];
                rev |= d << shift;
            }
        }
    }
    SB_HD uint64_t in_index(const NttPassArgs &p, uint32_t j, uint32_t gg) const {
        if (p.kind == NTT_STRIDED) return base + ((uint64_t)j << p.log_c) + gg;
        if (p.kind == NTT_LAST) return ((((k1_0 + gg) << log_rest) + rest) << p.r) + j;
        return j;
    }
    // where output digit k of column gg goes
    SB_HD uint64_t out_index(const NttPassArgs &p, uint32_t k, uint32_t gg) const {
        if (p.kind == NTT_STRIDED) return base + ((uint64_t)k << p.log_c) + gg;
        if (p.kind == NTT_LAST) return (k1_0 + gg) + ((rev + ((uint64_t)k << (p.log_a - p.log_r1))) << p.log_r1);
        return k;
    }
};

typedef Fp<FrParams> nfr_t;

// ONE shared copy of the Montgomery product on the device (operands by value = in registers): inlined, the ~120 products of a pass kernel
// are 425 KB of SASS and the instruction cache becomes the bottleneck (the same lesson as msm.cu's ec_mul)
#ifdef __CUDACC__
static __device__ __noinline__ nfr_t ntt_mul_dev(nfr_t a, nfr_t b) { return mul(a, b); }
#endif
SB_HD nfr_t ntt_mul(const nfr_t &a, const nfr_t &b) {
#ifdef __CUDA_ARCH__
    return ntt_mul_dev(a, b);
#else
    return mul(a, b);
#endif
}

// ---- the DIF levels of one stage on the eight registers -------------------------------------------------------------------------------
// TW: functor e -> omega_R^e (e < R / 2).  `low` = first j-bit NOT yet processed by earlier stages (levels at j-bits >= low are skipped).
template <int EB, class TW>
SB_HD void ntt_stage_butterflies(nfr_t x[1 << EB], const NttGeom &G, uint32_t tid, uint32_t p, uint32_t low, const TW &tw) {
    const uint32_t jb0 = p - G.jshift;                          // j-bit of window bit 0
    const uint32_t tl = (tid & ((1u << p) - 1u)) >> G.jshift;   // the thread's j bits below the window
#pragma unroll
    for (int wb = EB - 1; wb >= 0; wb--) {
        const uint32_t q = jb0 + (uint32_t)wb;                  // j-bit of this level: pairs differ in it, h = 2^q
        if (q >= low || q >= G.r) continue;
        const uint32_t sh = G.r - 1 - q;
#pragma unroll
        for (int b = 0; b < (1 << EB); b++) {
            if (b & (1 << wb)) continue;
            const int b1 = b | (1 << wb);
            const uint32_t jlow = (((uint32_t)b & ((1u << wb) - 1u)) << jb0) | tl;   // j mod 2^q
            const nfr_t u = x[b], v = x[b1];
            x[b] = add(u, v);
            nfr_t d = sub(u, v);
            if (q != 0) d = ntt_mul(d, tw(jlow << sh));              // the last level's twiddle is omega^0
            x[b1] = d;
        }
    }
}

// inter-pass twiddle (strided passes) for output row k, column c of the tile's a-block
SB_HD nfr_t ntt_load_fr(const uint4 *p) {
    nfr_t r;
    const uint4 a = p[0], b = p[1];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
SB_HD void ntt_store_fr(uint4 *p, const nfr_t &x) {
    uint4 a, b;
    a.x = x.v[0]; a.y = x.v[1]; a.z = x.v[2]; a.w = x.v[3];
    b.x = x.v[4]; b.y = x.v[5]; b.z = x.v[6]; b.w = x.v[7];
    p[0] = a; p[1] = b;
}
SB_HD nfr_t ntt_zero() { return nfr_t::zero(); }

// element offsets of transform `bi` of a batched launch
struct NttBatch {
    uint64_t src, dst, pre, post;
    SB_HD NttBatch(const NttPassArgs &p, uint32_t bi) : src(bi * p.src_bstride), dst(bi * p.dst_bstride), pre(bi * p.pre_bstride), post(bi * p.post_bstride) {}
};

// input element `gi` of the first pass with the optional fused pre-operations (zero padding, vector / pattern scaling)
SB_HD nfr_t ntt_fetch_input(const NttPassArgs &p, const NttBatch &bo, uint64_t gi) {
    if (gi >= p.n_in) return ntt_zero();
    nfr_t x = ntt_load_fr(p.src + 2 * (bo.src + gi));
    if (p.pre_vec) x = ntt_mul(x, ntt_load_fr(p.pre_vec + 2 * (bo.pre + gi)));
    if (p.pre_m) x = ntt_mul(x, p.pre_pat[(uint32_t)gi % p.pre_m]);
    return x;
}

// finishing touch and store of output digit k (already bit-reversed) of column gg
SB_HD void ntt_emit(const NttPassArgs &p, const NttBatch &bo, const NttTileCoord &tc, uint32_t k, uint32_t gg, nfr_t x) {
    const uint64_t o = tc.out_index(p, k, gg);
    if (p.kind == NTT_STRIDED) {
        const uint32_t c = tc.c0 + gg;
        if (p.tw_full) {
            x = ntt_mul(x, ntt_load_fr(p.tw_full + 2 * (((uint64_t)k << p.log_c) + c)));
        } else {
            const uint64_t E = ((uint64_t)c * k) << p.log_a;  // < N
            nfr_t tw = ntt_load_fr(p.t_lo + 2 * (E & ((1ull << p.log_tlo) - 1)));
            const uint64_t eh = E >> p.log_tlo;
            if (eh) tw = ntt_mul(tw, ntt_load_fr(p.t_hi + 2 * eh));
            if (p.has_tw_scale) tw = ntt_mul(tw, p.tw_scale);
            x = ntt_mul(x, tw);
        }
    } else {
        if (o >= p.n_out) return;
        if (p.post_m) x = ntt_mul(x, p.post_pat[(uint32_t)o % p.post_m]);
        if (p.post_vec) x = ntt_mul(x, ntt_load_fr(p.post_vec + 2 * (bo.post + o)));
    }
    ntt_store_fr(p.dst + 2 * (bo.dst + o), x);
}

}  // namespace sb
