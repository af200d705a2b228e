// Pippenger MSM over BN254 G1 for sm_100a.
//
// Replaces halo2_proofs::arithmetic::best_multiexp (reached from
// zk_prover/src/circuits/utils.rs:75-76 (keygen, 17 calls) and :94-102,171-178 (create_proof,
// 16 calls) through ParamsKZG::{commit, commit_lagrange}).  The result is a unique group element,
// so the GPU algorithm is free; it is chosen for B200:
//
//   1. recode   every scalar -> W = ceil(255/c) signed c-bit digits; key = window * 2^(c-1) + |d| - 1
//   2. sort     counting sort of the (key, point index | sign) pairs: histogram with warp-aggregated
//               atomics, exclusive scan, scatter with warp-aggregated cursor claims (order inside a
//               bucket is irrelevant, so no stable multi-pass radix sort is needed)
//   3. reduce-by-key, perfectly load-balanced regardless of the scalar distribution:
//               every thread owns a fixed-size chunk of the sorted list (not a bucket), walks it with
//               one XYZZ accumulator (mixed addition, 8M+2S), stores runs that lie strictly inside
//               its chunk straight into their bucket and hands the first / last run of the chunk to
//               the next level as a (key, partial sum) pair.  Levels shrink 8x; a last single-CTA
//               segmented scan finishes.  A bucket holding 90 % of all points (constant columns,
//               selector columns, Z polynomials) costs exactly the same as a uniform one.
//   4. bucket reduction  sum_b b * B_b per window, in segments of S buckets (running sums + one small
//               double-and-add), then a CTA tree per window
//   5. the <= 64 window sums go to the host, which folds them (254 dependent doublings) and
//               normalises (host_g1.cpp).
// Control logic is pinned on the CPU by tests/models/msm_model.py.
#include <stdlib.h>

#ifndef SB_EC_INLINE_MUL
#define SB_EC_NOINLINE_MUL 1
#endif
#include "common.cuh"
#include "ec.cuh"

namespace sb {

void host_fold_windows(const uint8_t *win, int n_windows, int c, uint8_t out_affine[64]);
void host_bucket_combine(const uint8_t *fin, int n_bits, int shift, int x_slot, uint8_t out_xyzz[128]);
void host_residue_fixup(uint8_t r_xyzz[128], const uint8_t total_xyzz[128], int log_mod, int res);

static const uint32_t INVALID_KEY = 0xffffffffu;
// chunk length of reduce levels >= 2: these levels are chains of dependent EC additions on few threads, so short chunks (more, smaller
// levels) finish sooner than long ones; 8 measured best from 2^13 to 2^22 points (SB_MSM_LK overrides)
#define LK (ctx->tune.msm_lk)
static const int FINAL_MAX = 256;  // slots handled by the final single-CTA level

struct MsmShape {
    uint32_t c, W, B;   // window bits, windows, buckets per window
    uint32_t L1;        // chunk length of level 1
    uint32_t seg_log;   // bucket-reduction segment = 2^seg_log buckets
    uint64_t n, t_max;  // points, sorted-list capacity (W * n rounded up to 4)
    uint32_t w_lo, W_all;  // this launch set covers windows [w_lo, w_lo + W) of the W_all windows of the scalar (window-sharded MSM)
    uint32_t Wb;           // bucket sets: W, or 1 when the bases come from fixed-base window tables (all windows share one set)
    uint32_t batch;        // scalar vectors sharing the bases (tables only): vector j = scalars[j n .. (j+1) n), bucket set j
    uint32_t piece_off[8]; // table-entry offset of vector j (0, or the distance to the other basis' tables inside one slab)
    uint32_t tab_stride;   // 0, or the table stride: the point for (window w, base i) is tables[w * tab_stride + i] = 2^(c w) * P_i
    uint32_t res, log_mod; // residue shard (multi-GPU, tables): only digits with (|d| - 1) mod 2^log_mod == res are kept, in bucket (|d| - 1) >> log_mod; B is the reduced count
};

// ------------------------------------------------------------------ 1+2: recode, histogram, scatter
// digit w of canonical scalar s (8 x u32), with the running carry of the signed recoding
__device__ __forceinline__ uint32_t raw_window(const uint32_t s[8], uint32_t bit, uint32_t c) {
    const uint32_t limb = bit >> 5, off = bit & 31;
    uint64_t two = s[limb];
    if (limb + 1 < 8) two |= (uint64_t)s[limb + 1] << 32;
    return (uint32_t)(two >> off) & ((1u << c) - 1);
}

template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_sort_kernel(const uint4 *scalars, uint64_t n, MsmShape sh, uint32_t *counts /*histogram or cursor*/,
                                                       uint32_t *svals) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // warp-uniform trip count so that every lane reaches the shuffles / ballots below
    const uint64_t n_all = n * sh.batch;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n_all; base += stride) {
        const uint64_t gi = base + lane;
        const bool live = gi < n_all;
        const uint32_t piece = (uint32_t)(gi / n);
        const uint64_t i = gi - (uint64_t)piece * n;
        uint32_t s[8];
#pragma unroll
        for (int q = 0; q < 8; q++) s[q] = 0;
        bool nonzero = false;
        if (live) {
            const fr_t raw = load_fp<FrParams>(scalars + 2 * gi);
            nonzero = !raw.is_zero();  // Montgomery form of zero is zero
            if (nonzero) {
                const fr_t x = from_mont(raw);
#pragma unroll
                for (int q = 0; q < 8; q++) s[q] = x.v[q];
            }
        }
        // witness and lookup columns are almost entirely zero: a warp of zero scalars has no digits to count or scatter
        if (!__any_sync(0xffffffffu, nonzero)) continue;
        uint32_t carry = 0;
        const uint32_t half = 1u << (sh.c - 1);
        for (uint32_t w = 0; w < sh.w_lo + sh.W; w++) {
            const uint32_t bit = w * sh.c;
            uint32_t d = (bit < 256 ? raw_window(s, bit, sh.c) : 0) + carry;
            uint32_t neg = 0;
            if (d > half) {
                d = (1u << sh.c) - d;
                neg = 1;
                carry = 1;
            } else {
                carry = 0;
            }
            if (w < sh.w_lo) continue;  // warp-uniform: earlier windows only feed the carry
            uint32_t key = INVALID_KEY;
            if (live && d) {
                const uint32_t dm = d - 1;
                if ((dm & ((1u << sh.log_mod) - 1u)) == sh.res) key = (sh.tab_stride ? piece * sh.B : (w - sh.w_lo) * sh.B) + (dm >> sh.log_mod);
            }
            // warp aggregation by RUNS of equal keys in neighbouring lanes: one atomic per run.  Equal scalars sit in neighbouring rows (constant and
            // selector-like columns, the sorted permuted lookup columns), so runs catch the hot buckets; match.any would also catch scattered
            // duplicates but costs ~70 cycles of a shared unit per digit (ncu: 89 % busy, profiles/r02l), twice the whole histogram otherwise.
            const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
            const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || prev != key);
            const uint32_t leader = 31u - __clz(heads & (0xffffffffu >> (31u - lane)));       // head of my run
            const uint32_t after = heads & ~(0xffffffffu >> (31u - lane));                    // heads above my lane
            const uint32_t run_end = after ? (uint32_t)__ffs(after) - 1u : 32u;               // first lane of the next run
            uint32_t pos = 0;
            if (lane == leader && key != INVALID_KEY) pos = atomicAdd(counts + key, run_end - leader);
            if (SCATTER) {
                pos = __shfl_sync(0xffffffffu, pos, leader);
                if (key != INVALID_KEY) {
                    pos += lane - leader;
                    svals[pos] = (uint32_t)(sh.tab_stride ? sh.piece_off[piece & 7] + w * sh.tab_stride + i : i) | (neg << 31);
                }
            }
        }
    }
}

// exclusive scan of m counts (three small kernels; m <= 2^25)
static const int SCAN_ITEMS = 8, SCAN_THREADS = 256, SCAN_TILE = SCAN_ITEMS * SCAN_THREADS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
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
        warp_sums[lane] = wi - ws;  // exclusive prefix of warp sums
        if (lane == 31) *total = wi;
    }
    __syncthreads();
    uint32_t r = incl - v + warp_sums[wid];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const uint32_t *in, uint64_t m, uint32_t *tile_sums) {
    __shared__ uint32_t total;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++)
        if (base + q < m) s += in[base + q];
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_of_tile_sums_kernel(uint32_t *tile_sums, uint64_t ntiles, uint32_t *grand_total) {
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (uint64_t base = 0; base < ntiles; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < ntiles) tile_sums[i] = ex + carry;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t *in, uint64_t m, const uint32_t *tile_sums, uint32_t *out_a, uint32_t *out_b) {
    __shared__ uint32_t total;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
        v[q] = base + q < m ? in[base + q] : 0;
        s += v[q];
    }
    uint32_t ex = block_exclusive_scan(s, &total) + tile_sums[blockIdx.x];
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; q++) {
        if (base + q < m) {
            out_a[base + q] = ex;
            out_b[base + q] = ex;
        }
        ex += v[q];
    }
}

// ------------------------------------------------------------------ 3: chunked reduce-by-key
__device__ __forceinline__ xyzz_t load_xyzz(const uint4 *p) {
    xyzz_t r;
    r.x = load_fp<FqParams>(p);
    r.y = load_fp<FqParams>(p + 2);
    r.zz = load_fp<FqParams>(p + 4);
    r.zzz = load_fp<FqParams>(p + 6);
    return r;
}
__device__ __forceinline__ void store_xyzz(uint4 *p, const xyzz_t &a) {
    store_fp(p, a.x);
    store_fp(p + 2, a.y);
    store_fp(p + 4, a.zz);
    store_fp(p + 6, a.zzz);
}

// Level 1: the sorted list holds only (base index | sign); the bucket of position `pos` follows from the exclusive offsets of the
// counting sort (offsets[b] <= pos < offsets[b + 1]), so the scatter writes 4 bytes per digit instead of 8 and nothing reads keys
// back.  A thread finds the bucket of its first entry by binary search and re-searches whenever it walks off the current bucket
// (a linear walk could cross arbitrarily many empty buckets).  offsets[nb] is the number of valid digits.
__device__ __forceinline__ uint32_t bucket_of(const uint32_t *offsets, uint32_t nb, uint32_t pos) {
    uint32_t lo = 0, hi = nb;  // largest b in [0, nb) with offsets[b] <= pos
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= pos) lo = mid;
        else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(128) msm_reduce_first_kernel(const uint32_t *offsets, uint32_t nb, const uint32_t *vals, const uint4 *bases, uint32_t L, uint4 *buckets,
                                                               uint32_t *keys_out, uint4 *pts_out, uint64_t nchunks) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchunks) return;
    const uint32_t n_valid = __ldg(offsets + nb);
    const uint64_t start64 = t * L;
    uint32_t k0 = INVALID_KEY, k1 = INVALID_KEY;  // partial slots of this chunk
    if (start64 < n_valid) {
        const uint32_t start = (uint32_t)start64;
        const uint32_t end = min(n_valid, start + L);
        uint32_t cur = bucket_of(offsets, nb, start);
        uint32_t next_off = __ldg(offsets + cur + 1);
        uint32_t nruns = 1;
        xyzz_t acc = xyzz_t::identity();
        for (uint32_t pos = start; pos < end; pos += 4) {
            const uint4 v4 = *reinterpret_cast<const uint4 *>(vals + pos);  // start and L are multiples of 4; the list is padded
            const uint32_t vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (pos + q >= end) break;
                if (pos + q >= next_off) {
                    if (nruns == 1) {
                        k0 = cur;
                        store_xyzz(pts_out + 8 * (2 * t), acc);
                    } else {
                        store_xyzz(buckets + 8 * (uint64_t)cur, acc);  // interior run: sole owner of its bucket
                    }
                    // the next bucket is almost always the neighbour: one load; otherwise (runs of empty buckets) search
                    cur++;
                    next_off = __ldg(offsets + cur + 1);
                    if (pos + q >= next_off) {
                        cur = bucket_of(offsets, nb, pos + q);
                        next_off = __ldg(offsets + cur + 1);
                    }
                    nruns++;
                    acc = xyzz_t::identity();
                }
                const uint32_t v = vv[q];
                const uint4 *bp = bases + 4 * (uint64_t)(v & 0x7fffffffu);
                affine_t p;
                p.x = ldg_fp<FqParams>(bp);
                p.y = ldg_fp<FqParams>(bp + 2);
                madd(acc, p, (v >> 31) != 0);
            }
        }
        if (nruns == 1) {
            k0 = cur;
            store_xyzz(pts_out + 8 * (2 * t), acc);
        } else {
            k1 = cur;
            store_xyzz(pts_out + 8 * (2 * t + 1), acc);
        }
    }
    keys_out[2 * t] = k0;
    keys_out[2 * t + 1] = k1;
}

// Levels >= 2: entries are (key, XYZZ partial) produced by the previous level; INVALID slots are skipped.
__global__ void __launch_bounds__(128) msm_reduce_kernel(const uint32_t *keys, const uint4 *pts_in, uint64_t n_in, uint32_t L, uint4 *buckets, uint32_t *keys_out,
                                                         uint4 *pts_out, uint64_t nchunks) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchunks) return;
    const uint64_t start = t * L;
    uint32_t cur = INVALID_KEY, nruns = 0;
    uint32_t k0 = INVALID_KEY, k1 = INVALID_KEY;  // partial slots of this chunk
    xyzz_t acc = xyzz_t::identity();
    for (uint32_t j = 0; j < L; j += 4) {
        const uint64_t pos = start + j;
        if (pos >= n_in) break;
        const uint4 k4 = *reinterpret_cast<const uint4 *>(keys + pos);
        const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t k = (pos + q < n_in) ? kk[q] : INVALID_KEY;  // n_in need not be a multiple of 4
            if (k == INVALID_KEY) continue;
            if (k != cur) {
                if (cur != INVALID_KEY) {
                    if (nruns == 1) {
                        k0 = cur;
                        store_xyzz(pts_out + 8 * (2 * t), acc);
                    } else {
                        store_xyzz(buckets + 8 * (uint64_t)cur, acc);  // interior run: sole owner of its bucket
                    }
                }
                cur = k;
                nruns++;
                acc = xyzz_t::identity();
            }
            xyzz_t p = load_xyzz(pts_in + 8 * (pos + q));
            add(acc, p);
        }
    }
    if (cur != INVALID_KEY) {
        if (nruns == 1) {
            k0 = cur;
            store_xyzz(pts_out + 8 * (2 * t), acc);
        } else {
            k1 = cur;
            store_xyzz(pts_out + 8 * (2 * t + 1), acc);
        }
    }
    keys_out[2 * t] = k0;
    keys_out[2 * t + 1] = k1;
}

// last level: <= FINAL_MAX slots, one CTA.  Compact the valid slots, run a Hillis-Steele segmented
// inclusive scan over equal keys, and let the tail of every run store its bucket.
__global__ void __launch_bounds__(FINAL_MAX) msm_reduce_final_kernel(const uint32_t *keys, const uint4 *pts_in, uint32_t n_in, uint4 *buckets) {
    __shared__ uint32_t s_keys[FINAL_MAX];
    __shared__ uint4 s_pts[FINAL_MAX * 8];
    __shared__ uint32_t total;
    const uint32_t tid = threadIdx.x;
    const uint32_t k = tid < n_in ? keys[tid] : INVALID_KEY;
    const uint32_t valid = k != INVALID_KEY;
    const uint32_t dense = block_exclusive_scan(valid, &total);
    if (valid) {
        s_keys[dense] = k;
        const uint4 *src = pts_in + 8 * (uint64_t)tid;
#pragma unroll
        for (int q = 0; q < 8; q++) s_pts[dense * 8 + q] = src[q];
    }
    __syncthreads();
    const uint32_t m = total;
    xyzz_t mine = xyzz_t::identity();
    uint32_t my_key = INVALID_KEY;
    if (tid < m) {
        my_key = s_keys[tid];
        mine = load_xyzz(s_pts + tid * 8);
    }
    for (uint32_t d = 1; d < m; d <<= 1) {
        const bool take = tid < m && tid >= d && s_keys[tid - d] == my_key;
        xyzz_t other = xyzz_t::identity();
        if (take) other = load_xyzz(s_pts + (tid - d) * 8);
        __syncthreads();
        if (take) {
            add(mine, other);
            store_xyzz(s_pts + tid * 8, mine);
        }
        __syncthreads();
    }
    if (tid < m && (tid == m - 1 || s_keys[tid + 1] != my_key)) store_xyzz(buckets + 8 * (uint64_t)my_key, mine);
}

// Middle levels with few slots (latency-bound): the same compaction + segmented scan, one CTA per FINAL_MAX slots.  A run that lies strictly
// inside the CTA's slots is complete (sorted keys: nobody else holds that key) and goes to its bucket; the runs touching the first / last
// valid slot may continue in the neighbouring CTAs and become this CTA's two partial slots.  8 dependent additions shrink the list 128 x,
// where the chunked kernel needs 8 dependent additions to shrink it 4 x.
__global__ void __launch_bounds__(FINAL_MAX) msm_reduce_cta_kernel(const uint32_t *keys, const uint4 *pts_in, uint64_t n_in, uint4 *buckets, uint32_t *keys_out,
                                                                   uint4 *pts_out) {
    __shared__ uint32_t s_keys[FINAL_MAX];
    __shared__ uint4 s_pts[FINAL_MAX * 8];
    __shared__ uint32_t total;
    const uint32_t tid = threadIdx.x;
    const uint64_t g = (uint64_t)blockIdx.x * FINAL_MAX + tid;
    const uint32_t k = g < n_in ? keys[g] : INVALID_KEY;
    const uint32_t valid = k != INVALID_KEY;
    const uint32_t dense = block_exclusive_scan(valid, &total);
    if (valid) {
        s_keys[dense] = k;
        const uint4 *src = pts_in + 8 * g;
#pragma unroll
        for (int q = 0; q < 8; q++) s_pts[dense * 8 + q] = src[q];
    }
    __syncthreads();
    const uint32_t m = total;
    xyzz_t mine = xyzz_t::identity();
    uint32_t my_key = INVALID_KEY;
    if (tid < m) {
        my_key = s_keys[tid];
        mine = load_xyzz(s_pts + tid * 8);
    }
    for (uint32_t d = 1; d < m; d <<= 1) {
        const bool take = tid < m && tid >= d && s_keys[tid - d] == my_key;
        xyzz_t other = xyzz_t::identity();
        if (take) other = load_xyzz(s_pts + (tid - d) * 8);
        __syncthreads();
        if (take) {
            add(mine, other);
            store_xyzz(s_pts + tid * 8, mine);
        }
        __syncthreads();
    }
    if (tid < 2) keys_out[2 * (uint64_t)blockIdx.x + tid] = INVALID_KEY;
    __syncthreads();
    if (tid < m && (tid == m - 1 || s_keys[tid + 1] != my_key)) {  // tail of a run: it holds the run's sum
        const bool first_run = my_key == s_keys[0], last_run = tid == m - 1;
        if (first_run) {
            keys_out[2 * (uint64_t)blockIdx.x] = my_key;
            store_xyzz(pts_out + 8 * (2 * (uint64_t)blockIdx.x), mine);
        } else if (last_run) {
            keys_out[2 * (uint64_t)blockIdx.x + 1] = my_key;
            store_xyzz(pts_out + 8 * (2 * (uint64_t)blockIdx.x + 1), mine);
        } else {
            store_xyzz(buckets + 8 * (uint64_t)my_key, mine);
        }
    }
}

// ------------------------------------------------------------------ 4: bucket reduction
// Window sum R = sum_i (i + 1) * B_i over the m = 2^(c-1) buckets of a window, as a hierarchy of running sums with no scalar
// multiplication: write R = sum_i [Q_i + lambda * i * P_i] (level 0: P = Q = B, lambda = 1).  A segment of S consecutive
// elements, i = s S + j, contributes  sum_j Q_j + lambda * sum_j j P_j  +  (lambda S) * s * T_s  with T_s = sum_j P_j, so the
// next level is the same problem on P'_s = T_s, Q'_s = sum_j Q_j + lambda * sum_j j P_j, lambda' = lambda S (a power of two:
// log2(lambda) doublings).  sum_j j P_j is the classic running sum (2 additions per element).  The recursion ends at m = 1,
// where R = Q_0.  Thread = (window, segment); level 0 does 2 general additions per bucket, every later level is 2^seg_log
// times smaller.
template <bool HAS_Q>
__global__ void __launch_bounds__(128) msm_bucket_level_kernel(const uint4 *P, const uint4 *Q /* !HAS_Q: Q aliases P (level 0) */, uint32_t n_windows, uint32_t m,
                                                               uint32_t seg_log, uint32_t lambda_log, uint4 *P_out, uint4 *Q_out) {
    const uint32_t nseg = m >> seg_log;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_windows * nseg) return;
    const uint32_t w = t / nseg, sg = t % nseg;
    const uint32_t S = 1u << seg_log;
    const uint64_t base = (uint64_t)w * m + (uint64_t)sg * S;
    xyzz_t run = xyzz_t::identity(), acc = xyzz_t::identity();
    for (uint32_t j = S - 1; j >= 1; j--) {
        xyzz_t b = load_xyzz(P + 8 * (base + j));
        add(run, b);
        add(acc, run);
    }
    {
        xyzz_t b = load_xyzz(P + 8 * base);
        add(run, b);  // run = T_s
    }
    store_xyzz(P_out + 8 * (uint64_t)t, run);
    for (uint32_t q = 0; q < lambda_log; q++) acc = dbl(acc);
    if (HAS_Q) {
        for (uint32_t j = 0; j < S; j++) {
            xyzz_t q = load_xyzz(Q + 8 * (base + j));
            add(acc, q);
        }
    } else {
        add(acc, run);
    }
    store_xyzz(Q_out + 8 * (uint64_t)t, acc);
}

// The remaining m elements of every window: V_s = Q_s + lambda * s * P_s by double-and-add over the bits of s (few elements, so the
// scalar multiplication is cheap here), then a CTA tree; one partial per CTA.  Q == nullptr: V_s = (s + 1) * P_s.
__global__ void __launch_bounds__(128) msm_bucket_finish_kernel(const uint4 *P, const uint4 *Q, uint32_t m, uint32_t lambda_log, uint32_t ctas_per_window, uint4 *partials) {
    __shared__ uint4 s_pts[128 * 8];
    const uint32_t w = blockIdx.x / ctas_per_window, cb = blockIdx.x % ctas_per_window, tid = threadIdx.x;
    const uint32_t sidx = cb * 128 + tid;
    xyzz_t v = xyzz_t::identity();
    if (sidx < m) {
        const uint64_t pos = (uint64_t)w * m + sidx;
        xyzz_t p = load_xyzz(P + 8 * pos);
        if (sidx && !p.is_identity()) {
            for (int bit = 31 - __clz(sidx); bit >= 0; bit--) {
                v = dbl(v);
                if ((sidx >> bit) & 1) add(v, p);
            }
            for (uint32_t q = 0; q < lambda_log; q++) v = dbl(v);
        }
        if (Q) {
            xyzz_t qv = load_xyzz(Q + 8 * pos);
            add(v, qv);
        } else {
            add(v, p);
        }
    }
    store_xyzz(s_pts + tid * 8, v);
    __syncthreads();
    for (uint32_t d = 64; d >= 1; d >>= 1) {
        if (tid < d) {
            xyzz_t o = load_xyzz(s_pts + (tid + d) * 8);
            add(v, o);
            store_xyzz(s_pts + tid * 8, v);
        }
        __syncthreads();
    }
    if (tid == 0) store_xyzz(partials + 8 * (uint64_t)blockIdx.x, v);
}

// one CTA per window: sum its partials (strided serial part + shared-memory tree)
__global__ void __launch_bounds__(128) msm_window_sum_kernel(const uint4 *seg_sums, uint32_t nseg_per_window, uint4 *win_sums) {
    __shared__ uint4 s_pts[128 * 8];
    const uint32_t w = blockIdx.x, tid = threadIdx.x;
    xyzz_t acc = xyzz_t::identity();
    for (uint32_t s = tid; s < nseg_per_window; s += blockDim.x) {
        xyzz_t p = load_xyzz(seg_sums + 8 * ((uint64_t)w * nseg_per_window + s));
        add(acc, p);
    }
    store_xyzz(s_pts + tid * 8, acc);
    __syncthreads();
    for (uint32_t d = blockDim.x >> 1; d >= 1; d >>= 1) {
        if (tid < d) {
            xyzz_t o = load_xyzz(s_pts + (tid + d) * 8);
            add(acc, o);
            store_xyzz(s_pts + tid * 8, acc);
        }
        __syncthreads();
    }
    if (tid == 0) store_xyzz(win_sums + 8 * (uint64_t)w, acc);
}

// ------------------------------------------------------------------ 4': bucket reduction as a tree (default)
// The running sums above are chains of dependent additions on few threads (16 + ~30 + 7 per launch set, ~4 us each: 0.5 ms per set whatever the
// load).  Here R = sum_i (i + 1) B_i = T + sum_b 2^b S_b with T = sum_i B_i and S_b = sum_{i : bit b of i set} B_i, and (T, S_0 .. S_{l-1}) of a
// block of 2^l consecutive buckets follow from its two halves by l + 1 independent additions (T = T_L + T_R, S_b = S_b^L + S_b^R, S_{l-1} = T_R):
// a CTA folds 256 buckets in 8 steps of at most one addition per thread (2 additions per bucket in total, like the running sums), a second launch
// folds the <= 256 node records of a set (CTA 0: the same tree over the nodes' T for the upper bits; CTAs 1..8: plain sums of the nodes' S_b), and
// the host finishes with the Horner fold over <= 16 bits that it does for windows anyway (host_g1.cpp).  Depth: 16 additions.
static const int BT_LOG = 8, BT = 1 << BT_LOG;         // buckets per CTA
static const int BT_SLOTS = BT_LOG + 1;                // node record: T, S_0 .. S_7
static const int BT_FIN = 2 * BT_LOG + 2;              // per set: T, S_0 .. S_15, and the plain sum of the Q vector when a running-sum level ran first
static const int BT_THREADS = BT / 2;                  // every step has at most BT / 2 additions
static const size_t BT_SMEM = (size_t)(BT + 3 * BT / 4) * 128;   // records of 2 buckets (BT points) | records of 4 buckets (3 BT / 4 points), ping-pong
extern __shared__ uint4 bt_smem[];
__global__ void __launch_bounds__(BT_THREADS, 3) msm_bucket_tree_kernel(const uint4 *in, uint64_t job_stride, uint32_t elem_stride, uint32_t m, uint32_t with_bits,
                                                                        uint32_t colsum, uint4 *out, uint32_t out_job_stride, uint32_t out_node_stride, uint32_t dst_slot0,
                                                                        uint32_t dst_bits0) {
    uint4 *buf[2] = {bt_smem + BT * 8, bt_smem};   // buf[1] (BT points) takes the records of 2 buckets, buf[0] those of 4, ...
    const uint32_t tid = threadIdx.x, job = blockIdx.y;
    // colsum: every CTA of the row reads the same m records, CTA x their slot x; only CTA 0 keeps the bit sums
    const uint32_t node = colsum ? 0u : blockIdx.x, src_off = colsum ? blockIdx.x : 0u;
    const bool bits = with_bits && (!colsum || blockIdx.x == 0);
    {   // records of 2 buckets straight from global memory: T = V_2t + V_2t+1, S_0 = V_2t+1
        const uint64_t e = (uint64_t)node * BT + 2 * tid;
        const uint4 *src = in + 8 * ((uint64_t)job * job_stride + src_off);
        xyzz_t a = xyzz_t::identity(), b = xyzz_t::identity();
        if (e < m) a = load_xyzz(src + 8 * (e * elem_stride));
        if (e + 1 < m) b = load_xyzz(src + 8 * ((e + 1) * elem_stride));
        if (bits) store_xyzz(buf[1] + 8 * (2 * tid + 1), b);
        add(a, b);
        store_xyzz(buf[1] + 8 * (bits ? 2 * tid : tid), a);
    }
    __syncthreads();
    int cur = 1;
    for (uint32_t l = 1; l < (uint32_t)BT_LOG; l++) {   // records of 2^l buckets -> records of 2^(l+1)
        const uint32_t per_in = bits ? l + 1 : 1, per_out = bits ? l + 2 : 1, n_out = (uint32_t)BT >> (l + 1);
        if (tid < n_out * per_in) {
            const uint32_t j = tid / per_in, q = tid - j * per_in;
            xyzz_t a = load_xyzz(buf[cur] + 8 * ((2 * j) * per_in + q));
            const xyzz_t b = load_xyzz(buf[cur] + 8 * ((2 * j + 1) * per_in + q));
            if (bits && q == 0) store_xyzz(buf[cur ^ 1] + 8 * (j * per_out + l + 1), b);  // S_l = T of the upper half
            add(a, b);
            store_xyzz(buf[cur ^ 1] + 8 * (j * per_out + q), a);
        }
        __syncthreads();
        cur ^= 1;
    }
    uint4 *dst = out + 8 * ((uint64_t)job * out_job_stride + (uint64_t)node * out_node_stride);
    const uint32_t n_rec = bits ? (uint32_t)BT_SLOTS : 1u;
    for (uint32_t i = tid; i < n_rec * 8; i += BT_THREADS) {
        const uint32_t slot = i >> 3;
        const uint32_t d = slot == 0 ? dst_slot0 + (colsum ? blockIdx.x : 0u) : dst_bits0 + slot - 1;
        dst[8 * d + (i & 7)] = buf[cur][i];
    }
}

// ------------------------------------------------------------------ fixed-base window tables
// tables[w * stride + i] = 2^(c w) * P_i (affine) for w < W: with them every window of the scalar addresses the SAME bucket
// set, so the window size can grow (c = 20: 13 windows instead of 16) at no bucket-reduction cost.  One thread per base: the
// chain of c doublings per window in XYZZ, then ONE inversion per base (Montgomery trick over its W - 1 table entries).
static const int TAB_MAX_W = 24;
__global__ void __launch_bounds__(128) msm_table_build_kernel(const uint4 *bases, uint64_t n, uint64_t stride, uint32_t c, uint32_t W, uint4 *tables) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= stride) return;
    affine_t p;
    p.x = fq_t::zero();
    p.y = fq_t::zero();
    if (i < n) {
        p.x = load_fp<FqParams>(bases + 4 * i);
        p.y = load_fp<FqParams>(bases + 4 * i + 2);
    }
    store_fp(tables + 4 * i, p.x);
    store_fp(tables + 4 * i + 2, p.y);
    xyzz_t pts[TAB_MAX_W];
    fq_t pref[TAB_MAX_W];
    xyzz_t cur = xyzz_t::from_affine(p);
    fq_t prod = fq_t::one();
    for (uint32_t w = 1; w < W; w++) {
        for (uint32_t q = 0; q < c; q++) cur = dbl(cur);
        pts[w] = cur;
        pref[w] = prod;
        if (!cur.is_identity()) prod = mul(prod, mul(cur.zz, cur.zzz));
    }
    fq_t inv_all = inv(prod);
    for (uint32_t w = W - 1; w >= 1; w--) {
        affine_t r;
        r.x = fq_t::zero();
        r.y = fq_t::zero();
        if (!pts[w].is_identity()) {
            fq_t zi = mul(inv_all, pref[w]);                    // 1 / (zz * zzz) of entry w
            inv_all = mul(inv_all, mul(pts[w].zz, pts[w].zzz));
            r.x = mul(pts[w].x, mul(zi, pts[w].zzz));
            r.y = mul(pts[w].y, mul(zi, pts[w].zz));
        }
        store_fp(tables + 4 * ((uint64_t)w * stride + i), r.x);
        store_fp(tables + 4 * ((uint64_t)w * stride + i) + 2, r.y);
    }
}

// ------------------------------------------------------------------ host driver
static uint32_t ilog2_floor(uint64_t x) {
    uint32_t r = 0;
    while (x >>= 1) r++;
    return r;
}

static MsmShape msm_shape(sb_ctx *ctx, uint64_t n, int32_t w_lo = 0, int32_t w_hi = -1, const MsmTables *tabs = nullptr, uint32_t batch = 1, uint32_t res = 0,
                          uint32_t log_mod = 0) {
    MsmShape sh;
    memset(&sh, 0, sizeof sh);
    int c = (int)ilog2_floor(n) - 3;
    if (c < 4) c = 4;
    if (c > 16) c = 16;
    if (ctx->tune.msm_c >= 2 && ctx->tune.msm_c <= 16) c = ctx->tune.msm_c;
    if (tabs) c = (int)tabs->c;
    sh.c = (uint32_t)c;
    sh.W_all = (255 + sh.c - 1) / sh.c;
    sh.tab_stride = tabs ? (uint32_t)tabs->stride : 0;
    sh.w_lo = (uint32_t)w_lo;
    sh.W = (w_hi < 0 ? sh.W_all : (uint32_t)w_hi) - sh.w_lo;
    sh.res = res;
    sh.log_mod = log_mod;
    sh.B = (1u << (sh.c - 1)) >> log_mod;
    sh.batch = batch;
    sh.Wb = tabs ? batch : sh.W;
    sh.n = n;
    sh.t_max = ((uint64_t)sh.W * n * batch + 3) & ~3ull;
    const uint64_t want = (uint64_t)ctx->sm_count * 1024;
    uint32_t L1 = 64;
    while (L1 > 8 && sh.t_max / L1 < want) L1 >>= 1;
    if (ctx->tune.msm_l1 >= 4 && ctx->tune.msm_l1 <= 1024 && (ctx->tune.msm_l1 % 4) == 0) L1 = (uint32_t)ctx->tune.msm_l1;
    sh.L1 = L1;
    sh.seg_log = sh.c - 1 >= 3 ? 3 : sh.c - 1;
    if (ctx->tune.msm_seg >= 0 && ctx->tune.msm_seg <= (int)sh.c - 1) sh.seg_log = (uint32_t)ctx->tune.msm_seg;
    return sh;
}

void msm_window_shape(sb_ctx *ctx, size_t n, uint32_t *c, uint32_t *W) {
    const MsmShape sh = msm_shape(ctx, n);
    *c = sh.c;
    *W = sh.W_all;
}

static int32_t msm_run_impl(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, int32_t w_lo, int32_t w_hi, const MsmTables *tabs, uint8_t *out, cudaStream_t st,
                            uint32_t batch = 1, const uint32_t *piece_off = nullptr, uint32_t res = 0, uint32_t log_mod = 0);

int32_t msm_run(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, uint8_t out_affine[64], cudaStream_t st) {
    return msm_run_impl(ctx, d_bases, d_scalars, n, 0, -1, nullptr, out_affine, st);
}
// windows [w_lo, w_hi) only: writes (w_hi - w_lo) XYZZ window sums (128 B each) instead of the folded point
int32_t msm_run_windows(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, uint32_t w_lo, uint32_t w_hi, uint8_t *win_out, cudaStream_t st) {
    SB_REQUIRE(w_lo < w_hi, "msm_run_windows: empty window range");
    return msm_run_impl(ctx, d_bases, d_scalars, n, (int32_t)w_lo, (int32_t)w_hi, nullptr, win_out, st);
}
void msm_fold_windows(const uint8_t *win, uint32_t W, uint32_t c, uint8_t out_affine[64]) { host_fold_windows(win, (int)W, (int)c, out_affine); }

// fixed-base tables: the whole MSM (out: 64 B affine) or windows [w_lo, w_hi) of it (out: ONE 128 B XYZZ partial that already carries its 2^(c w) factors)
int32_t msm_run_tables(sb_ctx *ctx, const MsmTables *tabs, const void *d_scalars, size_t n, int32_t w_lo, int32_t w_hi, uint8_t *out, cudaStream_t st) {
    SB_REQUIRE(tabs && tabs->d_tables && n <= tabs->stride, "msm_run_tables: more scalars than table columns");
    return msm_run_impl(ctx, tabs->d_tables, d_scalars, n, w_lo, w_hi, tabs, out, st);
}
// windows [w_lo, w_hi) of `batch` scalar vectors in one launch set: out = batch XYZZ partials (128 B each) that carry their 2^(c w) factors
int32_t msm_run_tables_batch_windows(sb_ctx *ctx, const MsmTables *tabs, const void *d_scalars, size_t n, uint32_t batch, int32_t w_lo, int32_t w_hi, uint8_t *out,
                                     cudaStream_t st, const uint32_t *piece_off, uint32_t res, uint32_t log_mod) {
    SB_REQUIRE(tabs && tabs->d_tables && n <= tabs->stride && batch >= 1 && w_lo < w_hi, "msm_run_tables_batch_windows: bad arguments");
    SB_REQUIRE((uint64_t)(w_hi - w_lo) * n * batch < (1ull << 32) - 8, "msm_run_tables_batch_windows: batch too large");
    return msm_run_impl(ctx, tabs->d_tables, d_scalars, n, w_lo, w_hi, tabs, out, st, batch, piece_off, res, log_mod);
}

int32_t msm_tables_build(sb_ctx *ctx, const void *d_bases, size_t n, uint32_t c, MsmTables *out, cudaStream_t st, void *d_dst) {
    SB_REQUIRE(c >= 11 && c <= 24, "msm tables: window bits must be 11..24");
    const uint32_t W = (255 + c - 1) / c;
    SB_REQUIRE(W <= (uint32_t)TAB_MAX_W && (uint64_t)W * n < (1ull << 31), "msm tables: too many windows / points");
    void *d_tab = d_dst;
    if (!d_tab) {
        cudaError_t e = cudaMalloc(&d_tab, (size_t)W * n * 64);
        if (e != cudaSuccess) { set_last_error("msm tables: cudaMalloc(%zu): %s", (size_t)W * n * 64, cudaGetErrorString(e)); return SB_ERR_ALLOC; }
    }
    SB_LAUNCH(ctx, msm_table_build_kernel, (unsigned)((n + 127) / 128), 128, 0, st, (const uint4 *)d_bases, (uint64_t)n, (uint64_t)n, c, W, (uint4 *)d_tab);
    SB_CUDA_TRY(sync_stream(ctx, st));
    out->d_tables = d_tab;
    out->c = c;
    out->W = W;
    out->stride = n;
    return SB_OK;
}

// `batch` scalar vectors (contiguous, n each) against the same table bases in ONE launch set: out = batch x 64 B affine commitments.
// The vectors only differ in which bucket set their digits land in, so sort, accumulation and bucket reduction run once over
// batch * W * n digits and the latency-bound tails (upper reduce levels, bucket hierarchy, host round trip) are paid once.
int32_t msm_run_tables_batch(sb_ctx *ctx, const MsmTables *tabs, const void *d_scalars, size_t n, uint32_t batch, uint8_t *out_affine, cudaStream_t st) {
    SB_REQUIRE(tabs && tabs->d_tables && n <= tabs->stride && batch >= 1, "msm_run_tables_batch: bad arguments");
    // one launch set as long as the sorted list fits 32-bit positions
    uint32_t done = 0;
    while (done < batch) {
        uint32_t take = batch - done;
        while (take > 1 && (uint64_t)tabs->W * n * take >= (1ull << 32) - 8) take--;
        SB_TRY(msm_run_impl(ctx, tabs->d_tables, (const uint8_t *)d_scalars + (size_t)done * n * 32, n, 0, -1, tabs, out_affine + (size_t)done * 64, st, take));
        done += take;
    }
    return SB_OK;
}

// a batch whose vectors commit to DIFFERENT bases of one SRS (advice columns over the Lagrange basis + the random polynomial over the monomial
// basis): both tables live in one slab, a vector's table is selected by an entry offset.  basis_of[j] in {0, 1}; at most 8 vectors.
int32_t msm_run_tables_batch_mixed(sb_ctx *ctx, const MsmTables *t0, const MsmTables *t1, const void *d_scalars, size_t n, uint32_t batch, const uint8_t *basis_of,
                                   uint8_t *out_affine, cudaStream_t st) {
    SB_REQUIRE(t0 && t1 && t0->d_tables && t1->d_tables && t0->c == t1->c && t0->W == t1->W && t0->stride == t1->stride, "mixed batch: the two tables must have one shape");
    SB_REQUIRE((const uint8_t *)t1->d_tables == (const uint8_t *)t0->d_tables + (size_t)t0->W * t0->stride * 64, "mixed batch: the tables must be adjacent in one slab");
    SB_REQUIRE(batch >= 1 && batch <= 8 && n <= t0->stride && (uint64_t)t0->W * n * batch < (1ull << 32) - 8 && 2ull * t0->W * t0->stride < (1ull << 31),
               "mixed batch: too large");
    uint32_t off[8];
    for (uint32_t j = 0; j < batch; j++) off[j] = basis_of[j] ? (uint32_t)((uint64_t)t0->W * t0->stride) : 0u;
    return msm_run_impl(ctx, t0->d_tables, d_scalars, n, 0, -1, t0, out_affine, st, batch, off);
}

static int32_t msm_run_impl(sb_ctx *ctx, const void *d_bases, const void *d_scalars, size_t n, int32_t w_lo, int32_t w_hi, const MsmTables *tabs, uint8_t *out_affine,
                            cudaStream_t st, uint32_t batch, const uint32_t *piece_off, uint32_t res, uint32_t log_mod) {
    if (n == 0) {
        memset(out_affine, 0, w_hi < 0 ? (size_t)64 * batch : (size_t)(tabs ? batch : (uint32_t)(w_hi - w_lo)) * 128);
        return SB_OK;
    }
    SB_REQUIRE(n < (1ull << 31), "msm: n must be < 2^31");
    SB_REQUIRE(batch == 1 || tabs, "msm: batches need table bases");
    SB_REQUIRE(log_mod == 0 || (tabs && w_hi >= 0 && res < (1u << log_mod) && tabs->c >= log_mod + 2), "msm: residue shards need table bases and a window range");
    MsmShape sh = msm_shape(ctx, n, w_lo, w_hi, tabs, batch, res, log_mod);
    SB_REQUIRE(!piece_off || batch <= 8, "msm: at most 8 vectors in a mixed-basis batch");
    for (uint32_t j = 0; j < 8; j++) sh.piece_off[j] = (piece_off && j < batch) ? piece_off[j] : 0;
    SB_REQUIRE(w_hi < 0 || (uint32_t)w_hi <= sh.W_all, "msm: window range exceeds the scalar");
    SB_REQUIRE(sh.t_max < (1ull << 32), "msm: window count * n must be < 2^32");
    const uint64_t nb = (uint64_t)sh.Wb * sh.B;

    // ---- scratch ----
    uint32_t *d_counts, *d_cursor, *d_tiles, *d_svals;
    uint4 *d_buckets, *d_seg, *d_win;
    const uint64_t ntiles = (nb + SCAN_TILE - 1) / SCAN_TILE;
    SB_TRY(scratch_get(ctx, "msm_counts", (nb + 4) * 4, (void **)&d_counts));
    SB_TRY(scratch_get(ctx, "msm_cursor", (nb + 4) * 4, (void **)&d_cursor));
    SB_TRY(scratch_get(ctx, "msm_tiles", (ntiles + 4) * 4, (void **)&d_tiles));
    SB_TRY(scratch_get(ctx, "msm_svals", (sh.t_max + 64) * 4, (void **)&d_svals));
    SB_TRY(scratch_get(ctx, "msm_buckets", nb * 128, (void **)&d_buckets));
    const uint32_t nseg = sh.B >> sh.seg_log;
    {
        const uint64_t cap = (uint64_t)sh.Wb * (sh.B >> 1) + 128;
        SB_TRY(scratch_get(ctx, "msm_seg", (4 * cap + (uint64_t)sh.Wb * ((sh.B + 127) / 128) + 8 + sh.Wb + 8) * 128, (void **)&d_seg));  // bucket hierarchy scratch
    }
    (void)nseg;
    const uint64_t nchunks1 = (sh.t_max + sh.L1 - 1) / sh.L1;
    uint64_t slots_a = 2 * nchunks1;                                   // level-1 output
    uint64_t slots_b = 2 * ((slots_a + LK - 1) / LK);                  // level-2 output
    uint32_t *d_keys_a, *d_keys_b;
    uint4 *d_pts_a, *d_pts_b;
    SB_TRY(scratch_get(ctx, "msm_keys_a", (slots_a + 4) * 4, (void **)&d_keys_a));
    SB_TRY(scratch_get(ctx, "msm_pts_a", slots_a * 128, (void **)&d_pts_a));
    SB_TRY(scratch_get(ctx, "msm_keys_b", (slots_b + 4) * 4, (void **)&d_keys_b));
    SB_TRY(scratch_get(ctx, "msm_pts_b", slots_b * 128, (void **)&d_pts_b));

    for (int e = 0; e < 5; e++)
        if (!ctx->msm_ev[e]) SB_CUDA_TRY(cudaEventCreate(&ctx->msm_ev[e]));
    SB_CUDA_TRY(cudaEventRecord(ctx->msm_ev[0], st));

    // ---- 1+2: histogram, scan, scatter ----
    SB_CUDA_TRY(cudaMemsetAsync(d_counts, 0, (nb + 4) * 4, st));
    SB_CUDA_TRY(cudaMemsetAsync(d_buckets, 0, nb * 128, st));
    unsigned sort_grid = (unsigned)((n * batch + 255) / 256);
    const unsigned max_grid = (unsigned)ctx->sm_count * 16;
    if (sort_grid > max_grid) sort_grid = max_grid;
    SB_LAUNCH(ctx, msm_sort_kernel<false>, sort_grid, 256, 0, st, (const uint4 *)d_scalars, (uint64_t)n, sh, d_counts, nullptr);
    SB_LAUNCH(ctx, scan_tile_sums_kernel, (unsigned)ntiles, SCAN_THREADS, 0, st, d_counts, nb, d_tiles);
    SB_LAUNCH(ctx, scan_of_tile_sums_kernel, 1, 1024, 0, st, d_tiles, ntiles, d_counts + nb);
    SB_LAUNCH(ctx, scan_apply_kernel, (unsigned)ntiles, SCAN_THREADS, 0, st, d_counts, nb, d_tiles, d_counts, d_cursor);
    SB_LAUNCH(ctx, msm_sort_kernel<true>, sort_grid, 256, 0, st, (const uint4 *)d_scalars, (uint64_t)n, sh, d_cursor, d_svals);

    // ---- 3: reduce-by-key levels ----
    SB_CUDA_TRY(cudaEventRecord(ctx->msm_ev[1], st));
    SB_LAUNCH(ctx, msm_reduce_first_kernel, (unsigned)((nchunks1 + 127) / 128), 128, 0, st, (const uint32_t *)d_counts, (uint32_t)nb, d_svals, (const uint4 *)d_bases,
              sh.L1, d_buckets, d_keys_a, d_pts_a, nchunks1);
    SB_CUDA_TRY(cudaEventRecord(ctx->msm_ev[2], st));
    uint64_t slots = slots_a;
    uint32_t *kin = d_keys_a, *kout = d_keys_b;
    uint4 *pin = d_pts_a, *pout = d_pts_b;
    const uint64_t CTA_SCAN_MAX = (uint64_t)ctx->tune.msm_cta_scan_max;  // below this many slots the levels are latency-bound: one CTA-wide scan per 256 slots
    while (slots > FINAL_MAX) {
        uint64_t nch;
        if (slots <= CTA_SCAN_MAX && !ctx->tune.msm_no_cta_scan) {
            nch = (slots + FINAL_MAX - 1) / FINAL_MAX;
            SB_LAUNCH(ctx, msm_reduce_cta_kernel, (unsigned)nch, FINAL_MAX, 0, st, kin, (const uint4 *)pin, slots, d_buckets, kout, pout);
        } else {
            nch = (slots + LK - 1) / LK;
            SB_LAUNCH(ctx, msm_reduce_kernel, (unsigned)((nch + 127) / 128), 128, 0, st, kin, (const uint4 *)pin, slots, (uint32_t)LK, d_buckets, kout, pout, nch);
        }
        slots = 2 * nch;
        uint32_t *tk = kin; kin = kout; kout = tk;
        uint4 *tp = pin; pin = pout; pout = tp;
    }
    SB_LAUNCH(ctx, msm_reduce_final_kernel, 1, FINAL_MAX, 0, st, kin, (const uint4 *)pin, (uint32_t)slots, d_buckets);

    // ---- 4: bucket reduction ----
    SB_CUDA_TRY(cudaEventRecord(ctx->msm_ev[3], st));
    const bool tree = !ctx->tune.msm_no_bucket_tree && (size_t)sh.Wb * BT_FIN * 128 + 64 <= ctx->pinned_bytes;
    SB_REQUIRE(tree || log_mod == 0, "msm: residue shards need the tree bucket reduction");
    uint32_t tree_bits = 0, tree_shift = 0;   // R = X + 2^shift * sum_{b < bits} 2^b S_b
    uint4 *d_fin = nullptr;
    if (tree) {
        const uint64_t cap = (uint64_t)sh.Wb * (sh.B >> 1) + 128;
        const uint4 *bP = d_buckets, *bQ = nullptr;
        uint32_t m = sh.B;
        if (m > (uint32_t)BT * BT) {   // more than 2^16 buckets per set: one running-sum level first (segments of m / 2^16 buckets per thread)
            tree_shift = ilog2_floor(m) - 2 * BT_LOG;
            const uint32_t threads = sh.Wb * (m >> tree_shift);
            SB_LAUNCH(ctx, msm_bucket_level_kernel<false>, (threads + 127) / 128, 128, 0, st, bP, bQ, sh.Wb, m, tree_shift, 0u, d_seg, d_seg + 8 * cap);
            bP = d_seg;
            bQ = d_seg + 8 * cap;
            m >>= tree_shift;
        }
        tree_bits = ilog2_floor(m);
        const uint32_t n_nodes = (m + BT - 1) / BT;
        uint4 *d_nodes;
        SB_TRY(scratch_get(ctx, "msm_fin", ((uint64_t)sh.Wb * BT_FIN + 8) * 128, (void **)&d_fin));
        SB_TRY(scratch_get(ctx, "msm_nodes", (2 * (uint64_t)sh.Wb * n_nodes * BT_SLOTS + 8) * 128, (void **)&d_nodes));
        static bool attr_set[64] = {false};  // the attribute is per device: a process may drive several GPUs through several contexts
        if (ctx->device >= 64 || !attr_set[ctx->device]) {
            SB_CUDA_TRY(cudaFuncSetAttribute(msm_bucket_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BT_SMEM));
            if (ctx->device < 64) attr_set[ctx->device] = true;
        }
        const size_t smem = BT_SMEM;
        if (n_nodes == 1) {
            SB_LAUNCH(ctx, msm_bucket_tree_kernel, dim3(1, sh.Wb), BT_THREADS, smem, st, bP, (uint64_t)m, 1u, m, 1u, 0u, d_fin, (uint32_t)BT_FIN, 0u, 0u, 1u);
        } else {
            SB_LAUNCH(ctx, msm_bucket_tree_kernel, dim3(n_nodes, sh.Wb), BT_THREADS, smem, st, bP, (uint64_t)m, 1u, m, 1u, 0u, d_nodes, n_nodes * BT_SLOTS, (uint32_t)BT_SLOTS, 0u, 1u);
            SB_LAUNCH(ctx, msm_bucket_tree_kernel, dim3(BT_SLOTS, sh.Wb), BT_THREADS, smem, st, (const uint4 *)d_nodes, (uint64_t)n_nodes * BT_SLOTS, (uint32_t)BT_SLOTS, n_nodes, 1u, 1u,
                      d_fin, (uint32_t)BT_FIN, 0u, 0u, (uint32_t)BT_SLOTS);
        }
        if (bQ) {   // plain sum of the Q vector (m = 2^16 elements per set) into slot BT_FIN - 1
            uint4 *d_nodes_q = d_nodes + 8 * ((uint64_t)sh.Wb * n_nodes * BT_SLOTS);
            SB_LAUNCH(ctx, msm_bucket_tree_kernel, dim3(n_nodes, sh.Wb), BT_THREADS, smem, st, bQ, (uint64_t)m, 1u, m, 0u, 0u, d_nodes_q, n_nodes * BT_SLOTS, (uint32_t)BT_SLOTS, 0u, 1u);
            SB_LAUNCH(ctx, msm_bucket_tree_kernel, dim3(1, sh.Wb), BT_THREADS, smem, st, (const uint4 *)d_nodes_q, (uint64_t)n_nodes * BT_SLOTS, (uint32_t)BT_SLOTS, n_nodes, 0u, 1u,
                      d_fin, (uint32_t)BT_FIN, 0u, (uint32_t)BT_FIN - 1, 0u);
        }
    } else {
        // scratch layout (XYZZ slots): P/Q ping-pong of the level kernels, CTA partials, window sums
        const uint64_t cap = (uint64_t)sh.Wb * (sh.B >> 1) + 128;
        uint4 *bufP[2] = {d_seg, d_seg + 8 * cap}, *bufQ[2] = {d_seg + 16 * cap, d_seg + 24 * cap};
        uint4 *d_part = d_seg + 32 * cap;
        d_win = d_part + 8 * ((uint64_t)sh.Wb * ((sh.B + 127) / 128) + 8);
        const uint4 *bP = d_buckets, *bQ = nullptr;
        uint32_t m = sh.B, lambda_log = 0;
        int pp = 0;
        const uint32_t finish_at = (uint32_t)ctx->tune.msm_finish_at;  // below this the level kernels are latency-bound and the finish (double-and-add + CTA tree) is shorter
        for (int level = 0; level < 2 && m > finish_at; level++) {
            const uint32_t sl = level == 0 ? (sh.seg_log ? sh.seg_log : 1) : (uint32_t)ctx->tune.msm_seg1;  // level 1 is latency-bound: segments of 4
            const uint32_t threads = sh.Wb * (m >> sl);
            if (bQ) SB_LAUNCH(ctx, msm_bucket_level_kernel<true>, (threads + 127) / 128, 128, 0, st, bP, bQ, sh.Wb, m, sl, lambda_log, bufP[pp], bufQ[pp]);
            else SB_LAUNCH(ctx, msm_bucket_level_kernel<false>, (threads + 127) / 128, 128, 0, st, bP, bQ, sh.Wb, m, sl, lambda_log, bufP[pp], bufQ[pp]);
            bP = bufP[pp];
            bQ = bufQ[pp];
            pp ^= 1;
            lambda_log += sl;
            m >>= sl;
        }
        const uint32_t ctas = (m + 127) / 128;
        SB_LAUNCH(ctx, msm_bucket_finish_kernel, sh.Wb * ctas, 128, 0, st, bP, bQ, m, lambda_log, ctas, d_part);
        SB_LAUNCH(ctx, msm_window_sum_kernel, sh.Wb, 128, 0, st, (const uint4 *)d_part, ctas, d_win);
    }

    // ---- 5: window sums -> host fold ----
    SB_CUDA_TRY(cudaEventRecord(ctx->msm_ev[4], st));
    if (tree) SB_CUDA_TRY(cudaMemcpyAsync(ctx->pinned, d_fin, (size_t)sh.Wb * BT_FIN * 128, cudaMemcpyDeviceToHost, st));
    else SB_CUDA_TRY(cudaMemcpyAsync(ctx->pinned, d_win, (size_t)sh.Wb * 128, cudaMemcpyDeviceToHost, st));
    // the number of non-zero signed digits that were sorted and accumulated (the scan's grand total): the level-1 additions actually performed
    uint32_t *h_total = (uint32_t *)((uint8_t *)ctx->pinned + ctx->pinned_bytes - 64);
    SB_CUDA_TRY(cudaMemcpyAsync(h_total, d_counts + nb, 4, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(sync_stream(ctx, st));
    for (int e = 0; e < 4; e++) cudaEventElapsedTime(&ctx->msm_phase_ms[e], ctx->msm_ev[e], ctx->msm_ev[e + 1]);
    cudaEventElapsedTime(&ctx->msm_phase_ms[4], ctx->msm_ev[0], ctx->msm_ev[4]);
    for (int e = 0; e < 5; e++) ctx->acc_msm_ms[e] += ctx->msm_phase_ms[e];   // running totals since sb_perf_reset (one proof = several launch sets)
    ctx->acc_msm_digits += *h_total;
    ctx->acc_msm_sets++;
    ctx->acc_msm_d2h += (tree ? (uint64_t)sh.Wb * BT_FIN * 128 : (uint64_t)sh.Wb * 128) + 4;
    ctx->msm_last_shape[0] = sh.c; ctx->msm_last_shape[1] = sh.W; ctx->msm_last_shape[2] = sh.L1; ctx->msm_last_shape[3] = sh.seg_log;
    if (tree) {   // per set: X + 2^shift * sum_b 2^b S_b, written back over the head of the pinned buffer as one XYZZ point per set (set j's record starts
                  // at j * BT_FIN * 128 >= j * 128, so the compaction never overtakes its input)
        uint8_t *hp = (uint8_t *)ctx->pinned;
        for (uint32_t j = 0; j < sh.Wb; j++) {
            uint8_t pt[128];
            host_bucket_combine(hp + (size_t)j * BT_FIN * 128, (int)tree_bits, (int)tree_shift, tree_shift ? BT_FIN - 1 : 0, pt);
            // residue shard: bucket b' stands for the digit 2^log_mod * b' + res + 1, so the sum is 2^log_mod * sum (b' + 1) V_b' - (2^log_mod - res - 1) * sum V_b'
            if (log_mod) host_residue_fixup(pt, hp + (size_t)j * BT_FIN * 128 /* slot 0: sum of all buckets */, (int)log_mod, (int)res);
            memcpy(hp + (size_t)j * 128, pt, 128);
        }
    }
    if (w_hi >= 0) memcpy(out_affine, ctx->pinned, (size_t)sh.Wb * 128);
    else if (tabs)
        for (uint32_t j = 0; j < batch; j++) host_fold_windows((const uint8_t *)ctx->pinned + (size_t)j * 128, 1, (int)sh.c, out_affine + (size_t)j * 64);
    else host_fold_windows((const uint8_t *)ctx->pinned, (int)sh.Wb, (int)sh.c, out_affine);
    return SB_OK;
}

}  // namespace sb
