// Merkle sum tree of Summa on the GPU: Keccak-256 of the usernames, Poseidon (WIDTH 2, RATE 1, 8 + 56 rounds) leaf and
// middle-node hashes, level-by-level build, Merkle-proof extraction.
//
// Restates (not translates) zk_prover/src/merkle_sum_tree/{entry.rs:15-38, node.rs:16-85, mst.rs:96-134, tree.rs:22-137,
// utils/build_tree.rs:5-78}: the reference walks `Vec<Vec<Node>>` with rayon, one level at a time.  Here the whole tree
// lives in two flat HBM arrays (hashes, balances as one plane per currency) indexed by a closed-form level offset, every
// level is one launch of one-thread-per-node Poseidon sponges, and the levels with <= 256 nodes are folded into a single
// one-CTA kernel.  Bound: fmaheavy (a leaf costs N_CURRENCIES + 1 permutations of 417 field products, a middle node N_CURRENCIES + 2).
#include "common.cuh"
#include "handles.h"
#include "poseidon_constants.inc"

namespace sb {

struct PoseidonConsts {
    fr_t rc[64][2];
    fr_t mds[2][2];
    fr_t partial[56][4];  // rounds 4..59 in scaled-lane form: rc0, rc1 / d, w = m01 d, v = m10 / (m11 d)
    fr_t partial_scale;   // d after the last partial round
};
__constant__ PoseidonConsts c_pos;
static std::mutex g_pos_mu;
static bool g_pos_loaded[64] = {false};

static int32_t poseidon_consts_load(int device) {
    std::lock_guard<std::mutex> lk(g_pos_mu);
    if (device < 64 && g_pos_loaded[device]) return SB_OK;
    PoseidonConsts h;
    memcpy(h.rc, POSEIDON_RC_HOST, sizeof(h.rc));
    memcpy(h.mds, POSEIDON_MDS_HOST, sizeof(h.mds));
    memcpy(h.partial, POSEIDON_PARTIAL_HOST, sizeof(h.partial));
    memcpy(&h.partial_scale, POSEIDON_PARTIAL_SCALE_HOST, sizeof(h.partial_scale));
    SB_CUDA_TRY(cudaMemcpyToSymbol(c_pos, &h, sizeof(h)));
    if (device < 64) g_pos_loaded[device] = true;
    return SB_OK;
}

__device__ __forceinline__ fr_t pow5(const fr_t &x) {
    fr_t x2 = sqr(x);
    fr_t x4 = sqr(x2);
    return mul(x4, x);
}

// halo2_gadgets poseidon::primitives::permute for T = 2: add round constants, S-box on both lanes in the
// 4 + 4 full rounds and on lane 0 only in the 56 partial rounds, then the MDS product.  The partial rounds run in an exactly
// equivalent "scaled lane" form (lane 1 = d * s1_hat, d updated analytically; constants precomputed by
// tools/gen_poseidon_header.py): 6 instead of 7 products per round, 417 instead of 472 per permutation, same field elements out.
__device__ __forceinline__ void poseidon_full_round(fr_t &s0, fr_t &s1, int r) {
    fr_t a = pow5(add(s0, c_pos.rc[r][0]));
    fr_t b = pow5(add(s1, c_pos.rc[r][1]));
    s0 = add(mul(c_pos.mds[0][0], a), mul(c_pos.mds[0][1], b));
    s1 = add(mul(c_pos.mds[1][0], a), mul(c_pos.mds[1][1], b));
}
__device__ __noinline__ void poseidon_permute(fr_t &s0, fr_t &s1) {
#pragma unroll 1
    for (int r = 0; r < 4; r++) poseidon_full_round(s0, s1, r);
#pragma unroll 1
    for (int r = 0; r < 56; r++) {
        fr_t a = pow5(add(s0, c_pos.partial[r][0]));
        fr_t b = add(s1, c_pos.partial[r][1]);
        s0 = add(mul(c_pos.mds[0][0], a), mul(c_pos.partial[r][2], b));
        s1 = add(mul(c_pos.partial[r][3], a), b);
    }
    s1 = mul(s1, c_pos.partial_scale);
#pragma unroll 1
    for (int r = 60; r < 64; r++) poseidon_full_round(s0, s1, r);
}

// ConstantLength<L> sponge, RATE 1: capacity lane starts at L * 2^64
__device__ __forceinline__ fr_t sponge_init_capacity(uint32_t L) {
    fr_t c = fr_t::zero();
    c.v[2] = L;  // L * 2^64, canonical
    return to_mont(c);
}

// level offset inside the flat arrays: levels 0..depth hold 2^(depth - l) nodes each
__host__ __device__ __forceinline__ uint64_t level_off(uint32_t depth, uint32_t l) { return (2ull << depth) - (2ull << (depth - l)); }

// ---------------------------------------------------------------------------------------- keccak-256 (usernames)
__constant__ uint64_t c_keccak_rc[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull, 0x0000000080000001ull,
                                         0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
                                         0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull,
                                         0x000000000000800aull, 0x800000008000000aull, 0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }
__device__ void keccak_f1600(uint64_t A[25]) {
    const int rho[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    const int pi[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        uint64_t C[5];
#pragma unroll
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) {
            uint64_t d = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
#pragma unroll
            for (int y = 0; y < 25; y += 5) A[y + x] ^= d;
        }
        uint64_t cur = A[1];
#pragma unroll
        for (int i = 0; i < 24; i++) {
            int j = pi[i];
            uint64_t t = A[j];
            A[j] = rotl64(cur, rho[i]);
            cur = t;
        }
#pragma unroll
        for (int y = 0; y < 25; y += 5) {
            uint64_t t0 = A[y], t1 = A[y + 1], t2 = A[y + 2], t3 = A[y + 3], t4 = A[y + 4];
            A[y] = t0 ^ (~t1 & t2);
            A[y + 1] = t1 ^ (~t2 & t3);
            A[y + 2] = t2 ^ (~t3 & t4);
            A[y + 3] = t3 ^ (~t4 & t0);
            A[y + 4] = t4 ^ (~t0 & t1);
        }
        A[0] ^= c_keccak_rc[round];
    }
}

// Entry::new (entry.rs:15-27): hashed_username = BigUint::from_bytes_be(keccak256(username)); big_uint_to_fp reduces it mod r.
__device__ fr_t keccak_username_fr(const uint8_t *msg, uint32_t len) {
    uint64_t A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = 0;
    const uint32_t rate = 136;
    uint32_t pos = 0;
    while (len - pos >= rate) {
        for (uint32_t w = 0; w < rate / 8; w++) {
            uint64_t lane = 0;
            for (int b = 0; b < 8; b++) lane |= (uint64_t)msg[pos + 8 * w + b] << (8 * b);
            A[w] ^= lane;
        }
        keccak_f1600(A);
        pos += rate;
    }
    uint32_t rem = len - pos;
    for (uint32_t w = 0; w < rate / 8; w++) {
        uint64_t lane = 0;
        for (int b = 0; b < 8; b++) {
            uint32_t idx = 8 * w + b;
            uint64_t byte = idx < rem ? msg[pos + idx] : 0;
            if (idx == rem) byte ^= 0x01;
            if (idx == rate - 1) byte ^= 0x80;
            lane |= byte << (8 * b);
        }
        A[w] ^= lane;
    }
    keccak_f1600(A);
    // digest = first 32 bytes of the state, read as a big-endian integer
    fr_t v;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint64_t lane = A[3 - i];  // lane 3 holds the least significant 8 digest bytes
        uint32_t lo = (uint32_t)lane, hi = (uint32_t)(lane >> 32);  // byte-swap each half
        v.v[2 * i] = __byte_perm(hi, 0, 0x0123);
        v.v[2 * i + 1] = __byte_perm(lo, 0, 0x0123);
    }
    // v < 2^256 < 6 r
#pragma unroll
    for (int q = 0; q < 5; q++) {
        fr_t t = v;
        final_sub<FrParams>(v.v, t.v);
    }
    return to_mont(v);
}

// ---------------------------------------------------------------------------------------- tree kernels
// flat tree: hash[off(l) + i], bal[c * stride + off(l) + i] with stride = 2^(depth+1)
struct TreeView {
    uint4 *hash;
    uint4 *bal;
    uint4 *uname;  // hashed usernames of the (padded) entries, Montgomery
    uint32_t depth, n_cur;
};

__device__ __forceinline__ void leaf_from_preimage(const TreeView &t, uint64_t i, const fr_t &uname, const fr_t *bals) {
    const uint64_t stride = 2ull << t.depth;
    fr_t s0 = uname, s1 = sponge_init_capacity(t.n_cur + 1);  // state [0 + x_0, capacity]
    poseidon_permute(s0, s1);
    for (uint32_t c = 0; c < t.n_cur; c++) {
        s0 = add(s0, bals[c]);
        poseidon_permute(s0, s1);
    }
    store_fp(t.hash + 2 * i, s0);
    store_fp(t.uname + 2 * i, uname);
    for (uint32_t c = 0; c < t.n_cur; c++) store_fp(t.bal + 2 * (c * stride + i), bals[c]);
}

#define SB_MAX_CUR 32

// entries given as (username bytes, u64 balances); rows >= n_entries are zero entries (entry.rs:30-38)
// a balance as `limbs` little-endian u64 words (1: N_BYTES <= 8, the common case; 4: a 256-bit BigUint, entry.rs / csv/entry_16_bigints.csv), to Fr:
// `big_uint_to_fp` (operation_helpers.rs:10-12) reduces mod r, which to_mont does for any 256-bit input
__device__ __forceinline__ fr_t balance_to_fr(const uint64_t *w, uint32_t limbs) {
    fr_t x = fr_t::zero();
    for (uint32_t l = 0; l < limbs; l++) {
        x.v[2 * l] = (uint32_t)w[l];
        x.v[2 * l + 1] = (uint32_t)(w[l] >> 32);
    }
    return to_mont(x);
}

__global__ void mst_leaf_entries_kernel(TreeView t, const uint8_t *names, const uint32_t *offs, const uint64_t *bal64, uint32_t limbs, uint64_t n_entries) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (1ull << t.depth)) return;
    fr_t uname = fr_t::zero();
    fr_t bals[SB_MAX_CUR];
    if (i < n_entries) {
        uname = keccak_username_fr(names + offs[i], offs[i + 1] - offs[i]);
        for (uint32_t c = 0; c < t.n_cur; c++) bals[c] = balance_to_fr(bal64 + (i * t.n_cur + c) * limbs, limbs);
    } else {
        for (uint32_t c = 0; c < t.n_cur; c++) bals[c] = fr_t::zero();
    }
    leaf_from_preimage(t, i, uname, bals);
}

// leaves given as hash preimages [username, balances...] in Montgomery form (Node::leaf_node_from_preimage, node.rs:57-69)
__global__ void mst_leaf_preimage_kernel(TreeView t, const uint4 *pre) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (1ull << t.depth)) return;
    fr_t bals[SB_MAX_CUR];
    const uint4 *p = pre + 2 * i * (t.n_cur + 1);
    fr_t uname = load_fp<FrParams>(p);
    for (uint32_t c = 0; c < t.n_cur; c++) bals[c] = load_fp<FrParams>(p + 2 * (c + 1));
    leaf_from_preimage(t, i, uname, bals);
}

// Node::middle (node.rs:32-45): balances add per currency, hash = H(balances..., hash_l, hash_r)
__device__ __forceinline__ void middle_node(const TreeView &t, uint32_t level, uint64_t i) {
    const uint64_t stride = 2ull << t.depth;
    const uint64_t src = level_off(t.depth, level - 1) + 2 * i, dst = level_off(t.depth, level) + i;
    fr_t s0 = fr_t::zero(), s1 = sponge_init_capacity(t.n_cur + 2);
    for (uint32_t c = 0; c < t.n_cur; c++) {
        const uint4 *b = t.bal + 2 * (c * stride + src);
        fr_t sum = add(load_fp<FrParams>(b), load_fp<FrParams>(b + 2));
        store_fp(t.bal + 2 * (c * stride + dst), sum);
        s0 = add(s0, sum);
        poseidon_permute(s0, s1);
    }
    s0 = add(s0, load_fp<FrParams>(t.hash + 2 * src));
    poseidon_permute(s0, s1);
    s0 = add(s0, load_fp<FrParams>(t.hash + 2 * (src + 1)));
    poseidon_permute(s0, s1);
    store_fp(t.hash + 2 * dst, s0);
}

__global__ void mst_level_kernel(TreeView t, uint32_t level) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (1ull << (t.depth - level))) return;
    middle_node(t, level, i);
}

// levels first_level..depth (at most 256 nodes in the first of them) in one CTA
__global__ void mst_top_kernel(TreeView t, uint32_t first_level) {
    for (uint32_t level = first_level; level <= t.depth; level++) {
        if (threadIdx.x < (1u << (t.depth - level))) middle_node(t, level, threadIdx.x);
        __threadfence_block();
        __syncthreads();
    }
}

// MerkleProof pieces (tree.rs:85-137) of proof j = blockIdx.y for user index idx[j]; one thread per level.
// out layout per proof (Fr, Montgomery): entry preimage (n_cur+1) | sibling leaf preimage (n_cur+1) | (depth-1) x sibling middle preimage (n_cur+2)
__global__ void mst_proof_kernel(TreeView t, const uint64_t *idx, uint4 *out, uint8_t *path) {
    const uint32_t level = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t j = blockIdx.y;
    if (level >= max(t.depth, 1u)) return;
    const uint64_t stride = 2ull << t.depth;
    const uint64_t index = idx[j];
    const uint64_t per = 2ull * (t.n_cur + 1) + (uint64_t)(t.depth > 0 ? t.depth - 1 : 0) * (t.n_cur + 2);
    uint4 *o = out + 2 * j * per;
    if (level == 0) {
        for (int side = 0; side < 2; side++) {
            if (side == 1 && t.depth == 0) break;  // a single-leaf tree has no sibling
            const uint64_t e = side == 0 ? index : (index ^ 1);
            uint4 *q = o + 2 * side * (t.n_cur + 1);
            store_fp(q, load_fp<FrParams>(t.uname + 2 * e));
            for (uint32_t c = 0; c < t.n_cur; c++) store_fp(q + 2 * (c + 1), load_fp<FrParams>(t.bal + 2 * (c * stride + e)));
        }
        if (t.depth > 0) path[j * t.depth] = (uint8_t)(index & 1);
        return;
    }
    const uint64_t cur = index >> level;
    const uint64_t sib = cur ^ 1;
    path[j * t.depth + level] = (uint8_t)(cur & 1);
    // get_middle_node_hash_preimage(level, sib): children at level-1
    const uint64_t ch = level_off(t.depth, level - 1) + 2 * sib;
    uint4 *q = o + 2 * (2ull * (t.n_cur + 1) + (uint64_t)(level - 1) * (t.n_cur + 2));
    for (uint32_t c = 0; c < t.n_cur; c++) {
        const uint4 *b = t.bal + 2 * (c * stride + ch);
        store_fp(q + 2 * c, add(load_fp<FrParams>(b), load_fp<FrParams>(b + 2)));
    }
    store_fp(q + 2 * t.n_cur, load_fp<FrParams>(t.hash + 2 * ch));
    store_fp(q + 2 * (t.n_cur + 1), load_fp<FrParams>(t.hash + 2 * (ch + 1)));
}


// MerkleSumTree::update_leaf (mst.rs:158-197): new balances for one entry, then the path to the root; one thread (depth + 1 dependent hashes)
__global__ void mst_update_kernel(TreeView t, uint64_t index, const uint64_t *bal64, uint32_t limbs) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    fr_t bals[SB_MAX_CUR];
    for (uint32_t c = 0; c < t.n_cur; c++) bals[c] = balance_to_fr(bal64 + c * limbs, limbs);
    leaf_from_preimage(t, index, load_fp<FrParams>(t.uname + 2 * index), bals);
    for (uint32_t level = 1; level <= t.depth; level++) {
        __threadfence();
        middle_node(t, level, index >> level);
    }
}

// Tree::verify_proof (tree.rs:139-186), one thread per proof.  pre: the layout sb_mst_proofs writes.
__global__ void mst_verify_kernel(const uint4 *pre, const uint8_t *path, uint32_t depth, uint32_t n_cur, const uint4 *root /* hash | balances */, uint64_t n_proofs, uint8_t *ok) {
    const uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n_proofs) return;
    const uint64_t per = 2ull * (n_cur + 1) + (uint64_t)(depth > 0 ? depth - 1 : 0) * (n_cur + 2);
    const uint4 *p = pre + 2 * j * per;
    auto hash_n = [&](const uint4 *q, uint32_t len) {
        fr_t s0 = fr_t::zero(), s1 = sponge_init_capacity(len);
        for (uint32_t i = 0; i < len; i++) {
            s0 = add(s0, load_fp<FrParams>(q + 2 * i));
            poseidon_permute(s0, s1);
        }
        return s0;
    };
    fr_t bal[SB_MAX_CUR];
    fr_t node = hash_n(p, n_cur + 1);
    for (uint32_t c = 0; c < n_cur; c++) bal[c] = load_fp<FrParams>(p + 2 * (c + 1));
    for (uint32_t level = 0; level < depth; level++) {
        const uint4 *sp = level == 0 ? p + 2 * (n_cur + 1) : p + 2 * (2ull * (n_cur + 1) + (uint64_t)(level - 1) * (n_cur + 2));
        const uint32_t slen = level == 0 ? n_cur + 1 : n_cur + 2;
        const fr_t sib = hash_n(sp, slen);
        const uint4 *sbal = level == 0 ? sp + 2 : sp;  // leaf preimage: [username, balances...]; middle preimage: [balances..., hash_l, hash_r]
        fr_t s0 = fr_t::zero(), s1 = sponge_init_capacity(n_cur + 2);
        for (uint32_t c = 0; c < n_cur; c++) {
            bal[c] = add(bal[c], load_fp<FrParams>(sbal + 2 * c));
            s0 = add(s0, bal[c]);
            poseidon_permute(s0, s1);
        }
        const bool right = path[j * depth + level] != 0;
        s0 = add(s0, right ? sib : node);
        poseidon_permute(s0, s1);
        s0 = add(s0, right ? node : sib);
        poseidon_permute(s0, s1);
        node = s0;
    }
    bool good = node == load_fp<FrParams>(root);
    for (uint32_t c = 0; c < n_cur; c++) good = good && (bal[c] == load_fp<FrParams>(root + 2 * (c + 1)));
    ok[j] = good ? 1 : 0;
}

}  // namespace sb

using namespace sb;

struct sb_mst {
    sb_ctx *ctx = nullptr;
    uint32_t depth = 0, n_cur = 0;
    void *d_hash = nullptr, *d_bal = nullptr, *d_uname = nullptr;
    float build_ms = 0;
    TreeView view() const {
        TreeView t;
        t.hash = (uint4 *)d_hash;
        t.bal = (uint4 *)d_bal;
        t.uname = (uint4 *)d_uname;
        t.depth = depth;
        t.n_cur = n_cur;
        return t;
    }
};

static int32_t mst_alloc(sb_ctx *ctx, uint32_t depth, uint32_t n_cur, sb_mst **out) {
    sb_mst *m = new sb_mst();
    m->ctx = ctx;
    ctx_retain(ctx);  // the tree keeps its context alive: sb_ctx_destroy before sb_mst_destroy is legal
    m->depth = depth;
    m->n_cur = n_cur;
    const size_t slots = 2ull << depth;
    cudaError_t e = cudaMalloc(&m->d_hash, slots * 32);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_bal, slots * 32 * n_cur);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_uname, (slots / 2) * 32);
    if (e != cudaSuccess) {
        set_last_error("sb_mst: cudaMalloc of a depth-%u tree with %u currencies: %s", depth, n_cur, cudaGetErrorString(e));
        cudaFree(m->d_hash);
        cudaFree(m->d_bal);
        cudaFree(m->d_uname);
        delete m;
        ctx_release(ctx);
        return SB_ERR_ALLOC;
    }
    *out = m;
    return SB_OK;
}

static int32_t mst_build_levels(sb_ctx *ctx, sb_mst *m, cudaStream_t st) {
    TreeView t = m->view();
    uint32_t level = 1;
    for (; level <= m->depth && (1ull << (m->depth - level)) > 256; level++) {
        const uint64_t cnt = 1ull << (m->depth - level);
        SB_LAUNCH(ctx, mst_level_kernel, (unsigned)((cnt + 127) / 128), 128, 0, st, t, level);
    }
    if (level <= m->depth) SB_LAUNCH(ctx, mst_top_kernel, 1, 256, 0, st, t, level);
    return SB_OK;
}

static uint32_t depth_for(size_t n_entries) {  // mst.rs:106: ceil(log2(len))
    uint32_t d = 0;
    while ((1ull << d) < n_entries) d++;
    return d;
}

extern "C" {

static int32_t mst_build_impl(sb_ctx *ctx, const uint8_t *usernames, const uint32_t *offsets, const uint64_t *balances, uint32_t limbs, size_t n_entries,
                              uint32_t n_currencies, sb_mst **out_mst) {
    if (!ctx || !out_mst) return SB_ERR_ARG;
    SB_REQUIRE(n_entries >= 1 && usernames && offsets && balances, "sb_mst_build: empty input");
    SB_REQUIRE(n_currencies >= 1 && n_currencies <= SB_MAX_CUR, "sb_mst_build: n_currencies must be 1..32");
    SB_REQUIRE(n_entries <= (1ull << 30), "sb_mst_build: more than 2^30 entries");
    // the leaf kernel reads names[offs[i] .. offs[i+1]): a non-monotone array would make that length wrap and fault the device
    SB_REQUIRE(offsets[0] == 0, "sb_mst_build: offsets[0] must be 0");
    for (size_t i = 0; i < n_entries; i++) SB_REQUIRE(offsets[i] <= offsets[i + 1], "sb_mst_build: offsets must be non-decreasing");
    CtxGuard g(ctx);
    SB_TRY(poseidon_consts_load(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t depth = depth_for(n_entries);
    sb_mst *m = nullptr;
    SB_TRY(mst_alloc(ctx, depth, n_currencies, &m));
    const size_t name_bytes = offsets[n_entries];
    void *d_names = nullptr, *d_offs = nullptr, *d_b64 = nullptr;
    int32_t rc = scratch_get(ctx, "mst_names", name_bytes + 16, &d_names);
    if (rc == SB_OK) rc = scratch_get(ctx, "mst_offs", (n_entries + 1) * 4, &d_offs);
    if (rc == SB_OK) rc = scratch_get(ctx, "mst_b64", n_entries * n_currencies * 8 * limbs, &d_b64);
    if (rc != SB_OK) { sb_mst_destroy(m); return rc; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto body = [&]() -> int32_t {
        SB_CUDA_TRY(cudaEventRecord(e0, st));
        SB_CUDA_TRY(cudaMemcpyAsync(d_names, usernames, name_bytes, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaMemcpyAsync(d_offs, offsets, (n_entries + 1) * 4, cudaMemcpyHostToDevice, st));
        SB_CUDA_TRY(cudaMemcpyAsync(d_b64, balances, n_entries * n_currencies * 8 * limbs, cudaMemcpyHostToDevice, st));
        const uint64_t leaves = 1ull << depth;
        SB_LAUNCH(ctx, mst_leaf_entries_kernel, (unsigned)((leaves + 127) / 128), 128, 0, st, m->view(), (const uint8_t *)d_names, (const uint32_t *)d_offs,
                  (const uint64_t *)d_b64, limbs, (uint64_t)n_entries);
        SB_TRY(mst_build_levels(ctx, m, st));
        SB_CUDA_TRY(cudaEventRecord(e1, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
        SB_CUDA_TRY(cudaEventElapsedTime(&m->build_ms, e0, e1));
        return SB_OK;
    };
    rc = body();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != SB_OK) { sb_mst_destroy(m); return rc; }
    *out_mst = m;
    return SB_OK;
}

int32_t sb_mst_build(sb_ctx *ctx, const uint8_t *usernames, const uint32_t *offsets, const uint64_t *balances, size_t n_entries, uint32_t n_currencies,
                     sb_mst **out_mst) {
    return mst_build_impl(ctx, usernames, offsets, balances, 1, n_entries, n_currencies, out_mst);
}
int32_t sb_mst_build_wide(sb_ctx *ctx, const uint8_t *usernames, const uint32_t *offsets, const uint8_t *balances_le32, size_t n_entries, uint32_t n_currencies,
                          sb_mst **out_mst) {
    return mst_build_impl(ctx, usernames, offsets, (const uint64_t *)balances_le32, 4, n_entries, n_currencies, out_mst);
}

int32_t sb_mst_build_from_preimages(sb_ctx *ctx, const uint8_t *leaf_preimages, size_t n_leaves, uint32_t n_currencies, sb_mst **out_mst) {
    if (!ctx || !out_mst) return SB_ERR_ARG;
    SB_REQUIRE(leaf_preimages && n_leaves >= 1 && (n_leaves & (n_leaves - 1)) == 0, "sb_mst_build_from_preimages: the leaf layer must be a power of two (build_tree.rs:17)");
    SB_REQUIRE(n_currencies >= 1 && n_currencies <= SB_MAX_CUR, "sb_mst_build_from_preimages: n_currencies must be 1..32");
    SB_REQUIRE(n_leaves <= (1ull << 30), "sb_mst_build_from_preimages: more than 2^30 leaves");
    CtxGuard g(ctx);
    SB_TRY(poseidon_consts_load(ctx->device));
    cudaStream_t st = ctx->stream;
    const uint32_t depth = depth_for(n_leaves);
    sb_mst *m = nullptr;
    SB_TRY(mst_alloc(ctx, depth, n_currencies, &m));
    void *d_pre = nullptr;
    const size_t bytes = n_leaves * (n_currencies + 1) * 32;
    int32_t rc = scratch_get(ctx, "mst_pre", bytes, &d_pre);
    if (rc != SB_OK) { sb_mst_destroy(m); return rc; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto body = [&]() -> int32_t {
        SB_CUDA_TRY(cudaEventRecord(e0, st));
        SB_CUDA_TRY(cudaMemcpyAsync(d_pre, leaf_preimages, bytes, cudaMemcpyHostToDevice, st));
        SB_LAUNCH(ctx, mst_leaf_preimage_kernel, (unsigned)((n_leaves + 127) / 128), 128, 0, st, m->view(), (const uint4 *)d_pre);
        SB_TRY(mst_build_levels(ctx, m, st));
        SB_CUDA_TRY(cudaEventRecord(e1, st));
        SB_CUDA_TRY(cudaStreamSynchronize(st));
        SB_CUDA_TRY(cudaEventElapsedTime(&m->build_ms, e0, e1));
        return SB_OK;
    };
    rc = body();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != SB_OK) { sb_mst_destroy(m); return rc; }
    *out_mst = m;
    return SB_OK;
}

int32_t sb_mst_destroy(sb_mst *mst) {
    if (!mst) return SB_OK;
    {
        CtxGuard g(mst->ctx);
        cudaStreamSynchronize(mst->ctx->stream);
        cudaFree(mst->d_hash);
        cudaFree(mst->d_bal);
        cudaFree(mst->d_uname);
    }
    ctx_release(mst->ctx);
    delete mst;
    return SB_OK;
}

int32_t sb_mst_shape(const sb_mst *mst, uint32_t *out_depth, uint32_t *out_n_currencies, float *out_build_ms) {
    if (!mst) return SB_ERR_ARG;
    if (out_depth) *out_depth = mst->depth;
    if (out_n_currencies) *out_n_currencies = mst->n_cur;
    if (out_build_ms) *out_build_ms = mst->build_ms;
    return SB_OK;
}

int32_t sb_mst_node(const sb_mst *mst, uint32_t level, size_t index, uint8_t out_hash[32], uint8_t *out_balances) {
    if (!mst || !out_hash || !out_balances) return SB_ERR_ARG;
    SB_REQUIRE(level <= mst->depth && index < (1ull << (mst->depth - level)), "sb_mst_node: node not found (tree.rs:35-38)");
    sb_ctx *ctx = mst->ctx;
    CtxGuard g(ctx);
    const uint64_t pos = level_off(mst->depth, level) + index, stride = 2ull << mst->depth;
    SB_CUDA_TRY(cudaMemcpyAsync(out_hash, (const uint8_t *)mst->d_hash + pos * 32, 32, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaMemcpy2DAsync(out_balances, 32, (const uint8_t *)mst->d_bal + pos * 32, stride * 32, 32, mst->n_cur, cudaMemcpyDeviceToHost, ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int32_t sb_mst_root(const sb_mst *mst, uint8_t out_hash[32], uint8_t *out_balances) {
    if (!mst) return SB_ERR_ARG;
    return sb_mst_node(mst, mst->depth, 0, out_hash, out_balances);
}

int32_t sb_mst_level_hashes(const sb_mst *mst, uint32_t level, uint8_t *out_hashes) {
    if (!mst || !out_hashes) return SB_ERR_ARG;
    SB_REQUIRE(level <= mst->depth, "sb_mst_level_hashes: invalid level");
    sb_ctx *ctx = mst->ctx;
    CtxGuard g(ctx);
    SB_CUDA_TRY(cudaMemcpyAsync(out_hashes, (const uint8_t *)mst->d_hash + level_off(mst->depth, level) * 32, (32ull << (mst->depth - level)), cudaMemcpyDeviceToHost,
                                ctx->stream));
    SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return SB_OK;
}

int32_t sb_mst_proofs(const sb_mst *mst, const uint64_t *indices, size_t n_proofs, uint8_t *out_preimages, uint8_t *out_path_indices) {
    if (!mst || !indices || !out_preimages || !out_path_indices) return SB_ERR_ARG;
    if (n_proofs == 0) return SB_OK;
    for (size_t j = 0; j < n_proofs; j++) SB_REQUIRE(indices[j] < (1ull << mst->depth), "sb_mst_proofs: index out of bounds (tree.rs:99-101)");
    sb_ctx *ctx = mst->ctx;
    CtxGuard g(ctx);
    cudaStream_t st = ctx->stream;
    const uint32_t d = mst->depth;
    const size_t per = 2ull * (mst->n_cur + 1) + (size_t)(d > 0 ? d - 1 : 0) * (mst->n_cur + 2);
    void *d_idx = nullptr, *d_out = nullptr, *d_path = nullptr;
    SB_TRY(scratch_get(ctx, "mst_pidx", n_proofs * 8, &d_idx));
    SB_TRY(scratch_get(ctx, "mst_pout", n_proofs * per * 32, &d_out));
    SB_TRY(scratch_get(ctx, "mst_ppath", n_proofs * (d ? d : 1), &d_path));
    SB_CUDA_TRY(cudaMemcpyAsync(d_idx, indices, n_proofs * 8, cudaMemcpyHostToDevice, st));
    SB_CUDA_TRY(cudaMemsetAsync(d_out, 0, n_proofs * per * 32, st));
    const size_t max_y = 65535;
    for (size_t j0 = 0; j0 < n_proofs; j0 += max_y) {
        const size_t cnt = n_proofs - j0 < max_y ? n_proofs - j0 : max_y;
        SB_LAUNCH(ctx, mst_proof_kernel, dim3(1, (unsigned)cnt), 32, 0, st, mst->view(), (const uint64_t *)d_idx + j0, (uint4 *)d_out + 2 * j0 * per,
                  (uint8_t *)d_path + j0 * d);
    }
    SB_CUDA_TRY(cudaMemcpyAsync(out_preimages, d_out, n_proofs * per * 32, cudaMemcpyDeviceToHost, st));
    if (d) SB_CUDA_TRY(cudaMemcpyAsync(out_path_indices, d_path, n_proofs * d, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(cudaStreamSynchronize(st));
    return SB_OK;
}

static int32_t mst_update_impl(sb_mst *mst, size_t index, const uint64_t *new_balances, uint32_t limbs, uint8_t out_root_hash[32], uint8_t *out_root_balances) {
    if (!mst || !new_balances) return SB_ERR_ARG;
    SB_REQUIRE(index < (1ull << mst->depth), "sb_mst_update_leaf: index out of bounds");
    sb_ctx *ctx = mst->ctx;
    {
        CtxGuard g(ctx);
        void *d_b;
        SB_TRY(scratch_get(ctx, "mst_upd", (size_t)mst->n_cur * 8 * limbs, &d_b));
        SB_TRY(h2d_staged(ctx, d_b, new_balances, (size_t)mst->n_cur * 8 * limbs, ctx->stream));
        SB_LAUNCH(ctx, mst_update_kernel, 1, 32, 0, ctx->stream, mst->view(), (uint64_t)index, (const uint64_t *)d_b, limbs);
        SB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    if (out_root_hash && out_root_balances) return sb_mst_root(mst, out_root_hash, out_root_balances);
    return SB_OK;
}
int32_t sb_mst_update_leaf(sb_mst *mst, size_t index, const uint64_t *new_balances, uint8_t out_root_hash[32], uint8_t *out_root_balances) {
    return mst_update_impl(mst, index, new_balances, 1, out_root_hash, out_root_balances);
}
int32_t sb_mst_update_leaf_wide(sb_mst *mst, size_t index, const uint8_t *new_balances_le32, uint8_t out_root_hash[32], uint8_t *out_root_balances) {
    return mst_update_impl(mst, index, (const uint64_t *)new_balances_le32, 4, out_root_hash, out_root_balances);
}

int32_t sb_mst_verify_proofs(sb_ctx *ctx, uint32_t n_currencies, uint32_t depth, const uint8_t *preimages, const uint8_t *path_indices, const uint8_t root_hash[32],
                             const uint8_t *root_balances, size_t n_proofs, uint8_t *out_ok) {
    if (!ctx || !preimages || !path_indices || !root_hash || !root_balances || !out_ok) return SB_ERR_ARG;
    SB_REQUIRE(n_currencies >= 1 && n_currencies <= SB_MAX_CUR && depth <= 30, "sb_mst_verify_proofs: bad shape");
    if (n_proofs == 0) return SB_OK;
    CtxGuard g(ctx);
    SB_TRY(poseidon_consts_load(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t per = 2ull * (n_currencies + 1) + (size_t)(depth > 0 ? depth - 1 : 0) * (n_currencies + 2);
    void *d_pre, *d_path, *d_root, *d_ok;
    SB_TRY(scratch_get(ctx, "mst_vpre", n_proofs * per * 32, &d_pre));
    SB_TRY(scratch_get(ctx, "mst_vpath", n_proofs * (depth ? depth : 1), &d_path));
    SB_TRY(scratch_get(ctx, "mst_vroot", (n_currencies + 1) * 32, &d_root));
    SB_TRY(scratch_get(ctx, "mst_vok", n_proofs, &d_ok));
    SB_CUDA_TRY(cudaMemcpyAsync(d_pre, preimages, n_proofs * per * 32, cudaMemcpyHostToDevice, st));
    if (depth) SB_CUDA_TRY(cudaMemcpyAsync(d_path, path_indices, n_proofs * depth, cudaMemcpyHostToDevice, st));
    SB_CUDA_TRY(cudaMemcpyAsync(d_root, root_hash, 32, cudaMemcpyHostToDevice, st));
    SB_CUDA_TRY(cudaMemcpyAsync((uint8_t *)d_root + 32, root_balances, n_currencies * 32, cudaMemcpyHostToDevice, st));
    SB_LAUNCH(ctx, mst_verify_kernel, (unsigned)((n_proofs + 63) / 64), 64, 0, st, (const uint4 *)d_pre, (const uint8_t *)d_path, depth, n_currencies, (const uint4 *)d_root,
              (uint64_t)n_proofs, (uint8_t *)d_ok);
    SB_CUDA_TRY(cudaMemcpyAsync(out_ok, d_ok, n_proofs, cudaMemcpyDeviceToHost, st));
    SB_CUDA_TRY(cudaStreamSynchronize(st));
    return SB_OK;
}

}  // extern "C"
