// Host-side harness of the NTT: compiles the DEVICE tile code (csrc/ntt_core.cuh: window / index maps, radix-8 stage butterflies, fused
// pre / post operations, tile coordinates of every pass kind) with g++ and runs it for every thread of every tile, phase by phase (a barrier
// of the kernel = a loop boundary here), following exactly the phase sequence of ntt_pass_kernel in csrc/ntt.cu.  tests/test_host_logic.py
// compares the result with the oracle's best_fft, and checks that no shared-memory access pattern of the kernel has bank conflicts.
#include <vector_types.h>

#include <cstddef>
#include <cstring>
#include <vector>

#include "../../circuits_halo2_b200/csrc/ntt_core.cuh"
using namespace sb;

namespace {
nfr_t fpow(nfr_t b, uint64_t e) {
    nfr_t acc = nfr_t::one();
    while (e) {
        if (e & 1) acc = mul(acc, b);
        b = sqr(b);
        e >>= 1;
    }
    return acc;
}
// same policy as ntt_make_plan in ntt.cu
void make_plan(uint32_t log_n, uint32_t tile_log, uint32_t min_passes, int *npass, uint32_t *radix) {
    if (log_n <= tile_log) { *npass = 1; radix[0] = log_n; return; }
    uint32_t p = (log_n + tile_log - 1) / tile_log;
    if (min_passes > p) p = min_passes;
    const uint32_t base = log_n / p, extra = log_n % p;
    *npass = (int)p;
    for (uint32_t t = 0; t < p; t++) radix[t] = base + (t < extra ? 1 : 0);
}
struct HostTwiddle {
    const std::vector<nfr_t> *w;
    std::vector<uint32_t> *log;  // swizzled slots touched, in call order (bank-conflict audit)
    nfr_t operator()(uint32_t e) const {
        if (log) log->push_back(ntt_swz(e));
        return (*w)[e];
    }
};
// worst number of distinct 16-byte bank groups' collisions inside any quarter-warp (8 consecutive threads) for one access slot
uint32_t conflict_degree(const std::vector<std::vector<uint32_t>> &per_thread, size_t slot) {
    uint32_t worst = 1;
    for (size_t t0 = 0; t0 + 8 <= per_thread.size(); t0 += 8) {
        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        std::vector<uint32_t> seen;
        for (size_t t = t0; t < t0 + 8; t++) {
            if (slot >= per_thread[t].size()) continue;
            const uint32_t a = per_thread[t][slot];
            bool dup = false;
            for (uint32_t s : seen) dup = dup || s == a;   // same address = broadcast, not a conflict
            if (dup) continue;
            seen.push_back(a);
            cnt[a & 7]++;
        }
        for (int i = 0; i < 8; i++) worst = cnt[i] > worst ? cnt[i] : worst;
    }
    return worst;
}
}  // namespace

extern "C" {
// a: n = 2^log_n elements (8 x u32 each, Montgomery) transformed in place like sb_best_fft; fused operations as in NttFuse (null = off).
// Returns the worst shared-memory bank-conflict degree seen (1 = conflict-free), or 0 on a usage error.
uint32_t ht_ntt(uint32_t *a_words, const uint32_t *omega_words, uint32_t log_n, uint32_t tile_log, uint32_t min_passes, int use_full_tw, const uint32_t *scale_words,
                const uint32_t *pre_vec_words, const uint32_t *pre_pat_words, uint32_t pre_m, const uint32_t *post_pat_words, uint32_t post_m, uint64_t n_in, uint64_t n_out, const uint32_t *post_vec_words, uint32_t eb) {
    const uint64_t n = 1ull << log_n;
    if (log_n < 3) return 0;
    nfr_t omega, scale = nfr_t::one();
    memcpy(omega.v, omega_words, 32);
    if (scale_words) memcpy(scale.v, scale_words, 32);
    int npass;
    uint32_t radix[NTT_MAX_PASS];
    make_plan(log_n, tile_log, min_passes, &npass, radix);
    std::vector<uint4> in(2 * n), work(2 * n), out(2 * n);
    memcpy(in.data(), a_words, 32 * n);
    memset(out.data(), 0xAB, 32 * n);  // truncated outputs must stay untouched
    const uint32_t log_tlo = (log_n + 1) / 2;
    std::vector<uint4> t_lo(2ull << log_tlo), t_hi(2ull << (log_n - log_tlo));
    {
        nfr_t cur = nfr_t::one();
        for (uint64_t i = 0; i < (1ull << log_tlo); i++) { ntt_store_fr(&t_lo[2 * i], cur); cur = mul(cur, omega); }
        const nfr_t step = fpow(omega, 1ull << log_tlo);
        cur = nfr_t::one();
        for (uint64_t i = 0; i < (1ull << (log_n - log_tlo)); i++) { ntt_store_fr(&t_hi[2 * i], cur); cur = mul(cur, step); }
    }
    uint32_t worst = 1;
    uint32_t log_a = 0;
    for (int t = 0; t < npass; t++) {
        NttPassArgs p;
        memset(&p, 0, sizeof p);
        const bool last = t == npass - 1;
        p.log_n = log_n; p.r = radix[t]; p.log_a = log_a; p.log_c = log_n - log_a - p.r; p.log_r1 = radix[0]; p.log_tlo = log_tlo; p.npass = (uint32_t)npass;
        p.kind = npass == 1 ? NTT_SINGLE : (last ? NTT_LAST : NTT_STRIDED);
        p.g = npass == 1 ? 0 : tile_log - p.r;
        if (p.kind == NTT_STRIDED && p.g > p.log_c) p.g = p.log_c;
        if (p.kind == NTT_LAST && p.g > p.log_r1) p.g = p.log_r1;
        for (int m = 1; m < npass - 1; m++) p.mid_bits[p.n_mid++] = radix[m];
        p.t_lo = t_lo.data(); p.t_hi = t_hi.data();
        p.n_in = n; p.n_out = n;
        const bool scale_in_post = scale_words && npass == 1;
        if (t == 0) {
            p.pre_vec = (const uint4 *)pre_vec_words;
            p.pre_m = pre_m;
            for (uint32_t i = 0; i < pre_m; i++) memcpy(p.pre_pat[i].v, pre_pat_words + 8 * i, 32);
            if (n_in) p.n_in = n_in;
        }
        if (last) {
            if (n_out) p.n_out = n_out;
            p.post_m = post_m;
            p.post_vec = (const uint4 *)post_vec_words;
            for (uint32_t i = 0; i < post_m; i++) { memcpy(p.post_pat[i].v, post_pat_words + 8 * i, 32); if (scale_in_post) p.post_pat[i] = mul(p.post_pat[i], scale); }
            if (scale_in_post && post_m == 0) { p.post_m = 1; p.post_pat[0] = scale; }
        }
        // butterfly twiddles omega_R^e
        const uint32_t half_r = 1u << (p.r - 1);
        std::vector<nfr_t> w(half_r);
        {
            const nfr_t wr = fpow(omega, 1ull << (log_n - p.r));
            nfr_t cur = nfr_t::one();
            for (uint32_t e = 0; e < half_r; e++) { w[e] = cur; cur = mul(cur, wr); }
        }
        std::vector<uint4> full;
        if (!last && use_full_tw) {
            const uint64_t cnt = 1ull << (log_n - log_a);
            full.resize(2 * cnt);
            for (uint64_t i = 0; i < cnt; i++) {
                const uint64_t c = i & ((1ull << p.log_c) - 1), k = i >> p.log_c;
                nfr_t tw = fpow(omega, (c * k) << log_a);
                if (t == 0 && scale_words) tw = mul(tw, scale);
                ntt_store_fr(&full[2 * i], tw);
            }
            p.tw_full = full.data();
        } else if (!last && t == 0 && scale_words) {
            p.has_tw_scale = 1;
            p.tw_scale = scale;
        }
        p.src = npass == 1 ? in.data() : (t == 0 ? in.data() : work.data());
        p.dst = npass == 1 ? out.data() : (last ? out.data() : work.data());

        p.eb = eb;
        const NttGeom G(p);
        const NttBatch bo(p, 0);
        const uint32_t E = 1u << G.eb;
        const uint32_t tile = 1u << G.t, T = tile >> G.eb;
        const uint64_t tiles = 1ull << (log_n - p.r - p.g);
        std::vector<nfr_t> smem(tile);
        std::vector<nfr_t> regs((size_t)T * E);
        for (uint64_t tile_id = 0; tile_id < tiles; tile_id++) {
            const NttTileCoord tc(p, tile_id);
            const bool audit = tile_id == 0;  // the access pattern is the same for every tile
            uint32_t pw = G.window(0);
            // ---- load phase ----
            if (p.kind == NTT_LAST) {
                for (uint32_t i = 0; i < tile; i++) smem[ntt_swz(i)] = ntt_load_fr(p.src + 2 * tc.in_index(p, G.j_of(i), G.gg_of(i)));
                for (uint32_t tid = 0; tid < T; tid++)
                    for (uint32_t b = 0; b < E; b++) regs[tid * E + b] = smem[ntt_swz(G.idx(tid, pw, b))];
            } else {
                for (uint32_t tid = 0; tid < T; tid++)
                    for (uint32_t b = 0; b < E; b++) {
                        const uint32_t i = G.idx(tid, pw, b);
                        const uint64_t gi = tc.in_index(p, G.j_of(i), G.gg_of(i));
                        regs[tid * E + b] = p.log_a == 0 ? ntt_fetch_input(p, bo, gi) : ntt_load_fr(p.src + 2 * gi);
                    }
            }
            uint32_t low = G.r, prev = pw;
            for (uint32_t s = 0; s < G.n_stages; s++) {
                pw = G.window(s);
                if (s > 0) {
                    std::vector<std::vector<uint32_t>> wr(T), rd(T);
                    for (uint32_t tid = 0; tid < T; tid++)
                        for (uint32_t b = 0; b < E; b++) { const uint32_t a = ntt_swz(G.idx(tid, prev, b)); smem[a] = regs[tid * E + b]; wr[tid].push_back(a); }
                    for (uint32_t tid = 0; tid < T; tid++)
                        for (uint32_t b = 0; b < E; b++) { const uint32_t a = ntt_swz(G.idx(tid, pw, b)); regs[tid * E + b] = smem[a]; rd[tid].push_back(a); }
                    if (audit)
                        for (size_t slot = 0; slot < E; slot++) { const uint32_t c1 = conflict_degree(wr, slot), c2 = conflict_degree(rd, slot); worst = c1 > worst ? c1 : worst; worst = c2 > worst ? c2 : worst; }
                }
                std::vector<std::vector<uint32_t>> twlog(T);
                for (uint32_t tid = 0; tid < T; tid++) {
                    HostTwiddle tw{&w, audit ? &twlog[tid] : nullptr};
                    if (G.eb == 2) ntt_stage_butterflies<2>(&regs[tid * E], G, tid, pw, low, tw);
                    else ntt_stage_butterflies<3>(&regs[tid * E], G, tid, pw, low, tw);
                }
                if (audit)
                    for (size_t slot = 0; slot < 12; slot++) { const uint32_t c1 = conflict_degree(twlog, slot); worst = c1 > worst ? c1 : worst; }
                low = pw - G.jshift;
                prev = pw;
            }
            // ---- write back ----
            if (p.kind == NTT_LAST && p.g > 0) {
                std::vector<std::vector<uint32_t>> rd(T);
                for (uint32_t tid = 0; tid < T; tid++)
                    for (uint32_t b = 0; b < E; b++) smem[ntt_swz(G.idx(tid, prev, b))] = regs[tid * E + b];
                for (uint32_t tid = 0; tid < T; tid++)
                    for (uint32_t b = 0; b < E; b++) {
                        const uint32_t m = tid + T * b;
                        const uint32_t gg = m & ((1u << p.g) - 1u), jj = m >> p.g;
                        const uint32_t a = ntt_swz((gg << G.r) | jj);
                        rd[tid].push_back(a);
                        ntt_emit(p, bo, tc, ntt_brev(jj, G.r), gg, smem[a]);
                    }
                if (audit)
                    for (size_t slot = 0; slot < E; slot++) { const uint32_t c1 = conflict_degree(rd, slot); worst = c1 > worst ? c1 : worst; }
            } else {
                for (uint32_t tid = 0; tid < T; tid++)
                    for (uint32_t b = 0; b < E; b++) {
                        const uint32_t i = G.idx(tid, prev, b);
                        ntt_emit(p, bo, tc, ntt_brev(G.j_of(i), G.r), G.gg_of(i), regs[tid * E + b]);
                    }
            }
        }
        log_a += p.r;
    }
    memcpy(a_words, out.data(), 32 * n);
    return worst;
}
}
