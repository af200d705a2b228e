// Column-wise evaluation of PLONKish expressions on the GPU: halo2's `GraphEvaluator`
// re-designed as a compile-then-interpret pipeline.
//
// Host: every term (gate polynomial, permutation / lookup identity) is hash-consed into a DAG,
// linearised, and register-allocated onto a handful of value slots; terms are folded Horner-style
// with the challenge y (acc = acc * y + term), so the whole quotient numerator h(X) of
// `Evaluator::evaluate_h` -- 19 gate polynomials + 5 permutation + 5 lookup terms for the Summa
// circuit -- is ONE program and ONE pass over the extended domain: each of the ~30 extended columns
// is streamed from HBM once per rotation it is queried at, nothing intermediate is written back.
// Device: one thread per row; value slots live in shared memory ([slot][thread], 128-bit accesses,
// conflict-free), instructions and constants are warp-uniform loads, column reads are coalesced
// 32-byte loads at (row + rotation) mod n.  The only code is one Montgomery product and one add/sub,
// so the kernel body stays resident in the instruction cache.
#include <algorithm>
#include <array>
#include <map>
#include <tuple>

#include "prover.h"

namespace sb {

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_NEG = 3, OP_COPY = 4 };
enum { K_REG = 0, K_CONST = 1, K_INPUT = 2 };
static inline uint32_t operand(uint32_t kind, uint32_t idx) { return (kind << 30) | idx; }

ExprP e_const(const fr_t &c) { auto e = std::make_shared<Expr>(); e->kind = Expr::CONST; e->c = c; e->col = e->rot = 0; return e; }
ExprP e_col(int col, int rot) { auto e = std::make_shared<Expr>(); e->kind = Expr::COL; e->col = col; e->rot = rot; e->c = fr_t::zero(); return e; }
static ExprP mk(Expr::Kind k, ExprP a, ExprP b) { auto e = std::make_shared<Expr>(); e->kind = k; e->a = a; e->b = b; e->col = e->rot = 0; e->c = fr_t::zero(); return e; }
ExprP e_neg(ExprP a) { return mk(Expr::NEG, a, nullptr); }
ExprP e_add(ExprP a, ExprP b) { return mk(Expr::ADD, a, b); }
ExprP e_sub(ExprP a, ExprP b) { return mk(Expr::SUB, a, b); }
ExprP e_mul(ExprP a, ExprP b) { return mk(Expr::MUL, a, b); }

namespace {

struct Node {
    int op;            // OP_* for interior nodes, -1 for leaves
    uint32_t leaf;     // operand encoding for leaves
    int a, b;          // children (node ids), -1 if unused
    int uses = 0;
    int slot = -1;
    int nmul = 0;      // products in the subtree (0: an addition chain over leaves, cheap to recompute)
};

struct Builder {
    Program &p;
    std::map<std::array<uint32_t, 8>, uint32_t> const_ids;
    std::map<std::pair<int, int>, uint32_t> input_ids;
    std::vector<Node> nodes;
    std::map<std::tuple<int, int, int, int>, int> interior;  // (op, a, b, scope) -> node id; scope = -1 shared by all terms, else the term
    std::map<uint32_t, int> leaves;                           // leaf operand -> node id
    int cur_term = 0;
    explicit Builder(Program &prog) : p(prog) {}

    uint32_t const_id(const fr_t &c) {
        std::array<uint32_t, 8> key;
        for (int i = 0; i < 8; i++) key[i] = c.v[i];
        auto it = const_ids.find(key);
        if (it != const_ids.end()) return it->second;
        uint32_t id = (uint32_t)p.consts.size();
        p.consts.push_back(c);
        const_ids[key] = id;
        return id;
    }
    uint32_t input_id(int col, int rot) {
        auto key = std::make_pair(col, rot);
        auto it = input_ids.find(key);
        if (it != input_ids.end()) return it->second;
        uint32_t id = (uint32_t)(p.inputs.size() / 2);
        p.inputs.push_back(col);
        p.inputs.push_back(rot);
        input_ids[key] = id;
        return id;
    }
    int leaf_node(uint32_t enc) {
        auto it = leaves.find(enc);
        if (it != leaves.end()) return it->second;
        Node n; n.op = -1; n.leaf = enc; n.a = n.b = -1;
        nodes.push_back(n);
        return leaves[enc] = (int)nodes.size() - 1;
    }
    int build(const ExprP &e) {
        switch (e->kind) {
            case Expr::CONST: return leaf_node(operand(K_CONST, const_id(e->c)));
            case Expr::COL: return leaf_node(operand(K_INPUT, input_id(e->col, e->rot)));
            default: break;
        }
        int a = build(e->a);
        int b = e->b ? build(e->b) : -1;
        int op = e->kind == Expr::ADD ? OP_ADD : e->kind == Expr::SUB ? OP_SUB : e->kind == Expr::MUL ? OP_MUL : OP_NEG;
        if ((op == OP_ADD || op == OP_MUL) && b < a) std::swap(a, b);  // commutative: canonical order for CSE
        const int nmul = (op == OP_MUL ? 1 : 0) + nodes[a].nmul + (b >= 0 ? nodes[b].nmul : 0);
        auto key = std::make_tuple(op, a, b, -1);
        auto it = interior.find(key);
        if (it != interior.end()) return it->second;
        Node n; n.op = op; n.leaf = 0; n.a = a; n.b = b; n.nmul = nmul > 1000000 ? 1000000 : nmul;
        nodes.push_back(n);
        return interior[key] = (int)nodes.size() - 1;
    }
};

}  // namespace

Program compile_terms(const std::vector<ExprP> &terms, const fr_t *fold) {
    Program p;
    Builder bld(p);
    const uint32_t ACC = 0;  // slot 0 holds the running Horner accumulator
    std::vector<bool> slot_busy(1, true);
    uint32_t fold_const = fold ? bld.const_id(*fold) : 0;
    auto emit = [&](int op, uint32_t dst, uint32_t a, uint32_t b) {
        p.code.push_back(((uint32_t)op << 16) | dst);
        p.code.push_back(a);
        p.code.push_back(b);
        if (op == OP_MUL) p.n_mul++;
        else if (op != OP_COPY) p.n_addsub++;
    };
    // ONE hash-consed DAG for all terms: a sub-expression shared by several gates (an S-box output feeding both MDS rows, the
    // round polynomial that the two Poseidon configurations gate with different selectors) is computed once.  The fold
    // sum_i y^(T-1-i) term_i is linear, so the terms may be ACCUMULATED IN ANY ORDER with explicit powers of y as constants
    // (one product + one addition per term, like Horner): terms that share the most work are scheduled next to each other so
    // shared values die quickly and the program needs few value slots (slots are shared memory: they bound occupancy).
    const size_t T = terms.size();
    std::vector<int> roots(T);
    for (size_t ti = 0; ti < T; ti++) {
        bld.cur_term = (int)ti;
        roots[ti] = bld.build(terms[ti]);
    }
    std::vector<Node> &nd = bld.nodes;
    for (auto &n : nd) {
        if (n.op < 0) continue;
        nd[n.a].uses++;
        if (n.b >= 0) nd[n.b].uses++;
    }
    for (int r : roots) nd[r].uses++;
    // interior nodes reachable from every root (sorted id lists) and their cost (a product ~ 8 additions)
    std::vector<std::vector<int>> reach(T);
    for (size_t ti = 0; ti < T; ti++) {
        std::vector<char> seen(nd.size(), 0);
        std::vector<int> stack{roots[ti]};
        while (!stack.empty()) {
            int id = stack.back();
            stack.pop_back();
            if (seen[id] || nd[id].op < 0) continue;
            seen[id] = 1;
            reach[ti].push_back(id);
            stack.push_back(nd[id].a);
            if (nd[id].b >= 0) stack.push_back(nd[id].b);
        }
        std::sort(reach[ti].begin(), reach[ti].end());
    }
    std::vector<char> emitted(nd.size(), 0);
    auto alloc = [&]() -> uint32_t {
        for (uint32_t s = 1; s < slot_busy.size(); s++)
            if (!slot_busy[s]) { slot_busy[s] = true; return s; }
        slot_busy.push_back(true);
        return (uint32_t)slot_busy.size() - 1;
    };
    auto opnd = [&](int id) -> uint32_t { return nd[id].op < 0 ? nd[id].leaf : operand(K_REG, (uint32_t)nd[id].slot); };
    auto release = [&](int id) {
        if (nd[id].op < 0) return;
        if (--nd[id].uses == 0) slot_busy[nd[id].slot] = false;
    };
    // powers of the fold challenge: term i carries y^(T-1-i)
    std::vector<uint32_t> pow_const(T, 0);
    if (fold) {
        hfr::Fr yp = hfr::ONE;
        const hfr::Fr y = to_host(*fold);
        for (size_t k = 0; k < T; k++) {
            pow_const[T - 1 - k] = bld.const_id(to_dev(yp));
            yp = hfr::mul(yp, y);
        }
    }
    (void)fold_const;

    // ---- scheduling units.  Terms of the shape f * body_i that share the factor f (the gates behind one selector, or behind one
    //      selector polynomial q(1-q)(2-q)...) are accumulated as f * sum_i y^(e_i) body_i: |G| + 1 products instead of 2 |G|.
    struct Unit { int factor; std::vector<size_t> terms; };
    std::vector<Unit> units;
    {
        std::map<int, int> root_count, child_freq;
        for (size_t ti = 0; ti < T; ti++) root_count[roots[ti]]++;
        auto groupable = [&](size_t ti) {
            const Node &r = nd[roots[ti]];
            return fold && r.op == OP_MUL && r.uses == 1 && root_count[roots[ti]] == 1 && r.a != r.b;
        };
        for (size_t ti = 0; ti < T; ti++)
            if (groupable(ti)) { child_freq[nd[roots[ti]].a]++; child_freq[nd[roots[ti]].b]++; }
        std::map<int, size_t> unit_of_factor;
        for (size_t ti = 0; ti < T; ti++) {
            int f = -1;
            if (groupable(ti)) {
                const int a2 = nd[roots[ti]].a, b2 = nd[roots[ti]].b;
                const int fa = child_freq[a2], fb = child_freq[b2];
                if (std::max(fa, fb) >= 2) f = fa >= fb ? a2 : b2;
            }
            if (f < 0) { units.push_back({-1, {ti}}); continue; }
            auto it = unit_of_factor.find(f);
            if (it == unit_of_factor.end()) { unit_of_factor[f] = units.size(); units.push_back({f, {ti}}); }
            else units[it->second].terms.push_back(ti);
        }
        for (Unit &u : units) {
            if (u.factor >= 0 && u.terms.size() == 1) u.factor = -1;  // the other sharers chose a different factor
            if (u.factor >= 0 && nd[u.factor].op >= 0) nd[u.factor].uses -= (int)u.terms.size() - 1;  // one use by the grouped product
        }
    }
    std::vector<char> unit_used(units.size(), 0);
    // next unit: the one that consumes the most value already sitting in slots (a product ~ 8 additions), i.e. lets shared values
    // die soonest; nothing live to consume -> the first unscheduled unit in source order.  Plain sums keep source order.
    auto pick_next = [&]() -> size_t {
        size_t best = units.size();
        int best_w = 0;
        for (size_t c = 0; c < units.size(); c++) {
            if (unit_used[c]) continue;
            if (!fold) return c;
            int w = 0;
            for (size_t ti : units[c].terms)
                for (int id : reach[ti])
                    if (emitted[id] && nd[id].uses > 0) w += nd[id].op == OP_MUL ? 8 : 1;
            if (w > best_w) { best_w = w; best = c; }
        }
        if (best == units.size())
            for (size_t c = 0; c < units.size() && best == units.size(); c++)
                if (!unit_used[c]) best = c;
        return best;
    };

    // Product-free sub-expressions (negated cells, selector complements, sums of cells) are identified globally like everything
    // else, but their VALUES are not kept across evaluations: an evaluation that needs one recomputes it (one addition is cheaper than
    // a value slot held for a long time; slots are shared memory and bound occupancy).  term_uses counts, per evaluation, the edges
    // from the nodes emitted in it into each cheap node (+ 1 for the caller's use of a cheap root).
    std::vector<int> cheap_stamp(nd.size(), -1), touch_stamp(nd.size(), -1), term_uses(nd.size(), 0);
    int stamp = 0;
    auto is_cheap = [&](int id) { return nd[id].op >= 0 && nd[id].nmul == 0; };
    auto release2 = [&](int id) {
        if (nd[id].op < 0) return;
        if (is_cheap(id)) {
            if (--term_uses[id] == 0) slot_busy[nd[id].slot] = false;
        } else {
            release(id);
        }
    };
    // make opnd(root) valid; the caller owes exactly one release2(root)
    auto eval = [&](int root) {
        stamp++;
        {   // dry walk: uses of cheap nodes inside this evaluation
            std::vector<int> stack{root};
            auto touch = [&](int id) {
                if (touch_stamp[id] != stamp) { touch_stamp[id] = stamp; term_uses[id] = 0; }
                term_uses[id]++;
            };
            if (is_cheap(root)) touch(root);  // the caller's use
            std::vector<int> visited_exp;
            while (!stack.empty()) {
                const int id = stack.back();
                stack.pop_back();
                Node &n = nd[id];
                if (n.op < 0) continue;
                if (is_cheap(id)) {
                    if (cheap_stamp[id] == stamp) continue;  // already expanded in this evaluation
                    cheap_stamp[id] = stamp;
                } else {
                    if (emitted[id]) continue;
                    if (n.slot == -2) continue;  // already expanded in this dry walk
                    n.slot = -2;
                    visited_exp.push_back(id);
                }
                for (int ch : {n.a, n.b}) {
                    if (ch < 0) continue;
                    if (is_cheap(ch)) touch(ch);
                    stack.push_back(ch);
                }
            }
            for (int id : visited_exp) nd[id].slot = -1;
            for (size_t q = 0; q < nd.size(); q++)
                if (cheap_stamp[q] == stamp) cheap_stamp[q] = -1 - stamp;  // "to be emitted in this evaluation", not yet emitted
        }
        // post-order walk from the root, skipping what is already available
        std::vector<std::pair<int, int>> stack{{root, 0}};
        while (!stack.empty()) {
            auto &top = stack.back();
            const int id = top.first;
            Node &n = nd[id];
            if (n.op < 0 || (is_cheap(id) ? cheap_stamp[id] == stamp : (bool)emitted[id])) { stack.pop_back(); continue; }
            if (top.second == 0) { top.second = 1; stack.push_back({n.a, 0}); continue; }
            if (top.second == 1) { top.second = 2; if (n.b >= 0) { stack.push_back({n.b, 0}); continue; } }
            const uint32_t a = opnd(n.a), b = n.b >= 0 ? opnd(n.b) : 0;
            // operands may be released before the destination is chosen: dst may alias a dying source
            release2(n.a);
            if (n.b >= 0) release2(n.b);
            n.slot = (int)alloc();
            emit(n.op, (uint32_t)n.slot, a, b);
            if (is_cheap(id)) cheap_stamp[id] = stamp;
            else emitted[id] = 1;
            stack.pop_back();
        }
    };
    bool acc_live = false;
    // acc (+)= value * y^(T-1-ti); value is an operand the caller has already released (it may die into the destination)
    auto accumulate = [&](uint32_t v_op, bool scaled, uint32_t pconst) {
        if (!acc_live) {
            if (scaled) emit(OP_MUL, ACC, v_op, operand(K_CONST, pconst));
            else emit(OP_COPY, ACC, v_op, 0);
            acc_live = true;
        } else if (scaled) {
            const uint32_t tmp = alloc();
            emit(OP_MUL, tmp, v_op, operand(K_CONST, pconst));
            emit(OP_ADD, ACC, operand(K_REG, ACC), operand(K_REG, tmp));
            slot_busy[tmp] = false;
        } else {
            emit(OP_ADD, ACC, operand(K_REG, ACC), v_op);
        }
    };
    for (size_t oi = 0; oi < units.size(); oi++) {
        const size_t ui = pick_next();
        unit_used[ui] = 1;
        const Unit &u = units[ui];
        if (u.factor < 0) {
            const size_t ti = u.terms[0];
            const int root = roots[ti];
            eval(root);
            const uint32_t t_op = opnd(root);
            release2(root);
            accumulate(t_op, fold && ti != T - 1, pow_const[ti]);  // the last term carries y^0
        } else {
            // gacc = sum_i y^(e_i) body_i, then acc += factor * gacc.  The roots factor * body_i themselves are never emitted.
            uint32_t gacc = 0;
            bool g_live = false;
            for (size_t ti : u.terms) {
                const Node &r = nd[roots[ti]];
                const int body = r.a == u.factor ? r.b : r.a;
                eval(body);
                const uint32_t b_op = opnd(body);
                release2(body);
                const bool scaled = ti != T - 1;
                if (!g_live) {
                    gacc = alloc();
                    if (scaled) emit(OP_MUL, gacc, b_op, operand(K_CONST, pow_const[ti]));
                    else emit(OP_COPY, gacc, b_op, 0);
                    g_live = true;
                } else if (scaled) {
                    const uint32_t tmp = alloc();
                    emit(OP_MUL, tmp, b_op, operand(K_CONST, pow_const[ti]));
                    emit(OP_ADD, gacc, operand(K_REG, gacc), operand(K_REG, tmp));
                    slot_busy[tmp] = false;
                } else {
                    emit(OP_ADD, gacc, operand(K_REG, gacc), b_op);
                }
            }
            eval(u.factor);
            const uint32_t f_op = opnd(u.factor);
            release2(u.factor);
            emit(OP_MUL, gacc, operand(K_REG, gacc), f_op);
            slot_busy[gacc] = false;
            accumulate(operand(K_REG, gacc), false, 0);
        }
    }
    p.n_slots = (uint32_t)slot_busy.size();
    p.out_slot = ACC;
    return p;
}

// Host interpreter of a compiled program for ONE row (CPU tests of the compiler; the device kernel below runs the same code).
// inputs[i] is the value of p.inputs pair i (column, rotation) at that row.
fr_t program_eval_host(const Program &p, const std::vector<fr_t> &inputs) {
    std::vector<hfr::Fr> slot(p.n_slots, hfr::ZERO);
    auto fetch = [&](uint32_t o) -> hfr::Fr {
        const uint32_t kind = o >> 30, idx = o & 0x3fffffffu;
        if (kind == K_REG) return slot[idx];
        if (kind == K_CONST) return to_host(p.consts[idx]);
        return to_host(inputs[idx]);
    };
    for (size_t pc = 0; pc < p.code.size() / 3; pc++) {
        const uint32_t w0 = p.code[3 * pc], wa = p.code[3 * pc + 1], wb = p.code[3 * pc + 2];
        const uint32_t op = w0 >> 16, dst = w0 & 0xffffu;
        const hfr::Fr a = fetch(wa);
        hfr::Fr r;
        if (op == OP_MUL) r = hfr::mul(a, fetch(wb));
        else if (op == OP_ADD) r = hfr::add(a, fetch(wb));
        else if (op == OP_SUB) r = hfr::sub(a, fetch(wb));
        else if (op == OP_NEG) r = hfr::neg(a);
        else r = a;
        slot[dst] = r;
    }
    return to_dev(slot[p.out_slot]);
}

// ------------------------------------------------------------------ device interpreter
struct ExprArgs {
    const uint32_t *code;
    uint32_t n_instr;
    const uint4 *consts;
    const int32_t *inputs;
    const uint4 *const *cols;
    const uint8_t *col_shift;  // per column: log2 of the element stride (0 = compact column, 3 = one coset of an extended column)
    uint32_t log_n, rot_scale_log;
    uint4 *out;
    uint32_t out_slot;  // (n_slots << 16) | output slot
};

static const int EXPR_THREADS = 128;

__device__ __forceinline__ fr_t expr_fetch(uint32_t opnd, const ExprArgs &A, const uint4 *s_lo, const uint4 *s_hi, uint32_t tid, uint64_t row) {
    const uint32_t kind = opnd >> 30, idx = opnd & 0x3fffffffu;
    fr_t r;
    if (kind == K_REG) {
        uint4 a = s_lo[idx * EXPR_THREADS + tid], b = s_hi[idx * EXPR_THREADS + tid];
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    } else if (kind == K_CONST) {
        r = ldg_fp<FrParams>(A.consts + 2 * idx);
    } else {
        const int32_t col = __ldg(A.inputs + 2 * idx), rot = __ldg(A.inputs + 2 * idx + 1);
        const uint64_t mask = (1ull << A.log_n) - 1;
        const uint64_t j = (row + (uint64_t)((int64_t)rot * (int64_t)(1ll << A.rot_scale_log))) & mask;
        const uint4 *base = A.cols[col];
        r = ldg_fp<FrParams>(base + 2 * (j << __ldg(A.col_shift + col)));
    }
    return r;
}

__global__ void __launch_bounds__(EXPR_THREADS) expr_eval_kernel(const ExprArgs A) {
    extern __shared__ uint4 smem[];
    const uint32_t tid = threadIdx.x;
    const uint64_t row = blockIdx.x * (uint64_t)EXPR_THREADS + tid;
    // value slots: [slot][thread]; low 128 bits of every slot first, then the high 128 bits
    const uint32_t n_slots = A.out_slot >> 16;
    uint4 *lo = smem, *hi = smem + (size_t)n_slots * EXPR_THREADS;
    const uint32_t out_slot = A.out_slot & 0xffffu;
    for (uint32_t pc = 0; pc < A.n_instr; pc++) {
        const uint32_t w0 = __ldg(A.code + 3 * pc), wa = __ldg(A.code + 3 * pc + 1), wb = __ldg(A.code + 3 * pc + 2);
        const uint32_t op = w0 >> 16, dst = w0 & 0xffffu;
        fr_t a = expr_fetch(wa, A, lo, hi, tid, row);
        fr_t r;
        if (op == OP_MUL) {
            r = mul(a, expr_fetch(wb, A, lo, hi, tid, row));
        } else if (op == OP_ADD) {
            r = add(a, expr_fetch(wb, A, lo, hi, tid, row));
        } else if (op == OP_SUB) {
            r = sub(a, expr_fetch(wb, A, lo, hi, tid, row));
        } else if (op == OP_NEG) {
            r = neg(a);
        } else {
            r = a;
        }
        lo[dst * EXPR_THREADS + tid] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        hi[dst * EXPR_THREADS + tid] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    }
    A.out[2 * row] = lo[out_slot * EXPR_THREADS + tid];
    A.out[2 * row + 1] = hi[out_slot * EXPR_THREADS + tid];
}

int32_t expr_eval(sb_ctx *ctx, const Program &prog, const std::vector<const void *> &cols, uint32_t log_n, uint32_t rot_scale_log, void *d_out, cudaStream_t st,
                  const std::vector<uint8_t> *col_shift) {
    const uint64_t n = 1ull << log_n;
    SB_REQUIRE(n >= (uint64_t)EXPR_THREADS, "expr_eval: domain smaller than one CTA");
    SB_REQUIRE(prog.n_slots >= 1 && prog.n_slots <= 48, "expr_eval: too many live values");
    const size_t code_b = prog.code.size() * 4, const_b = prog.consts.size() * 32, in_b = prog.inputs.size() * 4, col_b = cols.size() * sizeof(void *);
    const size_t off_const = (code_b + 31) & ~(size_t)31, off_in = off_const + ((const_b + 31) & ~(size_t)31), off_col = off_in + ((in_b + 31) & ~(size_t)31);
    const size_t off_shift = off_col + ((col_b + 31) & ~(size_t)31);
    const size_t total = off_shift + cols.size() + 32;
    SB_REQUIRE(!col_shift || col_shift->size() == cols.size(), "expr_eval: one stride per column");
    std::vector<uint8_t> host(total, 0);
    memcpy(host.data(), prog.code.data(), code_b);
    if (const_b) memcpy(host.data() + off_const, prog.consts.data(), const_b);
    if (in_b) memcpy(host.data() + off_in, prog.inputs.data(), in_b);
    if (col_b) memcpy(host.data() + off_col, cols.data(), col_b);
    if (col_shift && !cols.empty()) memcpy(host.data() + off_shift, col_shift->data(), cols.size());
    // each launch gets its own staging slice so that programs queued back-to-back on one stream do not clobber each other
    static thread_local uint32_t ring = 0;
    const uint32_t slice = ring++ % 8;
    uint8_t *d_prog;
    const size_t slice_bytes = 1 << 20;
    SB_REQUIRE(total <= slice_bytes, "expr_eval: program too large");
    SB_TRY(scratch_get(ctx, "expr_prog", 8 * slice_bytes, (void **)&d_prog));
    d_prog += (size_t)slice * slice_bytes;
    SB_TRY(h2d_staged(ctx, d_prog, host.data(), total, st));  // pinned ring: no stream synchronisation, `host` may die at return
    ExprArgs A;
    A.code = (const uint32_t *)d_prog;
    A.n_instr = (uint32_t)(prog.code.size() / 3);
    A.consts = (const uint4 *)(d_prog + off_const);
    A.inputs = (const int32_t *)(d_prog + off_in);
    A.cols = (const uint4 *const *)(d_prog + off_col);
    A.col_shift = (const uint8_t *)(d_prog + off_shift);
    A.log_n = log_n;
    A.rot_scale_log = rot_scale_log;
    A.out = (uint4 *)d_out;
    A.out_slot = (prog.n_slots << 16) | prog.out_slot;
    const size_t smem = (size_t)prog.n_slots * EXPR_THREADS * 32;
    // per device and idempotent: safe to repeat from concurrent contexts
    SB_CUDA_TRY(cudaFuncSetAttribute(expr_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * EXPR_THREADS * 32));
    SB_LAUNCH(ctx, expr_eval_kernel, (unsigned)(n / EXPR_THREADS), EXPR_THREADS, smem, st, A);
    return SB_OK;
}

}  // namespace sb
