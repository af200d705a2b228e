// Witness generation for `MstInclusionCircuit<LEVELS, N_CURRENCIES, N_BYTES>`: the advice cells `Circuit::synthesize` assigns, straight from a
// Merkle proof of the device-resident Merkle sum tree (sb_mst_proofs), so that the batched inclusion-proof service (BASELINE configs[4]:
// backend/src/apis/round.rs:153-174 -- `Tree::generate_proof` -> `MstInclusionCircuit::init` -> `gen_proof_solidity_calldata`) runs from tree to
// proof without the Rust front-end.
//
// Restates the WITNESS side of zk_prover/src/circuits/merkle_sum_tree.rs:228-520 (synthesize), chips/merkle_sum_tree.rs:107-227 (swap / sum regions),
// chips/range/range_check.rs:93-153 (running-sum byte decomposition), chips/poseidon/hash.rs:75-87 over halo2_gadgets' Pow5Chip region layout
// (initial state | per absorbed word: add input (3 rows) + permute state (37 rows: 4 full, 28 double-partial, 4 full rounds)), placed the way
// halo2's SimpleFloorPlanner places regions: a region starts at the highest next-free row of the columns it touches (selectors and fixed columns
// count), and a region's constants are appended to the first constants column (fixed 2) after it.  The key (fixed columns, copy constraints,
// selectors) does not depend on the witness and is NOT produced here.  Host C++ (the reference does this on the CPU as well, once per proof; it is
// ~1e5 field products); parity: tests/test_host_logic.py compares every cell with the oracle's synthesis for several (LEVELS, N_CURRENCIES).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <vector>

#include "../../include/summa_b200.h"
#include "hostfr.h"
#include "poseidon_constants.inc"

namespace sb {
void set_last_error(const char *fmt, ...);
}

namespace {
using sb::hfr::Fr;
namespace hfr = sb::hfr;

// column ids of the floor planner: advice 0..2, fixed 0..4 (5..9), selectors (10..18)
enum { A0 = 0, A1 = 1, A2 = 2, F0 = 5, F1 = 6, F2 = 7, F3 = 8, F4 = 9, SEL0 = 10 };
enum { S_BOOL_SWAP = 0, S_SUM, S_LOOKUP, S_FULL_E, S_PARTIAL_E, S_PAD_E, S_FULL_M, S_PARTIAL_M, S_PAD_M };
const int N_COLS = 19;

Fr load_const(const uint32_t w[8]) {
    Fr r;
    memcpy(r.v, w, 32);
    return r;
}

struct Cell {
    int col;
    uint32_t row;
    Fr value;
};

// SingleChipLayouter: every region is described by the columns it touches and its height; cells are written relative to its start
struct Planner {
    uint32_t next_free[N_COLS];
    uint32_t n_rows;
    std::vector<uint32_t> cells;   // (col, row) pairs of the assigned advice cells
    std::vector<Fr> values;
    bool overflow = false;
    explicit Planner(uint32_t n) : n_rows(n) { memset(next_free, 0, sizeof next_free); }

    uint32_t place(std::initializer_list<int> cols, uint32_t rows, uint32_t n_constants) {
        uint32_t start = 0;
        for (int c : cols) start = std::max(start, next_free[c]);
        for (int c : cols) next_free[c] = start + rows;
        next_free[F2] += n_constants;  // the region's constants go to the first constants column, one row each, after the region is placed
        if (start + rows > n_rows) overflow = true;
        return start;
    }
    Cell put(int col, uint32_t row, const Fr &v) {
        if (!hfr::is_zero(v)) {  // sparse output: zero cells are the default
            cells.push_back((uint32_t)col);
            cells.push_back(row);
            values.push_back(v);
        }
        return Cell{col, row, v};
    }
};

struct Poseidon {
    Fr rc[64][2], mds[2][2];
    Poseidon() {
        for (int r = 0; r < 64; r++)
            for (int i = 0; i < 2; i++) rc[r][i] = load_const(POSEIDON_RC_HOST[r][i]);
        for (int i = 0; i < 2; i++)
            for (int j = 0; j < 2; j++) mds[i][j] = load_const(POSEIDON_MDS_HOST[i][j]);
    }
    static Fr pow5(const Fr &x) {
        const Fr x2 = hfr::sqr(x);
        return hfr::mul(hfr::sqr(x2), x);
    }
    void mix(const Fr in[2], Fr out[2]) const {
        out[0] = hfr::add(hfr::mul(mds[0][0], in[0]), hfr::mul(mds[0][1], in[1]));
        out[1] = hfr::add(hfr::mul(mds[1][0], in[0]), hfr::mul(mds[1][1], in[1]));
    }
};

// halo2_gadgets Hash<_, _, S, ConstantLength<L>, 2, 1>::init(..).hash(..) with Pow5Chip: returns the cell holding the digest
Cell poseidon_hash_chip(Planner &pl, const Poseidon &P, const std::vector<Cell> &inputs, int s_full, int s_partial, int s_pad) {
    const uint64_t L = inputs.size();
    // "initial state": [0, L * 2^64] assigned from constants
    uint32_t st0 = pl.place({A0, A1}, 1, 2);
    uint64_t cap[4] = {0, L, 0, 0};
    Cell state[2] = {pl.put(A0, st0, hfr::ZERO), pl.put(A1, st0, hfr::from_canonical(cap))};
    for (const Cell &word : inputs) {
        // "add input": rows 0 (state copy), 1 (input in column 0; the rate is 1), 2 (state + input)
        uint32_t r0 = pl.place({SEL0 + s_pad, A0, A1}, 3, 0);
        pl.put(A0, r0, state[0].value);
        pl.put(A1, r0, state[1].value);
        pl.put(A0, r0 + 1, word.value);
        state[0] = pl.put(A0, r0 + 2, hfr::add(state[0].value, word.value));
        state[1] = pl.put(A1, r0 + 2, state[1].value);
        // "permute state": 4 full rounds, 28 rows of two partial rounds each, 4 full rounds; row i + 1 holds the state after row i's round(s)
        uint32_t p0 = pl.place({A0, A1, SEL0 + s_full, F0, F1, SEL0 + s_partial, A2, F2, F3}, 37, 0);
        Fr cur[2] = {state[0].value, state[1].value};
        pl.put(A0, p0, cur[0]);
        pl.put(A1, p0, cur[1]);
        auto full = [&](int rnd, uint32_t off) {
            Fr t[2] = {Poseidon::pow5(hfr::add(cur[0], P.rc[rnd][0])), Poseidon::pow5(hfr::add(cur[1], P.rc[rnd][1]))};
            P.mix(t, cur);
            pl.put(A0, p0 + off + 1, cur[0]);
            pl.put(A1, p0 + off + 1, cur[1]);
        };
        auto partial = [&](int rnd, uint32_t off) {
            Fr t[2] = {Poseidon::pow5(hfr::add(cur[0], P.rc[rnd][0])), hfr::add(cur[1], P.rc[rnd][1])};
            pl.put(A2, p0 + off, t[0]);  // partial_sbox
            Fr mid[2];
            P.mix(t, mid);
            Fr u[2] = {Poseidon::pow5(hfr::add(mid[0], P.rc[rnd + 1][0])), hfr::add(mid[1], P.rc[rnd + 1][1])};
            P.mix(u, cur);
            pl.put(A0, p0 + off + 1, cur[0]);
            pl.put(A1, p0 + off + 1, cur[1]);
        };
        for (int i = 0; i < 4; i++) full(i, (uint32_t)i);
        for (int i = 0; i < 28; i++) partial(4 + 2 * i, (uint32_t)(4 + i));
        for (int i = 0; i < 4; i++) full(60 + i, (uint32_t)(32 + i));
        state[0] = Cell{A0, p0 + 36, cur[0]};
        state[1] = Cell{A1, p0 + 36, cur[1]};
    }
    return state[0];
}

Cell assign_value(Planner &pl, const Fr &v, int col) {
    uint32_t r = pl.place({col}, 1, 0);
    return pl.put(col, r, v);
}

// RangeCheckU64Chip-style running sum: z_0 = value, z_{i+1} = (z_i - byte_i) / 256 for the N_BYTES low bytes; z_N is constrained to 0
bool range_check(Planner &pl, const Cell &value, uint32_t n_bytes, const Fr &inv256) {
    uint32_t r = pl.place({SEL0 + S_LOOKUP, A0}, n_bytes + 1, 1);
    uint64_t canon[4];
    hfr::to_canonical(value.value, canon);
    const uint8_t *bytes = (const uint8_t *)canon;
    Fr z = value.value;
    pl.put(A0, r, z);
    for (uint32_t i = 0; i < n_bytes; i++) {
        z = hfr::mul(hfr::sub(z, hfr::from_u64(bytes[i])), inv256);
        pl.put(A0, r + i + 1, z);
    }
    return hfr::is_zero(z);  // a balance outside N_BYTES does not decompose: the prover would fail its lookup / constant constraint
}

}  // namespace

extern "C" int32_t sb_mst_inclusion_witness(uint32_t levels, uint32_t n_currencies, uint32_t n_bytes, uint32_t k, const uint8_t *preimages, const uint8_t *path_indices,
                                            uint32_t *out_cells, uint8_t *out_values, size_t cap_cells, size_t *out_n_cells, uint8_t *out_instances) {
    if (!preimages || (levels && !path_indices) || !out_n_cells || !out_instances || (cap_cells && (!out_cells || !out_values))) return SB_ERR_ARG;
    if (levels < 1 || levels > 30 || n_currencies < 1 || n_currencies > 32 || n_bytes < 1 || n_bytes > 31 || k < 7 || k > 28) {
        sb::set_last_error("sb_mst_inclusion_witness: LEVELS in [1, 30], N_CURRENCIES in [1, 32], N_BYTES in [1, 31], k in [7, 28]");
        return SB_ERR_ARG;
    }
    static const Poseidon P;
    static const Fr inv256 = hfr::inv(hfr::from_u64(256));
    const uint32_t nc = n_currencies;
    Planner pl((1u << k) - 6);  // the last blinding_factors + 1 = 6 rows are not usable
    auto fr_at = [&](size_t i) { Fr v; memcpy(v.v, preimages + i * 32, 32); return v; };
    // entry preimage (nc + 1) | sibling leaf preimage (nc + 1) | (levels - 1) x sibling middle-node preimage (nc + 2): the layout sb_mst_proofs writes
    size_t off = 0;
    Cell username = assign_value(pl, fr_at(off), A0);
    std::vector<Cell> balances;
    for (uint32_t c = 0; c < nc; c++) balances.push_back(assign_value(pl, fr_at(off + 1 + c), A1));
    off += nc + 1;
    std::vector<Cell> in;
    in.push_back(username);
    in.insert(in.end(), balances.begin(), balances.end());
    Cell current = poseidon_hash_chip(pl, P, in, S_FULL_E, S_PARTIAL_E, S_PAD_E);
    const Fr leaf_hash = current.value;
    pl.place({F4}, 256, 0);  // the byte table
    bool in_range = true;
    for (uint32_t level = 0; level < levels; level++) {
        std::vector<Cell> sib_bal;
        Cell sibling;
        if (level == 0) {
            Cell su = assign_value(pl, fr_at(off), A0);
            for (uint32_t c = 0; c < nc; c++) sib_bal.push_back(assign_value(pl, fr_at(off + 1 + c), A1));
            off += nc + 1;
            in.clear();
            in.push_back(su);
            in.insert(in.end(), sib_bal.begin(), sib_bal.end());
            sibling = poseidon_hash_chip(pl, P, in, S_FULL_E, S_PARTIAL_E, S_PAD_E);
            for (uint32_t c = 0; c < nc; c++) {
                in_range = range_check(pl, balances[c], n_bytes, inv256) && in_range;
                in_range = range_check(pl, sib_bal[c], n_bytes, inv256) && in_range;
            }
        } else {
            for (uint32_t c = 0; c < nc; c++) sib_bal.push_back(assign_value(pl, fr_at(off + c), A1));
            Cell lh = assign_value(pl, fr_at(off + nc), A2);
            Cell rh = assign_value(pl, fr_at(off + nc + 1), A2);
            off += nc + 2;
            in = sib_bal;
            in.push_back(lh);
            in.push_back(rh);
            sibling = poseidon_hash_chip(pl, P, in, S_FULL_M, S_PARTIAL_M, S_PAD_M);
            for (uint32_t c = 0; c < nc; c++) in_range = range_check(pl, sib_bal[c], n_bytes, inv256) && in_range;
        }
        const bool right = path_indices[level] != 0;
        Cell bit = assign_value(pl, right ? hfr::ONE : hfr::ZERO, A0);
        // swap region: row 0 = (current, sibling, bit), row 1 = (left, right)
        uint32_t sr = pl.place({SEL0 + S_BOOL_SWAP, A0, A1, A2}, 2, 0);
        pl.put(A0, sr, current.value);
        pl.put(A1, sr, sibling.value);
        pl.put(A2, sr, bit.value);
        Cell left = pl.put(A0, sr + 1, right ? sibling.value : current.value);
        Cell rght = pl.put(A1, sr + 1, right ? current.value : sibling.value);
        std::vector<Cell> next;
        for (uint32_t c = 0; c < nc; c++) {
            uint32_t r = pl.place({SEL0 + S_SUM, A0, A1, A2}, 1, 0);
            pl.put(A0, r, balances[c].value);
            pl.put(A1, r, sib_bal[c].value);
            next.push_back(pl.put(A2, r, hfr::add(balances[c].value, sib_bal[c].value)));
        }
        in = next;
        in.push_back(left);
        in.push_back(rght);
        current = poseidon_hash_chip(pl, P, in, S_FULL_M, S_PARTIAL_M, S_PAD_M);
        balances = next;
    }
    if (pl.overflow) {
        sb::set_last_error("sb_mst_inclusion_witness: the circuit does not fit 2^%u rows (NotEnoughRowsAvailable)", k);
        return SB_ERR_ARG;
    }
    if (!in_range) {
        sb::set_last_error("sb_mst_inclusion_witness: a balance does not fit N_BYTES = %u bytes (the range check cannot be satisfied)", n_bytes);
        return SB_ERR_ARG;
    }
    // instances: leaf hash, root hash, root balances (merkle_sum_tree.rs:54-58)
    memcpy(out_instances, leaf_hash.v, 32);
    memcpy(out_instances + 32, current.value.v, 32);
    for (uint32_t c = 0; c < nc; c++) memcpy(out_instances + 64 + 32 * c, balances[c].value.v, 32);
    *out_n_cells = pl.values.size();
    if (pl.values.size() > cap_cells) {
        if (cap_cells) sb::set_last_error("sb_mst_inclusion_witness: %zu cells, buffer holds %zu", pl.values.size(), cap_cells);
        return cap_cells ? SB_ERR_ARG : SB_OK;  // cap 0 = size query
    }
    memcpy(out_cells, pl.cells.data(), pl.cells.size() * 4);
    memcpy(out_values, pl.values.data(), pl.values.size() * 32);
    return SB_OK;
}
