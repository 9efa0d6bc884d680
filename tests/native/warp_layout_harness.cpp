// CPU test harness for the shared-memory strip layout of the warp-per-codeword kernel
// (csrc/code_tables.cpp: color_warp_layout) and the coloured rows of the n = 8192 register-table
// kernel (color_regular_rows).  Checks what the kernels rely on:
//   * every check-phase access (slot s, 32 lanes) and every variable-phase access (half t, edge k,
//     32 lanes) touches 32 DIFFERENT banks -- real edges by the colouring, unused slots through the
//     free bank they were given (identity / zero row for reads, own word / dummy row for writes);
//   * positions of real edges are a bijection edge <-> word, and both tables agree on it.
// Codes: the shipped 32x64 code and random sparse codes with uneven degrees (argv: seed count).
#include "code_tables.h"
#include "ldpc535_default_code.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <vector>

using namespace ldpc535;

static int check_warp_layout(const CodeTables &t, const char *name)
{
    const int M = t.M, N = t.N, DC = t.dc_max, DV = t.dv_max;
    std::vector<int> seen((size_t)DC * 32, 0);
    for (int s = 0; s < DC; s++) {                       // check phase: lane j reads / writes slot s of check j
        std::set<int> banks;
        for (int j = 0; j < 32; j++) {
            const int pos = t.w_chk_pos[(size_t)s * 32 + j];
            if (pos / 32 != s) { std::printf("FAIL %s: slot %d lane %d sits in row %d\n", name, s, j, pos / 32); return 1; }
            banks.insert(pos % 32);
            const bool real = j < M && t.row_ptr[j] + s < t.row_ptr[j + 1];
            if (real) {
                if (t.w_pos_edge[pos] != t.row_ptr[j] + s) { std::printf("FAIL %s: pos_edge\n", name); return 1; }
                seen[pos]++;
            }
        }
        if ((int)banks.size() != 32) { std::printf("FAIL %s: check access of slot %d hits %zu banks\n", name, s, banks.size()); return 1; }
    }
    for (int k = 0; k < DV; k++)
        for (int half = 0; half < 2; half++) {           // variable phase: lane l handles edge k of bit l + 32 half
            std::set<int> banks;
            for (int l = 0; l < 32; l++) {
                const int c = l + 32 * half;
                const uint16_t pos = t.w_var_pos[(size_t)k * 64 + c];
                const bool real = c < N && t.col_ptr[c] + k < t.col_ptr[c + 1];
                if (real) {
                    if (pos & 0x8000) { std::printf("FAIL %s: real edge flagged unused\n", name); return 1; }
                    const int e = t.edge_of_col[t.col_ptr[c] + k];
                    if (t.w_pos_edge[pos] != e) { std::printf("FAIL %s: the two tables disagree on edge %d\n", name, e); return 1; }
                    banks.insert(pos % 32);
                } else {
                    if (!(pos & 0x8000)) { std::printf("FAIL %s: unused slot (k=%d, bit %d) has no free bank\n", name, k, c); return 1; }
                    banks.insert(pos & 31);
                }
            }
            if ((int)banks.size() != 32) { std::printf("FAIL %s: variable access (half %d, edge %d) hits %zu banks\n", name, half, k, banks.size()); return 1; }
        }
    for (int e = 0; e < t.E; e++) {
        int hits = 0;
        for (size_t p = 0; p < t.w_pos_edge.size(); p++) hits += t.w_pos_edge[p] == e;
        if (hits != 1) { std::printf("FAIL %s: edge %d has %d words\n", name, e, hits); return 1; }
    }
    return 0;
}

static int check_regular_rows(const CodeTables &t, const char *name)
{
    if (t.row_color.empty()) { std::printf("FAIL %s: coloured rows not built\n", name); return 1; }
    const int M = t.M, N = t.N, DC = t.dc_max, DV = t.dv_max;
    for (int row = 0; row < M / 32; row++)               // check phase: a row of 32 storage columns, one slot
        for (int s = 0; s < DC; s++) {
            std::set<int> banks;
            for (int l = 0; l < 32; l++) banks.insert(t.row_color[(size_t)(row * 32 + l) * DC + s]);
            if ((int)banks.size() != 32) { std::printf("FAIL %s: row %d slot %d hits %zu banks\n", name, row, s, banks.size()); return 1; }
        }
    std::set<int> words;
    for (int k = 0; k < DV; k++)
        for (int w = 0; w < N / 32; w++) {               // variable phase: 32 consecutive bits, edge k
            std::set<int> banks;
            for (int l = 0; l < 32; l++) {
                const int pos = t.var_slot_colored[(size_t)k * N + w * 32 + l];
                banks.insert(pos % 32);
                words.insert(pos);
            }
            if ((int)banks.size() != 32) { std::printf("FAIL %s: bits %d.., edge %d hit %zu banks\n", name, w * 32, k, banks.size()); return 1; }
        }
    if ((int)words.size() != t.E) { std::printf("FAIL %s: %zu words for %d edges\n", name, words.size(), t.E); return 1; }
    return 0;
}

// random M x N parity-check matrix with an identity-like left part (so that it has an LU pivot order),
// check degrees 2..dc, bit degrees uneven
static bool random_code(std::mt19937 &rng, int M, int N, int dc, std::vector<int32_t> &row_ptr, std::vector<int32_t> &col_idx)
{
    row_ptr.assign(1, 0);
    col_idx.clear();
    for (int j = 0; j < M; j++) {
        std::set<int> cols;
        cols.insert(j);                                   // pivot
        const int d = 2 + (int)(rng() % (unsigned)(dc - 1));
        while ((int)cols.size() < d) cols.insert(M + (int)(rng() % (unsigned)(N - M)));
        for (int c : cols) col_idx.push_back(c);
        row_ptr.push_back((int32_t)col_idx.size());
    }
    return true;
}

int main(int argc, char **argv)
{
    const int n_random = argc > 1 ? std::atoi(argv[1]) : 20;
    CodeTables t;
    if (build_code_tables(ldpc535_default_row_ptr, ldpc535_default_col_idx, 32, 64, t)) { std::printf("FAIL: shipped code\n"); return 1; }
    if (check_warp_layout(t, "shipped")) return 1;
    std::mt19937 rng(535);
    int built = 0;
    for (int i = 0; i < n_random; i++) {
        const int M = 4 + (int)(rng() % 29), N = M + 1 + (int)(rng() % (unsigned)(64 - M));
        std::vector<int32_t> rp, ci;
        random_code(rng, M, N, std::min(6, N - M + 1), rp, ci);
        CodeTables r;
        if (build_code_tables(rp.data(), ci.data(), M, N, r)) continue;     // singular: not a code the library accepts
        if (r.dv_max > 8 || r.dc_max > 16) continue;
        built++;
        char name[64];
        std::snprintf(name, sizeof name, "random %d (%dx%d)", i, M, N);
        if (check_warp_layout(r, name)) return 1;
    }
    // small (3,6)-regular codes with the shape of BASELINE config 4's (M a multiple of 32): permutation
    // construction, retried until one is simple and has an LU pivot order
    int regular_built = 0;
    for (int attempt = 0; attempt < 200 && regular_built < 2; attempt++) {
        const int M = 256, N = 512;
        std::vector<std::vector<int>> rows(M);
        for (int rep = 0; rep < 3; rep++) {
            for (int retry = 0; retry < 2000; retry++) {     // a permutation that gives no row a column it already has
                std::vector<int> perm(N);
                for (int c = 0; c < N; c++) perm[c] = c;
                std::shuffle(perm.begin(), perm.end(), rng);
                bool clash = false;
                for (int c = 0; c < N && !clash; c++) {
                    const std::vector<int> &r = rows[perm[c] % M];
                    clash = std::find(r.begin(), r.end(), c) != r.end();
                }
                if (clash) continue;
                for (int c = 0; c < N; c++) rows[perm[c] % M].push_back(c);
                break;
            }
        }
        std::vector<int32_t> rp(1, 0), ci;
        bool regular = true;
        for (int j = 0; j < M; j++) {
            std::sort(rows[j].begin(), rows[j].end());
            if (std::adjacent_find(rows[j].begin(), rows[j].end()) != rows[j].end() || rows[j].size() != 6) regular = false;
            for (int c : rows[j]) ci.push_back(c);
            rp.push_back((int32_t)ci.size());
        }
        CodeTables r;
        if (!regular || build_code_tables(rp.data(), ci.data(), M, N, r) != 0) continue;
        if (check_regular_rows(r, "regular 256x512")) return 1;
        regular_built++;
    }
    if (regular_built == 0) { std::printf("FAIL: no regular test code could be built\n"); return 1; }
    built += regular_built;
    std::printf("ok %d\n", built);
    return 0;
}
