// See code_tables.h.  Everything here is GF(2) on bit-packed rows: the reference does
// the same steps on dense uBLAS int matrices (O(M^2 N) element operations and, in the
// encoder, two dense real-valued LAPACK solves PER FRAME); here the factorisation runs
// once per code and its effect is folded into one generator matrix P.
#include "code_tables.h"

#include <algorithm>
#include <cstring>

namespace ldpc535 {
namespace {

struct BitMat {
    int rows = 0, cols = 0, words = 0;
    std::vector<uint64_t> w;
    BitMat() {}
    BitMat(int r, int c) : rows(r), cols(c), words((c + 63) / 64), w((size_t)r * ((c + 63) / 64), 0) {}
    uint64_t *row(int r) { return w.data() + (size_t)r * words; }
    const uint64_t *row(int r) const { return w.data() + (size_t)r * words; }
    bool get(int r, int c) const { return (row(r)[c >> 6] >> (c & 63)) & 1u; }
    void set(int r, int c) { row(r)[c >> 6] |= (uint64_t)1 << (c & 63); }
    void flip(int r, int c) { row(r)[c >> 6] ^= (uint64_t)1 << (c & 63); }
};

// first set bit of a row at column >= from, or -1
int first_set_from(const uint64_t *r, int words, int from, int cols)
{
    int wi = from >> 6;
    if (wi >= words) return -1;
    uint64_t cur = r[wi] & (~(uint64_t)0 << (from & 63));
    while (true) {
        if (cur) {
            int c = (wi << 6) + __builtin_ctzll(cur);
            return c < cols ? c : -1;
        }
        if (++wi >= words) return -1;
        cur = r[wi];
    }
}

// call fn(c) for every set bit c of the row with lo <= c < hi
template <class Fn>
void for_each_set(const uint64_t *r, int lo, int hi, Fn fn)
{
    if (lo >= hi) return;
    const int wlo = lo >> 6, whi = (hi - 1) >> 6;
    for (int w = wlo; w <= whi; w++) {
        uint64_t bits = r[w];
        if (w == wlo) bits &= ~(uint64_t)0 << (lo & 63);
        if (w == whi && (hi & 63)) bits &= ((uint64_t)1 << (hi & 63)) - 1;
        while (bits) {
            fn((w << 6) + __builtin_ctzll(bits));
            bits &= bits - 1;
        }
    }
}

// Storage column of every check in the CTA-per-codeword kernel's message array
// (index = slot * M + column, bank = column mod 32).  The check phase walks columns in order
// and is conflict-free for any placement; in the variable phase lane l of a warp handles bit
// 32 w + l and touches, for its k-th check, the bank of THAT check's column -- with the checks
// in file order a random code gives ~3.4 wavefronts per access.  Greedy placement: take the
// checks one by one and give each the bank (with room left) where its edges collide least with
// the edges already placed in the same (bit-warp, k) access group; then a few passes move
// still-colliding checks to a better bank by swapping.  Relabels storage only: the arithmetic
// (edge order inside a bit's sums, slot order inside a check) is untouched.
void place_checks_for_banks(CodeTables &t)
{
    const int M = t.M, rows = M / 32, dv = std::max(t.dv_max, 1);
    const int n_groups = ((t.N + 31) / 32) * dv;
    // group of every CSR edge: (warp of its bit, rank of its check among the bit's checks)
    std::vector<int32_t> grp(t.E);
    for (int c = 0; c < t.N; c++)
        for (int q = t.col_ptr[c], k = 0; q < t.col_ptr[c + 1]; q++, k++)
            grp[t.edge_of_col[q]] = (c / 32) * dv + k;
    std::vector<uint16_t> cnt((size_t)n_groups * 32, 0);
    std::vector<int> bank(M, -1), fill(32, 0);
    auto cost = [&](int j, int b) {
        int c = 0;
        for (int e = t.row_ptr[j]; e < t.row_ptr[j + 1]; e++) c += cnt[(size_t)grp[e] * 32 + b];
        return c;
    };
    auto add = [&](int j, int b, int d) {
        for (int e = t.row_ptr[j]; e < t.row_ptr[j + 1]; e++) cnt[(size_t)grp[e] * 32 + b] += d;
    };
    for (int j = 0; j < M; j++) {
        int best = -1, best_cost = 1 << 30;
        for (int o = 0; o < 32; o++) {
            const int b = (j + o) & 31;                 // rotate the tie-break so banks fill evenly
            if (fill[b] >= rows) continue;
            const int c = cost(j, b);
            if (c < best_cost || (c == best_cost && fill[b] < fill[best])) { best = b; best_cost = c; }
        }
        bank[j] = best;
        fill[best]++;
        add(j, best, +1);
    }
    // improvement: swap a colliding check with a check of another bank when that lowers the total
    std::vector<std::vector<int>> members(32);
    for (int j = 0; j < M; j++) members[bank[j]].push_back(j);
    for (int pass = 0; pass < 40; pass++) {
        int moved = 0;
        for (int j = 0; j < M; j++) {
            const int b0 = bank[j];
            add(j, b0, -1);
            const int c0 = cost(j, b0);
            if (c0 == 0) { add(j, b0, +1); continue; }
            int best_b = -1, best_q = -1, best_gain = 0;
            for (int b = 0; b < 32; b++) {
                if (b == b0) continue;
                const int cj = cost(j, b);
                if (cj >= c0) continue;
                // partner from bank b that can move to b0 cheaply (sample a few)
                const auto &mb = members[b];
                for (size_t z = 0; z < mb.size() && z < 16; z++) {
                    const int q = mb[(j + pass * 13 + z * 7) % mb.size()];
                    add(q, b, -1);
                    const int gain = (c0 + cost(q, b)) - (cost(j, b) + cost(q, b0));
                    add(q, b, +1);
                    if (gain > best_gain) { best_gain = gain; best_b = b; best_q = q; }
                }
            }
            if (best_b >= 0) {
                add(best_q, best_b, -1);
                add(best_q, b0, +1);
                add(j, best_b, +1);
                bank[j] = best_b;
                bank[best_q] = b0;
                auto &ma = members[b0];
                auto &mb = members[best_b];
                *std::find(ma.begin(), ma.end(), j) = best_q;
                *std::find(mb.begin(), mb.end(), best_q) = j;
                moved++;
            } else {
                add(j, b0, +1);
            }
        }
        if (!moved) break;
    }
    std::vector<int> next(32, 0);
    for (int j = 0; j < M; j++) t.chk_pos[j] = bank[j] + 32 * next[bank[j]]++;
    // residual: access groups' wavefronts beyond one (0 = conflict-free variable phase)
    long long extra = 0;
    for (size_t g = 0; g < (size_t)n_groups; g++) {
        int mx = 0;
        for (int b = 0; b < 32; b++) mx = std::max<int>(mx, cnt[g * 32 + b]);
        extra += std::max(0, mx - 1);
    }
    t.bank_extra_wavefronts = (int)extra;
    t.bank_groups = n_groups;
}

// Proper edge colouring of a bipartite multigraph with 32 colours (every node of degree <= 32:
// Koenig), alternating-path algorithm.  eu / ev: left / right node of every edge.
void edge_color_32(const std::vector<int> &eu, const std::vector<int> &ev, int n_left, int n_right,
                   std::vector<int> &color)
{
    const int E = (int)eu.size();
    color.assign(E, -1);
    std::vector<int> at_u((size_t)n_left * 32, -1), at_v((size_t)n_right * 32, -1);   // node, colour -> edge
    std::vector<int> path;
    for (int e = 0; e < E; e++) {
        const int u = eu[e], v = ev[e];
        int a = 0, b = 0;
        while (at_u[(size_t)u * 32 + a] >= 0) a++;          // free at u
        while (at_v[(size_t)v * 32 + b] >= 0) b++;          // free at v
        if (a != b) {
            // walk the a/b alternating path from v (it cannot reach u) and swap its colours
            path.clear();
            int node = v, col = a;
            bool on_right = true;
            while (true) {
                const int f = on_right ? at_v[(size_t)node * 32 + col] : at_u[(size_t)node * 32 + col];
                if (f < 0) break;
                path.push_back(f);
                node = on_right ? eu[f] : ev[f];
                on_right = !on_right;
                col = (col == a) ? b : a;
            }
            for (int f : path) { at_u[(size_t)eu[f] * 32 + color[f]] = -1; at_v[(size_t)ev[f] * 32 + color[f]] = -1; }
            for (int f : path) {
                color[f] = (color[f] == a) ? b : a;
                at_u[(size_t)eu[f] * 32 + color[f]] = f;
                at_v[(size_t)ev[f] * 32 + color[f]] = f;
            }
        }
        color[e] = a;
        at_u[(size_t)u * 32 + a] = e;
        at_v[(size_t)v * 32 + a] = e;
    }
}

// Conflict-free layout for the register-table kernel (see code_tables.h: row_color).  A check's
// messages stay in its storage column's ROW (32 consecutive columns), slot-major as before, but
// inside the row each (column, slot) word sits at the bank its edge was coloured with: the check
// phase reads a permutation of one row (conflict-free whatever the permutation), the variable
// phase reads, per warp and k, 32 edges that all differ in colour.  Padded slots of a row take
// the colours left over.
void color_regular_rows(CodeTables &t)
{
    const int M = t.M, N = t.N, DC = t.dc_max, DV = t.dv_max, E = t.E;
    std::vector<int> eu(E), ev(E), color;
    for (int j = 0; j < M; j++)
        for (int e = t.row_ptr[j], s = 0; e < t.row_ptr[j + 1]; e++, s++) eu[e] = (t.chk_pos[j] / 32) * DC + s;
    for (int c = 0; c < N; c++)
        for (int q = t.col_ptr[c], k = 0; q < t.col_ptr[c + 1]; q++, k++)
            ev[t.edge_of_col[q]] = (c / 32) * DV + k;
    edge_color_32(eu, ev, (M / 32) * DC, ((N + 31) / 32) * DV, color);
    t.row_color.assign((size_t)M * DC, 0xFF);
    for (int j = 0; j < M; j++)
        for (int e = t.row_ptr[j], s = 0; e < t.row_ptr[j + 1]; e++, s++)
            t.row_color[(size_t)t.chk_pos[j] * DC + s] = (uint8_t)color[e];
    for (int row = 0; row < M / 32; row++)
        for (int s = 0; s < DC; s++) {
            bool used[32] = {};
            for (int l = 0; l < 32; l++) {
                const uint8_t c = t.row_color[(size_t)(row * 32 + l) * DC + s];
                if (c != 0xFF) used[c] = true;
            }
            int nb = 0;
            for (int l = 0; l < 32; l++) {
                uint8_t &c = t.row_color[(size_t)(row * 32 + l) * DC + s];
                if (c != 0xFF) continue;
                while (used[nb]) nb++;
                c = (uint8_t)nb;
                used[nb] = true;
            }
        }
    t.var_slot_colored.assign((size_t)DV * N, 0xFFFF);
    for (int j = 0; j < M; j++)
        for (int e = t.row_ptr[j], s = 0; e < t.row_ptr[j + 1]; e++, s++) {
            const int col = t.chk_pos[j];
            // position of this edge among its bit's edges (rows ascending)
            const int c = t.col_idx[e];
            int k = 0;
            for (int q = t.col_ptr[c]; q < t.col_ptr[c + 1]; q++, k++)
                if (t.edge_of_col[q] == e) break;
            t.var_slot_colored[(size_t)k * N + c] = (uint16_t)(s * M + (col & ~31) + color[e]);
        }
}

// Shared-memory layout of the warp-per-codeword kernel.  The strip of a codeword is dc_max rows
// of 32 words; the message of (check j, slot s) lives in row s at a bank chosen per EDGE.  Lane j
// reads row s in the check phase; in the variable phase lane l reads the k-th edge of bit
// l + 32 t.  Both are conflict-free iff the banks are a proper edge colouring of the bipartite
// multigraph {slots} x {(t, k) groups} whose edges are the Tanner-graph edges -- every node has
// degree <= 32, so 32 colours suffice (Koenig); found with the alternating-path algorithm.
void color_warp_layout(CodeTables &t)
{
    const int M = t.M, N = t.N, DC = t.dc_max, DV = t.dv_max, E = t.E;
    const int n_left = DC, n_right = 2 * DV;
    std::vector<int> eu(E), ev(E), color(E, -1);
    for (int j = 0; j < M; j++)
        for (int e = t.row_ptr[j], s = 0; e < t.row_ptr[j + 1]; e++, s++) eu[e] = s;
    for (int c = 0; c < N; c++)
        for (int q = t.col_ptr[c], k = 0; q < t.col_ptr[c + 1]; q++, k++)
            ev[t.edge_of_col[q]] = (c / 32) * DV + k;
    std::vector<int> at_u((size_t)n_left * 32, -1), at_v((size_t)n_right * 32, -1);   // node, colour -> edge
    for (int e = 0; e < E; e++) {
        const int u = eu[e], v = ev[e];
        int a = 0, b = 0;
        while (at_u[u * 32 + a] >= 0) a++;          // free at u
        while (at_v[v * 32 + b] >= 0) b++;          // free at v
        if (a != b) {
            // walk the a/b alternating path from v (it cannot reach u) and swap its colours
            std::vector<int> path;
            int node = v, col = a;
            bool on_right = true;
            while (true) {
                const int f = on_right ? at_v[node * 32 + col] : at_u[node * 32 + col];
                if (f < 0) break;
                path.push_back(f);
                node = on_right ? eu[f] : ev[f];
                on_right = !on_right;
                col = (col == a) ? b : a;
            }
            for (int f : path) { at_u[eu[f] * 32 + color[f]] = -1; at_v[ev[f] * 32 + color[f]] = -1; }
            for (int f : path) {
                color[f] = (color[f] == a) ? b : a;
                at_u[eu[f] * 32 + color[f]] = f;
                at_v[ev[f] * 32 + color[f]] = f;
            }
        }
        color[e] = a;
        at_u[u * 32 + a] = e;
        at_v[v * 32 + a] = e;
    }
    t.w_chk_pos.assign((size_t)DC * 32, 0);
    t.w_var_pos.assign((size_t)DV * 64, 0xFFFF);
    t.w_pos_edge.assign((size_t)DC * 32, -1);
    for (int s = 0; s < DC; s++) {
        std::vector<char> used(32, 0);
        std::vector<char> real(32, 0);
        for (int j = 0; j < M; j++) {
            const int e = t.row_ptr[j] + s;
            if (e < t.row_ptr[j + 1]) {
                t.w_chk_pos[(size_t)s * 32 + j] = (uint16_t)(s * 32 + color[e]);
                t.w_pos_edge[(size_t)s * 32 + color[e]] = e;
                used[color[e]] = 1;
                real[j] = 1;
            }
        }
        int nb = 0;                                  // padded slots take the banks left over in the row
        for (int j = 0; j < 32; j++) {
            if (real[j]) continue;
            while (used[nb]) nb++;
            t.w_chk_pos[(size_t)s * 32 + j] = (uint16_t)(s * 32 + nb);
            used[nb] = 1;
        }
    }
    for (int c = 0; c < N; c++)
        for (int q = t.col_ptr[c], k = 0; q < t.col_ptr[c + 1]; q++, k++) {
            const int e = t.edge_of_col[q];
            t.w_var_pos[(size_t)k * 64 + c] = (uint16_t)(eu[e] * 32 + color[e]);
        }
    // unused bit slots (bit degree <= k, or no such bit): 0x8000 | a bank that the real edges of the
    // access (t, k) = bits 32 t .. 32 t + 31, edge k leave free -- the kernel reads its zero row and
    // writes its dummy row there
    for (int k = 0; k < DV; k++)
        for (int tt = 0; tt < 2; tt++) {
            std::vector<char> used(32, 0);
            for (int l = 0; l < 32; l++) {
                const uint16_t pos = t.w_var_pos[(size_t)k * 64 + 32 * tt + l];
                if (pos != 0xFFFF) used[pos & 31] = 1;
            }
            int nb = 0;
            for (int l = 0; l < 32; l++) {
                uint16_t &pos = t.w_var_pos[(size_t)k * 64 + 32 * tt + l];
                if (pos != 0xFFFF) continue;
                while (used[nb]) nb++;
                pos = (uint16_t)(0x8000u | (unsigned)nb);
                used[nb] = 1;
            }
        }
}

}  // namespace

int build_code_tables(const int32_t *row_ptr_in, const int32_t *col_idx_in, int M, int N,
                      CodeTables &t)
{
    if (!row_ptr_in || !col_idx_in || M < 1 || N <= M || N > 65535) return 1;
    const int K = N - M;
    const int E = row_ptr_in[M];
    if (row_ptr_in[0] != 0 || E < 1) return 1;
    for (int j = 0; j < M; j++) {
        if (row_ptr_in[j + 1] < row_ptr_in[j]) return 1;
        for (int e = row_ptr_in[j]; e < row_ptr_in[j + 1]; e++)
            if (col_idx_in[e] < 0 || col_idx_in[e] >= N) return 1;
    }

    // F = H (original column order); it is column-swapped and row-reduced in place.
    BitMat F(M, N);
    for (int j = 0; j < M; j++)
        for (int e = row_ptr_in[j]; e < row_ptr_in[j + 1]; e++) F.set(j, col_idx_in[e]);

    t = CodeTables();
    t.M = M; t.N = N; t.K = K;
    t.pivots.assign(M, 0);
    t.col_origin.resize(N);
    for (int c = 0; c < N; c++) t.col_origin[c] = c;

    // L (unit lower) and U (unit upper), M x M, as bit rows -- reorderHMatrix's outputs.
    BitMat L(M, M), U(M, M);
    bool singular = false;

    for (int i = 0; i < M; i++) {
        // 'First' strategy (lib/ldpc_decoder_cb_impl.cc:269-277): first non-zero of row i
        // at or after the diagonal; the reference falls back to column 0 when there is none.
        int chosen = first_set_from(F.row(i), F.words, i, N);
        if (chosen < 0) { chosen = 0; singular = true; }
        t.pivots[i] = chosen;

        if (chosen != i) {   // swap columns i <-> chosen of F (and, by bookkeeping, of H) (:283-290)
            for (int r = 0; r < M; r++) {
                bool a = F.get(r, i), b = F.get(r, chosen);
                if (a != b) { F.flip(r, i); F.flip(r, chosen); }
            }
            std::swap(t.col_origin[i], t.col_origin[chosen]);
        }
        // column i of F -> column i of L (rows >= i) and of U (rows <= i) (:293-294)
        for (int r = i; r < M; r++) if (F.get(r, i)) L.set(r, i);
        for (int r = 0; r <= i; r++) if (F.get(r, i)) U.set(r, i);
        // add row i to every later row with a 1 in column i (:297-305).  Row i is zero
        // left of the diagonal by now, so only words from i/64 on can change.
        if (i < M - 1) {
            const uint64_t *src = F.row(i);
            const int w0 = i >> 6;
            for (int k = i + 1; k < M; k++) {
                if (F.get(k, i)) {
                    uint64_t *dst = F.row(k);
                    for (int w = w0; w < F.words; w++) dst[w] ^= src[w];
                }
            }
        }
    }
    if (singular) return 4;
    for (int i = 0; i < M; i++)
        if (!L.get(i, i) || !U.get(i, i)) return 4;

    // Re-ordered H: column c of H_perm is original column col_origin[c].
    std::vector<int32_t> new_col(N);
    for (int c = 0; c < N; c++) new_col[t.col_origin[c]] = c;
    t.row_ptr.assign(row_ptr_in, row_ptr_in + M + 1);
    t.col_idx.resize(E);
    t.E = E;
    for (int j = 0; j < M; j++) {
        int a = t.row_ptr[j], b = t.row_ptr[j + 1];
        for (int e = a; e < b; e++) t.col_idx[e] = new_col[col_idx_in[e]];
        std::sort(t.col_idx.begin() + a, t.col_idx.begin() + b);
        for (int e = a + 1; e < b; e++)
            if (t.col_idx[e] == t.col_idx[e - 1]) return 1;     // duplicate entry
    }
    // CSC: edges of a column with rows ascending
    t.col_ptr.assign(N + 1, 0);
    for (int e = 0; e < E; e++) t.col_ptr[t.col_idx[e] + 1]++;
    for (int c = 0; c < N; c++) t.col_ptr[c + 1] += t.col_ptr[c];
    t.edge_of_col.resize(E);
    {
        std::vector<int32_t> fill(t.col_ptr.begin(), t.col_ptr.end() - 1);
        for (int j = 0; j < M; j++)
            for (int e = t.row_ptr[j]; e < t.row_ptr[j + 1]; e++) t.edge_of_col[fill[t.col_idx[e]]++] = e;
    }

    // Generator: c = U^-1 L^-1 (B d) with B = H_perm[:, M:N]
    // (makeParityCheck, lib/ldpc_encoder_bc_impl.cc:275-294; the real-valued dgesv detour
    // equals GF(2) substitution mod 2 -- DESIGN.md "Encoder semantics").
    BitMat X(M, K);
    for (int j = 0; j < M; j++)
        for (int e = t.row_ptr[j]; e < t.row_ptr[j + 1]; e++)
            if (t.col_idx[e] >= M) X.set(j, t.col_idx[e] - M);
    for (int i = 0; i < M; i++) {                 // forward: X <- L^-1 X  (rows k < i are final)
        uint64_t *xi = X.row(i);
        for_each_set(L.row(i), 0, i, [&](int k) {
            const uint64_t *xk = X.row(k);
            for (int q = 0; q < X.words; q++) xi[q] ^= xk[q];
        });
    }
    for (int i = M - 1; i >= 0; i--) {            // back: X <- U^-1 X  (rows k > i are final)
        uint64_t *xi = X.row(i);
        for_each_set(U.row(i), i + 1, M, [&](int k) {
            const uint64_t *xk = X.row(k);
            for (int q = 0; q < X.words; q++) xi[q] ^= xk[q];
        });
    }
    t.kwords = (K + 31) / 32;
    t.mwords = (M + 31) / 32;
    t.P.assign((size_t)M * t.kwords, 0);
    t.Pt.assign((size_t)K * t.mwords, 0);
    for (int j = 0; j < M; j++)
        for (int k = 0; k < K; k++)
            if (X.get(j, k)) {
                t.P[(size_t)j * t.kwords + (k >> 5)] |= 1u << (k & 31);
                t.Pt[(size_t)k * t.mwords + (j >> 5)] |= 1u << (j & 31);
            }

    // Decoder tables (slot-major message layout).
    t.chk_deg.assign(M, 0);
    t.var_deg.assign(N, 0);
    for (int j = 0; j < M; j++) {
        int d = t.row_ptr[j + 1] - t.row_ptr[j];
        if (d > 255) return 1;
        t.chk_deg[j] = (uint8_t)d;
        t.dc_max = std::max(t.dc_max, d);
    }
    for (int c = 0; c < N; c++) {
        int d = t.col_ptr[c + 1] - t.col_ptr[c];
        if (d > 255) return 1;
        t.var_deg[c] = (uint8_t)d;
        t.dv_max = std::max(t.dv_max, d);
    }
    t.edge_slot.resize(E);
    t.chk_pos.resize(M);
    for (int j = 0; j < M; j++) t.chk_pos[j] = j;
    if (M >= 64 && M % 32 == 0) place_checks_for_banks(t);
    t.chk_deg_slot.assign(M, 0);
    for (int j = 0; j < M; j++) t.chk_deg_slot[t.chk_pos[j]] = t.chk_deg[j];
    if ((size_t)t.dc_max * M <= 65535) {
        t.chk_var.assign((size_t)t.dc_max * M, 0xFFFF);
        t.var_slot.assign((size_t)t.dv_max * N, 0xFFFF);
        for (int j = 0; j < M; j++)
            for (int e = t.row_ptr[j], s = 0; e < t.row_ptr[j + 1]; e++, s++) {
                t.chk_var[(size_t)s * M + t.chk_pos[j]] = (uint16_t)t.col_idx[e];
                t.edge_slot[e] = s * M + t.chk_pos[j];
            }
        for (int c = 0; c < N; c++)
            for (int q = t.col_ptr[c], k = 0; q < t.col_ptr[c + 1]; q++, k++)
                t.var_slot[(size_t)k * N + c] = (uint16_t)t.edge_slot[t.edge_of_col[q]];
    }
    if (M <= 32 && N <= 64) color_warp_layout(t);
    if (M >= 64 && M % 32 == 0 && (size_t)t.dc_max * M <= 65535) color_regular_rows(t);
    return 0;
}

}  // namespace ldpc535
