// Host-side, one-time code setup: the reference constructors' reorderHMatrix
// (lib/ldpc_decoder_cb_impl.cc:255-307 == lib/ldpc_encoder_bc_impl.cc:225-273) done on
// bit-packed rows, the encoder's generator P = (L U)^-1 B over GF(2), and the
// adjacency tables the decoder kernels read.  Plain C++, no CUDA.
#pragma once
#include <cstdint>
#include <vector>

namespace ldpc535 {

struct CodeTables {
    int M = 0, N = 0, K = 0, E = 0;
    int dc_max = 0, dv_max = 0;

    std::vector<int32_t> pivots;        // [M]   chosenCol of every re-ordering step
    std::vector<int32_t> col_origin;    // [N]   re-ordered column c holds original column col_origin[c]

    // re-ordered H
    std::vector<int32_t> row_ptr;       // [M+1]
    std::vector<int32_t> col_idx;       // [E]   columns ascending within a row (CSR edge order)
    std::vector<int32_t> col_ptr;       // [N+1]
    std::vector<int32_t> edge_of_col;   // [E]   CSR edge ids of a column, rows ascending

    // encoder: parity c_j = <P_j, d> mod 2
    int kwords = 0;                     // (K + 31) / 32
    std::vector<uint32_t> P;            // [M][kwords]  row-major, bit k of row j = P(j,k)
    std::vector<uint32_t> Pt;           // [K][mwords]  transposed: bit j of row k = P(j,k)
    int mwords = 0;                     // (M + 31) / 32

    // decoder: messages live slot-major, index s * M + j = slot s of check j
    std::vector<uint16_t> chk_var;      // [dc_max][M]  variable of (slot, check) or 0xFFFF
    std::vector<uint16_t> var_slot;     // [dv_max][N]  message index of the k-th check of a
                                        //              variable (checks ascending) or 0xFFFF
    std::vector<int32_t> chk_pos;       // [M]  storage column of check j in the message array (a
                                        //      permutation chosen to avoid shared-memory bank
                                        //      conflicts in the variable phase; identity for M < 64)
    std::vector<uint8_t> chk_deg_slot;  // [M]  degree of the check stored in column c
    int bank_extra_wavefronts = 0;      // variable-phase access groups' wavefronts beyond one
    int bank_groups = 0;
    // Register-table kernel for (3,6)-regular codes: on top of the storage columns, the 32 words of a
    // row (32 consecutive columns of one slot) are permuted by a proper 32-edge-colouring of the
    // bipartite multigraph {(row, slot)} x {(bit warp, k)}: BOTH phases are bank-conflict free.
    std::vector<uint8_t> row_color;     // [M][dc_max]  bank of (storage column, slot) inside its row; empty if not built
    std::vector<uint16_t> var_slot_colored;   // [dv_max][N]  message index under that layout
    // warp-per-codeword kernel (M <= 32, N <= 64): conflict-free shared-memory strip layout
    std::vector<uint16_t> w_chk_pos;    // [dc_max][32]  position of (slot s, check lane), padded slots included
    std::vector<uint16_t> w_var_pos;    // [dv_max][64]  position of the k-th edge of a bit, or 0x8000 | a bank free in that access
    std::vector<int32_t> w_pos_edge;    // [dc_max*32]   position -> CSR edge id or -1 (message dumps)
    std::vector<uint8_t> chk_deg;       // [M]
    std::vector<uint8_t> var_deg;       // [N]
    std::vector<int32_t> edge_slot;     // [E]  CSR edge id -> message index (message dumps)
};

// status: 0 ok, 1 invalid input, 4 singular (no pivot on some row)  (== LDPC535_* codes)
int build_code_tables(const int32_t *row_ptr, const int32_t *col_idx, int M, int N,
                      CodeTables &out);

}  // namespace ldpc535
