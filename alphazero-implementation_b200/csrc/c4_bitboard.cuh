// Connect4 (6x7, 4-in-line) rules on two 64-bit bitboards.
//
// Replaces the third-party `simulator.game.connect` State/Action calls the reference makes
// (SURVEY.md Appendix B): Action.sample_next_state() search.py:89 / node.py:38, State.actions
// models/games/connect4/model.py:29, State.has_ended node.py:63, State.reward node.py:68.
//
// Layout: bit index = column*7 + row, row 0 = bottom; each column owns 7 bits, the 7th (row 6)
// is a guard that stays zero so that shifted-AND line tests never wrap between columns.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define C4_HD __host__ __device__ __forceinline__
#else
#define C4_HD static inline
#endif

namespace c4 {

constexpr int H = 6;
constexpr int W = 7;
constexpr int STRIDE = 7;  // bits per column
constexpr uint64_t TOP_ROW = 0x810204081020ull;     // bit (7c+5) for c = 0..6
constexpr uint64_t BOARD = 0xFDFBF7EFDFBFull;       // 6 playable cells of every column
constexpr uint64_t LEGAL_MAGIC = 0x1041041041ull;   // sum 2^(36-6c): gathers bits 7c -> bits 36+c

C4_HD uint64_t col_cells(int c) { return 0x3Full << (STRIDE * c); }

// 4-in-line anywhere in one player's stones: vertical (1), horizontal (7), diagonals (6, 8)
C4_HD bool has4(uint64_t b) {
    uint64_t m;
    m = b & (b >> 1);
    if (m & (m >> 2)) return true;
    m = b & (b >> 7);
    if (m & (m >> 14)) return true;
    m = b & (b >> 6);
    if (m & (m >> 12)) return true;
    m = b & (b >> 8);
    return (m & (m >> 16)) != 0;
}

// branch-free variant for divergent warp code
C4_HD bool has4_nb(uint64_t b) {
    uint64_t v = b & (b >> 1), h = b & (b >> 7), d1 = b & (b >> 6), d2 = b & (b >> 8);
    return ((v & (v >> 2)) | (h & (h >> 14)) | (d1 & (d1 >> 12)) | (d2 & (d2 >> 16))) != 0;
}

C4_HD bool is_full(uint64_t occ) { return (occ & TOP_ROW) == TOP_ROW; }

// State.actions as a 7-bit column mask of a NON-ended position: column legal iff its top cell is empty
C4_HD uint32_t legal_mask(uint64_t occ) {
    uint64_t t = (~occ & TOP_ROW) >> 5;  // bit 7c set iff column c has room
    return (uint32_t)((t * LEGAL_MAGIC) >> 36) & 0x7Fu;
}

// the cell a stone dropped in column c lands on (column must have room): carry of occ + bottom(c)
C4_HD uint64_t drop_bit(uint64_t occ, int c) { return (occ + (1ull << (STRIDE * c))) & col_cells(c); }

// c-th legal column in ascending order -> column index (idx < popcount(legal))
C4_HD int nth_legal_column(uint32_t legal, int idx) {
    int col = 0;
#pragma unroll
    for (int c = 0; c < W; ++c) {
        uint32_t below = legal & ((1u << c) - 1u);
        int rank = 0;
#if defined(__CUDA_ARCH__)
        rank = __popc(below);
#else
        rank = __builtin_popcount(below);
#endif
        if (((legal >> c) & 1u) && rank == idx) col = c;
    }
    return col;
}

// has_ended / reward of an arbitrary position (root states handed in by the caller):
// winner = whoever owns a 4-in-line (player 0 tested first), else draw iff board full.
struct Terminal {
    bool ended;
    int8_t reward0;  // reward of player 0; player 1 gets the negation (+1 win / -1 loss / 0 draw)
};
C4_HD Terminal terminal_of(uint64_t b0, uint64_t b1) {
    Terminal t;
    bool w0 = has4(b0), w1 = has4(b1);
    t.ended = w0 || w1 || is_full(b0 | b1);
    t.reward0 = w0 ? 1 : (w1 ? -1 : 0);
    return t;
}

}  // namespace c4
