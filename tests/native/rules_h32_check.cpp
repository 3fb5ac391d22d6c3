// CPU check (test infrastructure): the half-word rules of csrc/c4_bitboard.cuh (namespace c4::h32, what k_env_step_h /
// k_state_info_h run) against the 64-bit formulation of the same header (what the tree kernels and the scalar kernels run), on
// positions from random playouts - every column byte 0..255, finished games and full columns included.  Prints what was covered; exit code 1 on the first mismatch.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../alphazero-implementation_b200/csrc/c4_bitboard.cuh"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}

struct Ref {
    uint64_t b0, b1;
    uint32_t pl, legal, ended, pair, status;
};

// k_env_step's body (csrc/az_engine.cu) on the 64-bit functions
static Ref ref_step(uint64_t b0, uint64_t b1, uint32_t pl, uint32_t cc) {
    c4::Terminal T = c4::terminal_of(b0, b1);
    uint32_t status = 1;
    if (!T.ended && cc < (uint32_t)c4::W && !(((b0 | b1) >> (c4::STRIDE * cc + 5)) & 1ull)) {
        const uint64_t bit = c4::drop_bit(b0 | b1, (int)cc);
        if (pl == 0) b0 |= bit; else b1 |= bit;
        const bool win = c4::has4_nb(pl ? b1 : b0);
        T.ended = win || c4::is_full(b0 | b1);
        T.reward0 = win ? (pl == 0 ? 1 : -1) : 0;
        pl ^= 1;
        status = 0;
    }
    Ref r;
    r.b0 = b0; r.b1 = b1; r.pl = pl; r.status = status;
    r.legal = T.ended ? 0u : c4::legal_mask(b0 | b1);
    r.ended = T.ended ? 1u : 0u;
    r.pair = (uint32_t)(uint8_t)T.reward0 | ((uint32_t)(uint8_t)(int8_t)-T.reward0 << 8);
    return r;
}

template <int MODE>
static int check(uint64_t b0, uint64_t b1, uint32_t pl, uint32_t cc) {
    const Ref r = ref_step(b0, b1, pl, cc);
    const uint32_t px = (pl & 1u) * 0x80u * 0x01000001u;  // the player's bit on the sign of bytes 0 and 3, as k_env_step_h forms it
    const uint32_t z = c4::h32::player_mask(px, 3);
    if (z != c4::h32::player_mask(px, 0) || c4::h32::player_mask(px, 1) != 0u || c4::h32::byte_of(0xA1B2C3D4u, 2) != 0xB2u) return 1;
    const c4::h32::Step s = c4::h32::env_step<MODE>((uint32_t)b0, (uint32_t)(b0 >> 32), (uint32_t)b1, (uint32_t)(b1 >> 32), z, cc);
    const uint64_t s0 = (uint64_t)s.lo0 | ((uint64_t)s.hi0 << 32), s1 = (uint64_t)s.lo1 | ((uint64_t)s.hi1 << 32);
    if (s0 != r.b0 || s1 != r.b1 || (pl ^ s.status ^ 1u) != r.pl || s.status != r.status || s.info.legal != r.legal || s.info.ended != r.ended || s.info.pair != r.pair) {
        printf("env_step<%d> mismatch: b0=%llx b1=%llx pl=%u col=%u\n", MODE, (unsigned long long)b0, (unsigned long long)b1, pl, cc);
        return 1;
    }
    const c4::Terminal T = c4::terminal_of(b0, b1);
    const c4::h32::Info i = c4::h32::state_info<MODE>((uint32_t)b0, (uint32_t)(b0 >> 32), (uint32_t)b1, (uint32_t)(b1 >> 32));
    const uint32_t pair = (uint32_t)(uint8_t)T.reward0 | ((uint32_t)(uint8_t)(int8_t)-T.reward0 << 8);
    if (i.ended != (T.ended ? 1u : 0u) || i.legal != (T.ended ? 0u : c4::legal_mask(b0 | b1)) || i.pair != pair ||
        c4::h32::has4<MODE>((uint32_t)b0, (uint32_t)(b0 >> 32)) != c4::has4(b0)) {
        printf("state_info<%d> mismatch: b0=%llx b1=%llx\n", MODE, (unsigned long long)b0, (unsigned long long)b1);
        return 1;
    }
    return 0;
}

int main(int argc, char **argv) {
    const long games = argc > 1 ? atol(argv[1]) : 20000;
    long checked = 0, wins = 0, draws = 0, full_cols = 0;
    for (long g = 0; g < games; ++g) {
        uint64_t b0 = 0, b1 = 0;
        uint32_t pl = 0;
        for (int ply = 0; ply < 60; ++ply) {  // keeps stepping after the end: finished positions are inputs too
            const uint32_t r = rnd();
            const uint32_t cc = (r & 15u) == 0 ? (r >> 8) & 255u : (r >> 8) % 7u;  // 1 in 16: any byte
            if (check<0>(b0, b1, pl, cc) || check<1>(b0, b1, pl, cc) || check<2>(b0, b1, pl, cc) || check<3>(b0, b1, pl, cc) || check<4>(b0, b1, pl, cc)) return 1;
            // every other column byte on this position now and then
            if ((r & 0xFF0000u) == 0)
                for (uint32_t c = 0; c < 256; ++c)
                    if (check<1>(b0, b1, pl, c)) return 1;
            ++checked;
            const Ref nx = ref_step(b0, b1, pl, cc);
            if (nx.status == 0 && nx.ended) (nx.pair ? wins : draws) += 1;
            if (nx.status == 1 && !c4::terminal_of(b0, b1).ended && cc < 7u) ++full_cols;
            b0 = nx.b0; b1 = nx.b1; pl = nx.pl;
        }
    }
    printf("%ld positions, %ld winning moves, %ld drawing moves, %ld moves into a full column\n", checked, wins, draws, full_cols);
    return 0;
}
