// Built-in deterministic evaluators (stand-ins for `Model.predict`, models/base/model.py:54-74)
// used by the bit-exact parity configuration (BASELINE config 2).  Normative definition:
// oracle/evaluators.py — this file restates it for the device.
//
//   AZ_EVAL_UNIFORM  prior_c = fp32(1)/fp32(k) on the k legal columns, value = [0, 0]
//   AZ_EVAL_HASH     h = mix64((bb0*K1) ^ mix64(bb1 + K2) ^ player)
//                    w_c = 1 + ((h >> 8c) & 0xFF);  prior_c = fp32(w_c) / fp32(sum of legal w)
//                    v0 = (((h >> 56) & 0xFF) - 128) / 128;  value = [v0, -v0]
#pragma once
#include <stdint.h>

#include "c4_bitboard.cuh"

namespace azeval {

C4_HD uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

C4_HD uint64_t board_hash(uint64_t b0, uint64_t b1, int player) {
    return mix64((b0 * 0x9E3779B97F4A7C15ull) ^ mix64(b1 + 0xD1B54A32D192ED03ull) ^ (uint64_t)player);
}

C4_HD uint32_t hash_weight(uint64_t h, int c) { return 1u + (uint32_t)((h >> (8 * c)) & 0xFFu); }

C4_HD uint32_t hash_weight_total(uint64_t h, uint32_t legal) {
    uint32_t tot = 0;
#pragma unroll
    for (int c = 0; c < c4::W; ++c)
        if ((legal >> c) & 1u) tot += hash_weight(h, c);
    return tot;
}

// value for player 0 (player 1 gets the negation); exact dyadic k/128
C4_HD float hash_value0(uint64_t h) { return (float)((int)((h >> 56) & 0xFFu) - 128) * (1.0f / 128.0f); }

}  // namespace azeval
