"""TEST INFRASTRUCTURE — deterministic evaluators with the reference `Model.predict` signature.

`predict(states) -> (list[dict[Action,float]], list[list[float]])`
(`src/alphazero_implementation/models/base/model.py:54-74`).  The definitions
below are NORMATIVE for the build: the C oracle (`c4_oracle.c`) and the CUDA
engine (`csrc/az_eval.cuh`) restate them bit for bit.

Board key: two u64 bitboards, bit index = column*7 + row, row 0 = bottom.

kind 1  UNIFORM  prior_c = fp32(1)/fp32(k) for each of the k legal columns,
                 value = [0.0, 0.0].
kind 2  HASH     h   = mix64( (bb0*0x9E3779B97F4A7C15) ^ mix64(bb1 + 0xD1B54A32D192ED03) ^ player )
                 w_c = 1 + ((h >> 8c) & 0xFF)                    c = 0..6
                 prior_c = fp32(w_c) / fp32(sum of w over legal columns)   (one fp32 IEEE divide)
                 v0  = (((h >> 56) & 0xFF) - 128) / 128 ;  value = [v0, -v0]
                 mix64 = splitmix64 finaliser.
Values are dyadic (k/128) so fp64 value sums are exact; priors are exact fp32
numbers widened to Python doubles, as `prob.item()` does in the reference
(`models/games/connect4/model.py:38`).
"""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1
UNIFORM = 1
HASH = 2


def mix64(x: int) -> int:
    x &= M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


def grid_to_bitboards(grid) -> tuple[int, int]:
    """grid[row][col] in {-1,0,1}, row 0 = bottom -> (bb0, bb1), bit = col*7+row."""
    bb = [0, 0]
    g = np.asarray(grid)
    for r in range(g.shape[0]):
        for c in range(g.shape[1]):
            v = int(g[r, c])
            if v >= 0:
                bb[v] |= 1 << (c * 7 + r)
    return bb[0], bb[1]


def bitboards_to_grid(bb0: int, bb1: int, height: int = 6, width: int = 7) -> np.ndarray:
    g = np.full((height, width), -1, dtype=np.int8)
    for c in range(width):
        for r in range(height):
            b = 1 << (c * 7 + r)
            if bb0 & b:
                g[r, c] = 0
            elif bb1 & b:
                g[r, c] = 1
    return g


def board_hash(bb0: int, bb1: int, player: int) -> int:
    return mix64(((bb0 * 0x9E3779B97F4A7C15) & M64) ^ mix64((bb1 + 0xD1B54A32D192ED03) & M64) ^ player)


def evaluate(kind: int, bb0: int, bb1: int, player: int, legal_cols: list[int]):
    """-> ({col: prior as python float (exact fp32)}, [v0, v1])."""
    if kind == UNIFORM:
        p = float(np.float32(1.0) / np.float32(len(legal_cols)))
        return {c: p for c in legal_cols}, [0.0, 0.0]
    if kind == HASH:
        h = board_hash(bb0, bb1, player)
        w = {c: 1 + ((h >> (8 * c)) & 0xFF) for c in legal_cols}
        tot = np.float32(sum(w.values()))
        pri = {c: float(np.float32(w[c]) / tot) for c in legal_cols}
        v0 = (((h >> 56) & 0xFF) - 128) / 128.0
        return pri, [v0, -v0]
    raise ValueError(f"unknown evaluator kind {kind}")


class DeterministicEvaluator:
    """Duck-typed stand-in for a reference `Model` (only what `search.py:18,22-25,82-84` touch)."""

    def __init__(self, kind: int):
        self.kind = kind
        self.num_predict_calls = 0
        self.num_states = 0

    def get_inference_clone(self):
        return self

    def state_dict(self):
        return {}

    def load_state_dict(self, sd):
        return None

    def eval(self):
        return self

    def predict(self, states):
        self.num_predict_calls += 1
        self.num_states += len(states)
        policies, values = [], []
        for s in states:
            acts = s.actions
            bb0, bb1 = grid_to_bitboards(s.grid)
            pri, val = evaluate(self.kind, bb0, bb1, s.player, [a.column for a in acts])
            policies.append({a: pri[a.column] for a in acts})
            values.append(val)
        return policies, values
