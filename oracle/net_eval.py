"""TEST INFRASTRUCTURE — the reference's `Connect4Model.predict` (models/games/connect4/model.py:19-43) as a batched
CPU callback for the C oracle (`c4oracle.search / selfplay(py_eval=...)`): positions -> input tensor
(`BasicNN._states_to_tensor`, basic.py:41-47: raw grid -1 / 0 / 1; `CNNModel._states_to_tensor`, cnn.py:77-100: planes
empty / side to move / opponent, row 0 = bottom) -> `model.forward` in fp32 on the host cores -> softmax over the legal
columns only -> (priors [n,7] f32 with 0 on illegal columns, values [n,2] f32).

Only tests/ and bench.py's CPU legs import this (it is the CPU arm's evaluator: the same network the GPU arm runs,
evaluated by stock PyTorch CPU kernels exactly as the reference does, one `predict` per simulation step).
"""
from __future__ import annotations

import numpy as np
import torch

_BITS = np.array([[c * 7 + r for c in range(7)] for r in range(6)], dtype=np.uint64)  # [row][col] -> bit index


def occupancy(bb: np.ndarray) -> np.ndarray:
    """u64 bitboards [n] -> bool [n, 6, 7] (row 0 = bottom)."""
    return ((np.asarray(bb, np.uint64)[:, None, None] >> _BITS[None]) & np.uint64(1)).astype(bool)


def grid_f32(bb0, bb1) -> np.ndarray:
    """`State.grid` as float32 [n, 6, 7]: -1 empty, 0 / 1 owner (basic.py:45)."""
    o0, o1 = occupancy(bb0), occupancy(bb1)
    g = np.full(o0.shape, -1.0, np.float32)
    g[o0] = 0.0
    g[o1] = 1.0
    return g


def planes_f32(bb0, bb1, player) -> np.ndarray:
    """[n, 3, 6, 7] float32: empty, stones of the side to move, stones of the opponent (cnn.py:93-95)."""
    o0, o1 = occupancy(bb0), occupancy(bb1)
    pl = np.asarray(player).astype(bool)[:, None, None]
    mine = np.where(pl, o1, o0)
    theirs = np.where(pl, o0, o1)
    return np.stack([~(o0 | o1), mine, theirs], axis=1).astype(np.float32)


class TorchNetEvaluator:
    """py_eval(bb0, bb1, player, legal) -> (priors, values) through `model.forward` (fp32, CPU)."""

    def __init__(self, model, threads: int | None = None):
        self.model = model.get_inference_clone().float().cpu()
        self.raw_grid = type(model).__name__ == "BasicNN"
        self.calls = 0
        self.positions = 0
        if threads:
            torch.set_num_threads(int(threads))

    @torch.no_grad()
    def forward(self, bb0, bb1, player):
        x = grid_f32(bb0, bb1) if self.raw_grid else planes_f32(bb0, bb1, player)
        logits, values = self.model(torch.from_numpy(x))
        return logits.float(), values.float()

    @torch.no_grad()
    def __call__(self, bb0, bb1, player, legal):
        self.calls += 1
        self.positions += len(bb0)
        logits, values = self.forward(bb0, bb1, player)
        mask = torch.from_numpy(((np.asarray(legal, np.uint8)[:, None] >> np.arange(7, dtype=np.uint8)[None]) & 1).astype(bool))
        priors = torch.softmax(logits.masked_fill(~mask, float("-inf")), dim=1)  # F.softmax over the legal logits (model.py:29-35)
        return priors.numpy(), values.numpy()
