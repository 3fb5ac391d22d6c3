"""Deterministic evaluators with the reference `Model` duck type (what `AlphaZeroSearch` touches:
`get_inference_clone`, `state_dict`/`load_state_dict`, `eval`, `predict`; search.py:18,22-25,82-84).

They are computed INSIDE the fused search kernel (`az_run_simulations`, csrc/az_eval.cuh), which is how
BASELINE config 2 (4096 games x 200 sims, uniform-prior evaluator, bit-exact visit counts) runs.
`predict` on explicit states goes through the same kernel: a 1-simulation search expands the root with
the evaluator's priors, and its backed-up value is the evaluator's value for the side to move.
"""
from __future__ import annotations

import numpy as np

from .engine import EVAL_HASH, EVAL_UNIFORM, Engine


class _BuiltinEvaluator:
    az_builtin_eval_kind: int = 0

    def get_inference_clone(self):
        return self

    def state_dict(self):
        return {}

    def load_state_dict(self, sd):
        return None

    def eval(self):
        return self

    def predict(self, states):
        n = len(states)
        eng = Engine(num_games=n, num_simulations=1)
        eng.set_roots(np.array([s.bb0 for s in states], np.uint64), np.array([s.bb1 for s in states], np.uint64),
                      np.array([s.player for s in states], np.uint8))
        eng.run_simulations(1, self.az_builtin_eval_kind)
        st = {k: v.cpu().numpy() for k, v in eng.root_stats().items()}
        policies, values = [], []
        for i, s in enumerate(states):
            policies.append({a: float(st["child_P"][i, a.column]) for a in s.actions})
            v = float(st["root_W"][i])  # value for the side to move; evaluators return [v0, -v0]
            v0 = (v if s.player == 0 else -v) + 0.0
            values.append([v0, -v0 + 0.0])
        eng.close()
        return policies, values


class UniformEvaluator(_BuiltinEvaluator):
    """prior = fp32(1)/fp32(k) on the k legal columns, value [0, 0]."""

    az_builtin_eval_kind = EVAL_UNIFORM


class HashEvaluator(_BuiltinEvaluator):
    """Pseudo-random priors and a dyadic value from a 64-bit hash of the position (exercises Q != 0)."""

    az_builtin_eval_kind = EVAL_HASH
