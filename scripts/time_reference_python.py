#!/usr/bin/env python
"""Worker of bench.py's `reference_python` leg: time the reference's own `EpisodeGenerator.generate_episodes`
(core/training/episode_generator.py:35-81 -> search.py:65-91 -> BasicNN.predict) from baseline/_ref on one host thread.

  --episodes 1   BASELINE config 1 (scripts/train.py:12-19 with one game at a time): S = 100, BasicNN, seeds 0
  --episodes 100 the reference's production batching (scripts/train.py:19)

Rounds of `generate_episodes` are repeated until --seconds have passed (the sample ends between two move steps); prints one JSON line.  A simulation = one iteration
of search.py:66 per tree, counted by wrapping `AlphaZeroSearch.run_simulations` (moves x S x trees)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

ap = argparse.ArgumentParser()
ap.add_argument("--episodes", type=int, default=1)
ap.add_argument("--sims", type=int, default=100)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--seconds", type=float, default=10.0)
args = ap.parse_args()

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_num_threads(1)
from alphazero_implementation.core.search.mcts import AlphaZeroSearch  # noqa: E402
from alphazero_implementation.core.training.episode_generator import EpisodeGenerator  # noqa: E402
from alphazero_implementation.models.games.connect4 import BasicNN  # noqa: E402
from simulator.game.connect import Config  # noqa: E402

torch.manual_seed(args.seed)
np.random.seed(args.seed)
counter = {"sims": 0}
_orig = AlphaZeroSearch.run_simulations


class _TimeUp(Exception):
    pass


def counted(self, nodes):
    if counter["sims"] and time.perf_counter() - t0 >= args.seconds:
        raise _TimeUp  # a round of E = 100 games lasts minutes: the sample ends between two move steps
    out = _orig(self, nodes)
    counter["sims"] += self.num_simulations * len(nodes)
    return out


AlphaZeroSearch.run_simulations = counted
state = Config(6, 7, 4).sample_initial_state()
gen = EpisodeGenerator(model=BasicNN(), num_simulations=args.sims, num_episodes=args.episodes, game_initial_state=state)
games = samples = 0
t0 = time.perf_counter()
try:
    while True:
        for ep in gen.generate_episodes():
            games += 1
            samples += len(ep.samples)
except _TimeUp:
    pass
dt = time.perf_counter() - t0
print(json.dumps({"episodes_concurrent": args.episodes, "num_simulations": args.sims, "seconds": dt, "games": games, "samples": samples,
                  "sims": counter["sims"], "sims_per_s": counter["sims"] / dt, "games_per_s": games / dt}))
