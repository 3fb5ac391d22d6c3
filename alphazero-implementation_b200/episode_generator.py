"""`EpisodeGenerator` with the reference's surface (core/training/episode_generator.py:12-81): E concurrent
self-play games on the GPU engine, episodes yielded in the reference's order.

Per move step: S simulations for every slot (`AlphaZeroSearch.simulate`), then one `az_sample_moves`
launch that records the samples, draws the moves from the uniforms this process takes from NumPy's global
stream (one per slot in slot order — exactly what `np.random.choice` consumes in node.py:33), advances or
recycles the slots and moves finished games to the device episode ring.
"""
from __future__ import annotations

from typing import Callable, Generator

import numpy as np
import torch

from .engine import EpisodeBatch
from .episode import Episode, episodes_from_batch
from .game import State
from .search import AlphaZeroSearch


class EpisodeGenerator:
    def __init__(self, *, model, num_simulations: int, num_episodes: int, game_initial_state: State,
                 exploration_weight: float = 1.0, uniform_source: Callable[[int, int], np.ndarray] | None = None, **search_kwargs):
        self.search = AlphaZeroSearch(model=model, num_simulations=num_simulations, exploration_weight=exploration_weight,
                                      **search_kwargs)
        self.num_episodes = int(num_episodes)
        self.game_initial_state = game_initial_state
        self.num_players = game_initial_state.config.num_players
        self.uniform_source = uniform_source  # (step, E) -> float64[E]; default: np.random.random_sample
        self.last_run: dict = {}

    def update_inference_model(self, model):
        self.search.update_inference_model(model)

    # ------------------------------------------------------------------------------------------
    def iter_steps(self, initial_state: State | None = None, max_steps: int | None = None):
        """Run the self-play loop; after every move step yield (step, EpisodeBatch-or-None, rng_state_before_draw).
        Host buffers in, host buffers out: the step's uniforms are copied from pinned host memory and the
        finished episodes (if any) plus the ring counters are read back."""
        if initial_state is None:
            initial_state = self.game_initial_state
        E = self.num_episodes
        eng = self.search.engine_for(E)
        if eng.num_games != E:
            raise RuntimeError("engine was sized for a different number of games")
        eng.reset_games(initial_state.bb0, initial_state.bb1, initial_state.player)
        u_host = torch.empty(E, dtype=torch.float64).pin_memory()
        u_dev = torch.empty(E, dtype=torch.float64, device=eng.device)
        step = 0
        while max_steps is None or step < max_steps:
            self.search.simulate(eng)
            rng_state = np.random.get_state() if self.uniform_source is None else None
            u = np.random.random_sample(E) if self.uniform_source is None else np.asarray(self.uniform_source(step, E), np.float64)
            u_host.copy_(torch.from_numpy(u))
            u_dev.copy_(u_host, non_blocking=True)
            eng.sample_moves(u_dev)
            n_ep, _ = eng.episode_counts()
            batch = eng.drain_episodes() if n_ep else None
            yield step, batch, rng_state
            step += 1

    def generate_episodes(self, initial_state: State | None = None) -> Generator[Episode, None, None]:
        """Reference semantics: yield episodes in (move step, slot) order, stop after `num_episodes`,
        abandoning games in flight (episode_generator.py:48-81)."""
        count = 0
        S = self.search.num_simulations
        for step, batch, rng_state in self.iter_steps(initial_state):
            if batch is None:
                continue
            episodes = episodes_from_batch(batch, S)
            for i, ep in enumerate(episodes):
                count += 1
                if count >= self.num_episodes:
                    if rng_state is not None:
                        # the reference stops drawing after this slot: leave NumPy's global stream where it would be
                        np.random.set_state(rng_state)
                        np.random.random_sample(int(batch.ep_slot[i]) + 1)
                    self.last_run = dict(steps=step + 1, episodes=count)
                    yield ep
                    return
                yield ep

    def generate_batches(self, initial_state: State | None = None, max_steps: int | None = None,
                         quota: int | None = None) -> Generator[EpisodeBatch, None, None]:
        """Array-level variant for large E: yields `EpisodeBatch`es (flat host arrays) as games finish.
        With `quota`, truncates at the reference's stopping point."""
        count = 0
        for step, batch, _ in self.iter_steps(initial_state, max_steps):
            if batch is None:
                continue
            if quota is not None and count + len(batch) >= quota:
                keep = quota - count
                yield _truncate(batch, keep)
                return
            count += len(batch)
            yield batch


def _truncate(b: EpisodeBatch, keep: int) -> EpisodeBatch:
    ns = int(b.ep_len[:keep].sum())
    return EpisodeBatch(b.ep_slot[:keep], b.ep_step[:keep], b.ep_len[:keep], b.ep_offset[:keep], b.ep_outcome[:keep],
                        b.s_bb0[:ns], b.s_bb1[:ns], b.s_player[:ns], b.s_counts[:ns])
