"""Device-resident replay buffer fed straight from the episode ring / all-gather.

Reference semantics (core/training/datamodule.py:57,114-130; models/base/model.py:76-82;
models/games/connect4/model.py:45-51): the buffer holds the last `buffer_size` EPISODES (a deque), the
training set is every sample of those episodes, inputs are `Model._states_to_tensor(states)`, policy
targets are dense [B,7] visit distributions (0 on illegal columns), value targets are [B,2] outcomes,
batches of 32, shuffled.  Here the samples stay on the GPU as bitboards + visit counts and are
expanded to planes by the `az_encode_states` kernel per minibatch.
"""
from __future__ import annotations

from collections import deque

import numpy as np
import torch

from .engine import EpisodeBatch
from .game import rules_engine


class ReplayBuffer:
    def __init__(self, buffer_size: int, num_simulations: int, device: torch.device | str = "cuda"):
        self.buffer_size = int(buffer_size)
        self.num_simulations = int(num_simulations)
        self.device = torch.device(device)
        self.episodes: deque[dict] = deque(maxlen=self.buffer_size)  # one dict of tensors per episode

    def __len__(self) -> int:
        return len(self.episodes)

    @property
    def num_samples(self) -> int:
        return sum(int(e["bb0"].numel()) for e in self.episodes)

    def extend(self, batch: EpisodeBatch | dict):
        """Append finished episodes in order (`buffer.append(episode)`, datamodule.py:29-30)."""
        if isinstance(batch, EpisodeBatch):
            d = dict(ep_len=batch.ep_len, ep_offset=batch.ep_offset, ep_outcome=batch.ep_outcome, s_bb0=batch.s_bb0.view(np.int64),
                     s_bb1=batch.s_bb1.view(np.int64), s_player=batch.s_player, s_counts=batch.s_counts)
            d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(self.device) for k, v in d.items()}
        else:
            d = {k: v.to(self.device) for k, v in batch.items()}
        lens = d["ep_len"].tolist()
        offs = d["ep_offset"].tolist()
        for e, (o, n) in enumerate(zip(offs, lens)):
            self.episodes.append(dict(bb0=d["s_bb0"][o:o + n], bb1=d["s_bb1"][o:o + n], player=d["s_player"][o:o + n],
                                      counts=d["s_counts"][o:o + n], outcome=d["ep_outcome"][e]))

    def tensors(self):
        """Flatten the deque (datamodule.py:114-122) -> (bb0, bb1, player, policy_target[B,7], value_target[B,2])."""
        eps = list(self.episodes)
        bb0 = torch.cat([e["bb0"] for e in eps]); bb1 = torch.cat([e["bb1"] for e in eps]); pl = torch.cat([e["player"] for e in eps])
        counts = torch.cat([e["counts"] for e in eps]).to(torch.float32)
        policy = counts / float(self.num_simulations - 1)  # Node.improved_policy (node.py:27), fp32 like torch.zeros(...)[i, col] = prob
        value = torch.cat([e["outcome"].to(torch.float32).expand(e["bb0"].numel(), 2) for e in eps])
        return bb0, bb1, pl, policy, value

    def batches(self, layout: int, batch_size: int = 32, shuffle: bool = True, generator: torch.Generator | None = None):
        """Yield (x, policy_target, value_target) minibatches on the device (DataLoader(batch 32, shuffle), datamodule.py:124-130)."""
        bb0, bb1, pl, policy, value = self.tensors()
        n = bb0.numel()
        perm = torch.randperm(n, generator=generator).to(self.device) if shuffle else torch.arange(n, device=self.device)
        eng = rules_engine()
        for i in range(0, n, batch_size):
            idx = perm[i:i + batch_size]
            x = eng.encode_states(bb0[idx].contiguous(), bb1[idx].contiguous(), pl[idx].contiguous(), layout)
            yield x, policy[idx], value[idx]
