"""AlphaZero iteration loop with the reference `Trainer.train(...)` signature (core/training/trainer.py:28-38),
without Lightning: self-play on the GPU engine -> replay buffer -> `Model.training_step` epochs -> weight sync.

Multi-GPU (one process per GPU): every rank plays its shard of the games; finished episodes are all-gathered
over NCCL into every rank's replay buffer; rank 0 runs the optimiser and broadcasts the new weights
(`update_inference_model`, search.py:22-25 / datamodule.py:100).  Like the reference (datamodule.py:89-101), self-play of
iteration k+1 runs on a background thread — here with its own CUDA stream — while iteration k trains.
"""
from __future__ import annotations

import time

import torch
import torch.distributed as dist

from .distributed import all_gather_episodes, broadcast_weights, shard_range
from .episode_generator import EpisodeGenerator
from .replay import ReplayBuffer


class Trainer:
    def __init__(self, model, device: int | None = None):
        self.model = model
        self.device_index = torch.cuda.current_device() if device is None else device
        self.device = torch.device("cuda", self.device_index)
        self.history: list[dict] = []

    def train(self, *, num_iterations: int, episodes_per_iter: int, simulations_per_episode: int, epochs_per_iter: int,
              initial_state, buffer_size: int, save_every_n_iterations: int = 0, batch_size: int = 32, seed: int = 0,
              overlap: bool = True, inference_dtype: torch.dtype | None = None, precision: str = "32-true"):
        """`overlap=True` reproduces the reference's pipeline (datamodule.py:89-101): the self-play of iteration k+1 runs on a
        background thread (own CUDA stream, weights as of the end of iteration k-1's training) while iteration k trains.
        `precision`: "32-true" (the reference's Lightning default) or "bf16-mixed" (forward / backward under bf16 autocast,
        fp32 master weights and optimiser state)."""
        import threading

        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
        lo, hi = shard_range(episodes_per_iter, rank, world)
        model = self.model.to(self.device)
        kw = {} if inference_dtype is None else dict(inference_dtype=inference_dtype)
        gen = EpisodeGenerator(model=model, num_simulations=simulations_per_episode, num_episodes=hi - lo,
                               game_initial_state=initial_state, device=self.device_index, **kw)
        replay = ReplayBuffer(buffer_size, simulations_per_episode, self.device)
        opt = model.configure_optimizers()
        g = torch.Generator().manual_seed(seed)
        play_stream = torch.cuda.Stream(device=self.device)
        box: dict = {}

        def play():  # the generator thread (EpisodeGeneratorThread.run, datamodule.py:24-30)
            try:
                torch.cuda.set_device(self.device)
                t = time.perf_counter()
                local = None
                with torch.cuda.stream(play_stream):
                    for batch in gen.generate_batches(quota=hi - lo):
                        local = batch if local is None else _concat(local, batch)
                    play_stream.synchronize()
                box["episodes"], box["selfplay_s"], box["error"] = local, time.perf_counter() - t, None
            except BaseException as exc:  # surfaced on the main thread
                box["error"] = exc

        def start():
            th = threading.Thread(target=play, daemon=True)
            th.start()
            return th

        thread = start() if overlap else None
        for it in range(num_iterations):
            t0 = time.perf_counter()
            if overlap:
                thread.join()
            else:
                play()
            if box.get("error") is not None:
                raise box["error"]
            local, selfplay_s = box["episodes"], box["selfplay_s"]
            t1 = time.perf_counter()
            if world > 1:
                merged = all_gather_episodes(_to_device(local, self.device), slot_offset=lo)
                replay.extend({k: merged[k] for k in ("ep_len", "ep_offset", "ep_outcome", "s_bb0", "s_bb1", "s_player", "s_counts")})
            else:
                replay.extend(local)
            t2 = time.perf_counter()
            # weight sync, then the next iteration's games start and overlap the training below
            nbytes = broadcast_weights(model, src=0) if world > 1 else 0
            gen.update_inference_model(model)
            if overlap and it + 1 < num_iterations:
                thread = start()
            losses = []
            if rank == 0:
                model.train()
                for _ in range(epochs_per_iter):
                    for x, pt, vt in replay.batches(model.input_layout, batch_size, True, g):
                        opt.zero_grad(set_to_none=True)
                        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=precision == "bf16-mixed"):
                            loss = model.training_step((x, pt, vt), 0)
                        loss.backward()
                        opt.step()
                        losses.append(loss.detach())
                model.eval()
            torch.cuda.current_stream(self.device).synchronize()
            t3 = time.perf_counter()
            self.history.append(dict(iteration=it, episodes=len(replay), samples=replay.num_samples, selfplay_s=selfplay_s,
                                     wait_for_selfplay_s=t1 - t0, gather_s=t2 - t1, train_s=t3 - t2, weight_bytes=nbytes,
                                     loss=float(torch.stack(losses).mean()) if losses else None))
        if world > 1:  # everyone ends with the trainer rank's final weights
            broadcast_weights(model, src=0)
        gen.update_inference_model(model)
        return self.history


def _concat(a, b):
    import numpy as np

    from .engine import EpisodeBatch

    off = np.concatenate([a.ep_offset, b.ep_offset + len(a.s_bb0)])
    return EpisodeBatch(np.concatenate([a.ep_slot, b.ep_slot]), np.concatenate([a.ep_step, b.ep_step]), np.concatenate([a.ep_len, b.ep_len]),
                        off, np.concatenate([a.ep_outcome, b.ep_outcome]), np.concatenate([a.s_bb0, b.s_bb0]),
                        np.concatenate([a.s_bb1, b.s_bb1]), np.concatenate([a.s_player, b.s_player]), np.concatenate([a.s_counts, b.s_counts]))


def _to_device(b, device):
    import numpy as np

    d = dict(ep_slot=b.ep_slot, ep_step=b.ep_step, ep_len=b.ep_len, ep_offset=b.ep_offset, ep_outcome=b.ep_outcome,
             s_bb0=b.s_bb0.view(np.int64), s_bb1=b.s_bb1.view(np.int64), s_player=b.s_player, s_counts=b.s_counts)
    return {k: torch.from_numpy(np.ascontiguousarray(v)).to(device) for k, v in d.items()}
