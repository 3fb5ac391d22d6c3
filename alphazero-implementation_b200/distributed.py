"""Multi-GPU plumbing: one process per GPU, games sharded, NCCL used for exactly two things
(SURVEY.md §8e): broadcasting new network weights and all-gathering finished episodes.

  broadcast_weights   <- `inference_model.load_state_dict(model.state_dict())`  search.py:22-25 / datamodule.py:100
  all_gather_episodes <- `buffer.append(episode)` into the shared replay deque    datamodule.py:29-30,57

Games never interact inside the search, so there is no collective on the per-simulation path.
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

_EP_FIELDS = ("ep_slot", "ep_step", "ep_len", "ep_offset", "ep_outcome")
_S_FIELDS = ("s_bb0", "s_bb1", "s_player", "s_counts")


def shard_range(total_games: int, rank: int, world: int, trainer_share: float | None = None, trainer_rank: int = 0) -> tuple[int, int]:
    """Static partition of game slots: rank r owns [lo, hi); sizes differ by at most one.
    `trainer_share` (0 <= share <= 1): the fraction of the games the trainer rank plays - that rank's GPU is time-shared between
    its self-play and the optimiser steps, so an equal split leaves the other ranks waiting for it; the rest is split evenly
    over the other ranks.  None (or one rank) = equal shards."""
    if trainer_share is None or world == 1:
        base, rem = divmod(total_games, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)
    if not 0.0 <= trainer_share <= 1.0:
        raise ValueError("trainer_share must be in [0, 1]")
    mine = min(total_games, int(round(total_games * trainer_share)))
    base, rem = divmod(total_games - mine, world - 1)
    sizes, k = [], 0
    for r in range(world):
        if r == trainer_rank:
            sizes.append(mine)
        else:
            sizes.append(base + (1 if k < rem else 0))
            k += 1
    lo = sum(sizes[:rank])
    return lo, lo + sizes[rank]


@torch.no_grad()
def broadcast_weights(model: torch.nn.Module, src: int = 0, group=None) -> int:
    """One flat broadcast of every parameter and buffer from `src`; returns the bytes sent."""
    tensors = [t for t in list(model.parameters()) + list(model.buffers()) if t.numel()]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    nbytes = 0
    by_dtype: dict[torch.dtype, list[torch.Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        nbytes += flat.numel() * flat.element_size()
        o = 0
        for t in ts:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return nbytes


def all_gather_episodes(local: dict, group=None, slot_offset: int | None = None) -> dict:
    """Merge every rank's drained episodes (dict of tensors as returned by `Engine.drain_episodes_device`)
    into one dict present on all ranks.  Two collectives: an all-gather of the (episodes, samples, slot offset)
    triples, then ONE all-gather of a byte buffer into which every field is packed (padded to the largest rank).
    Episode slots are made global by adding the owning rank's slot offset; sample offsets are rebased.
    Merge order: rank, then the rank's own order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    dev = local["ep_len"].device
    counts = torch.tensor([local["ep_len"].numel(), local["s_bb0"].numel(), slot_offset or 0], dtype=torch.int64, device=dev)
    all_counts = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(world, 3).cpu()
    ne, ns, offs = (all_counts[:, i].tolist() for i in range(3))
    max_e, max_s = max(ne + [1]), max(ns + [1])
    # layout of one rank's packed buffer: every field padded to the largest rank, regions 16-byte aligned
    layout, total = [], 0
    for fields, mx in ((_EP_FIELDS, max_e), (_S_FIELDS, max_s)):
        for f in fields:
            x = local[f]
            row = x.element_size() * math.prod(x.shape[1:])
            layout.append((f, total, mx, row, x.dtype, tuple(x.shape[1:])))
            total += (mx * row + 15) // 16 * 16
    buf = torch.zeros(total, dtype=torch.uint8, device=dev)
    for f, off, mx, row, dtype, tail in layout:
        x = local[f].contiguous()
        if x.shape[0]:
            buf[off:off + x.shape[0] * row] = x.view(-1).view(torch.uint8)
    gathered = torch.empty(world * total, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    gathered = gathered.view(world, total)
    out = {}
    for f, off, mx, row, dtype, tail in layout:
        cnt = ne if f in _EP_FIELDS else ns
        parts = [gathered[r, off:off + cnt[r] * row].view(dtype).view(cnt[r], *tail) for r in range(world)]
        out[f] = torch.cat(parts)
    # rebase sample offsets; make slots global
    s_base, e_pos = 0, 0
    for r in range(world):
        sl = slice(e_pos, e_pos + ne[r])
        out["ep_offset"][sl] += s_base
        if slot_offset is not None:
            out["ep_slot"][sl] += offs[r]
        s_base += ns[r]
        e_pos += ne[r]
    out["ep_rank"] = torch.cat([torch.full((c,), r, dtype=torch.int32, device=dev) for r, c in enumerate(ne)])
    return out
