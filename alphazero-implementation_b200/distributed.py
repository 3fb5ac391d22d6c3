"""Multi-GPU plumbing: one process per GPU, games sharded, NCCL used for exactly two things
(SURVEY.md §8e): broadcasting new network weights and all-gathering finished episodes.

  broadcast_weights   <- `inference_model.load_state_dict(model.state_dict())`  search.py:22-25 / datamodule.py:100
  all_gather_episodes <- `buffer.append(episode)` into the shared replay deque    datamodule.py:29-30,57

Games never interact inside the search, so there is no collective on the per-simulation path.
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


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


_ROW_BYTES = {"ep_slot": 4, "ep_step": 4, "ep_len": 4, "ep_offset": 8, "ep_outcome": 2, "s_bb0": 8, "s_bb1": 8, "s_player": 1, "s_counts": 28}
_HEADER = 32  # int64 [episodes, samples, slot offset, 0]


def _slab_layout(cap_e: int, cap_s: int):
    """Fixed layout of one rank's slab for given capacities: 32-byte header, then every field, regions 16-byte aligned."""
    layout, total = {}, _HEADER
    for fields, cap in ((_EP_FIELDS, cap_e), (_S_FIELDS, cap_s)):
        for f in fields:
            layout[f] = total
            total += (cap * _ROW_BYTES[f] + 15) // 16 * 16
    return layout, total


def _merge(parts: dict, ne: list, ns: list, offs: list, slot_offset, dev) -> dict:
    """Per-rank field tensors -> one dict: sample offsets rebased, slots made global.  Merge order: rank, then the rank's own order."""
    out = {f: torch.cat(ps) for f, ps in parts.items()}
    s_base, e_pos = 0, 0
    for r in range(len(ne)):
        sl = slice(e_pos, e_pos + ne[r])
        out["ep_offset"][sl] += s_base
        if slot_offset is not None:
            out["ep_slot"][sl] += offs[r]
        s_base += ns[r]
        e_pos += ne[r]
    out["ep_rank"] = torch.cat([torch.full((c,), r, dtype=torch.int32, device=dev) for r, c in enumerate(ne)])
    return out


def all_gather_episodes(local: dict, group=None, slot_offset: int | None = None, capacity: tuple[int, int] | None = None) -> dict:
    """Merge every rank's drained episodes (dict of tensors as returned by `Engine.drain_episodes_device`) into one dict present
    on all ranks (the analogue of `buffer.append(episode)` into the shared replay deque, datamodule.py:29-30,57).

    `capacity = (episodes, samples)` - the same on every rank, known before the call (e.g. the shard's episode quota and a bound on
    the samples) - selects the ONE-collective path: every rank packs its fields into a fixed-size slab whose 32-byte header carries
    its true counts, one all-gather moves the slabs, and the host reads the headers once, after the collective.  Should a rank
    exceed the capacity (seen by all ranks in the headers), or without `capacity`, the exact two-collective path runs: an all-gather
    of the counts, then one all-gather of slabs padded to the largest rank.
    Episode slots are made global by adding the owning rank's slot offset; sample offsets are rebased."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    dev = local["ep_len"].device
    my_e, my_s = int(local["ep_len"].numel()), int(local["s_bb0"].numel())

    def pack(cap_e: int, cap_s: int, header: bool):
        layout, total = _slab_layout(cap_e, cap_s)
        buf = torch.zeros(total, dtype=torch.uint8, device=dev)
        if header:
            buf[:_HEADER].view(torch.int64).copy_(torch.tensor([my_e, my_s, slot_offset or 0, 0], dtype=torch.int64))
        for f, off in layout.items():
            x = local[f].contiguous()
            if x.shape[0]:
                buf[off:off + x.shape[0] * _ROW_BYTES[f]] = x.view(-1).view(torch.uint8)
        gathered = torch.empty(world * total, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, buf, group=group)
        return layout, gathered.view(world, total)

    def unpack(layout, gathered, ne, ns):
        parts = {}
        for f, off in layout.items():
            cnt, x = (ne if f in _EP_FIELDS else ns), local[f]
            parts[f] = [gathered[r, off:off + cnt[r] * _ROW_BYTES[f]].view(x.dtype).view(cnt[r], *x.shape[1:]) for r in range(world)]
        return parts

    if capacity is not None:
        cap_e, cap_s = int(capacity[0]), int(capacity[1])
        fits = my_e <= cap_e and my_s <= cap_s
        if not fits:  # send the header only; every rank will see the overflow and take the exact path
            local_keep, local = local, {f: local[f][:0] for f in local}
            my_payload = (my_e, my_s)
        layout, gathered = pack(cap_e, cap_s, header=True)
        if not fits:
            local = local_keep
        heads = gathered[:, :_HEADER].contiguous().view(torch.int64).view(world, 4).cpu()  # the one host read, after the collective
        ne, ns, offs = (heads[:, i].tolist() for i in range(3))
        if all(e <= cap_e for e in ne) and all(x <= cap_s for x in ns):
            return _merge(unpack(layout, gathered, ne, ns), ne, ns, offs, slot_offset, dev)
    counts = torch.tensor([my_e, my_s, slot_offset or 0], dtype=torch.int64, device=dev)
    all_counts = torch.empty(world * 3, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_counts = all_counts.view(world, 3).cpu()
    ne, ns, offs = (all_counts[:, i].tolist() for i in range(3))
    layout, gathered = pack(max(ne + [1]), max(ns + [1]), header=False)
    return _merge(unpack(layout, gathered, ne, ns), ne, ns, offs, slot_offset, dev)
