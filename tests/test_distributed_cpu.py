"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: game sharding, weight broadcast, episode all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_episodes(rank, n_ep):
    rng = np.random.RandomState(100 + rank)
    lens = rng.randint(7, 43, n_ep).astype(np.int32)
    ns = int(lens.sum())
    off = np.zeros(n_ep, np.int64)
    if n_ep:
        off[1:] = np.cumsum(lens[:-1])
    return dict(ep_slot=torch.arange(n_ep, dtype=torch.int32), ep_step=torch.from_numpy(rng.randint(0, 50, n_ep).astype(np.int32)),
                ep_len=torch.from_numpy(lens), ep_offset=torch.from_numpy(off),
                ep_outcome=torch.from_numpy(rng.randint(-1, 2, (n_ep, 1)).astype(np.int8)).repeat(1, 2),
                s_bb0=torch.from_numpy(rng.randint(0, 2**40, ns)), s_bb1=torch.from_numpy(rng.randint(0, 2**40, ns)),
                s_player=torch.from_numpy(rng.randint(0, 2, ns).astype(np.uint8)),
                s_counts=torch.from_numpy(rng.randint(0, 200, (ns, 7)).astype(np.int32)))


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from alphazero_implementation_b200.distributed import all_gather_episodes, broadcast_weights, shard_range
    from alphazero_implementation_b200.models import BasicNN

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # 1. weight broadcast: rank 1 starts from different weights and must end up with rank 0's
        torch.manual_seed(rank)
        m = BasicNN()
        ref = None
        if rank == 0:
            ref = {k: v.clone() for k, v in m.state_dict().items()}
        nbytes = broadcast_weights(m, src=0)
        assert nbytes == sum(p.numel() * 4 for p in m.parameters())
        torch.manual_seed(0)
        expect = BasicNN().state_dict()
        for k, v in m.state_dict().items():
            assert torch.equal(v, expect[k]), k
        # 2. episode all-gather with different counts per rank (rank 1 has none in the second round)
        lo, hi = shard_range(10, rank, world)
        # exact two-collective path, one-collective path with a capacity that fits, and one that a rank overflows (falls back)
        for rnd, (n_ep, cap) in enumerate([((3, 5), None), ((4, 0), None), ((3, 5), (8, 8 * 42)), ((4, 0), (4, 4 * 42)), ((3, 5), (4, 4 * 42))]):
            local = _fake_episodes(rank, n_ep[rank])
            merged = all_gather_episodes(local, slot_offset=lo, capacity=cap)
            tot_e = sum(n_ep)
            assert merged["ep_len"].numel() == tot_e
            mine = slice(0, n_ep[0]) if rank == 0 else slice(n_ep[0], tot_e)
            assert torch.equal(merged["ep_len"][mine], local["ep_len"])
            assert torch.equal(merged["ep_slot"][mine], local["ep_slot"] + lo)
            assert (merged["ep_rank"][mine] == rank).all()
            # samples of every episode are where ep_offset says, for both ranks' episodes
            all_local = [_fake_episodes(r, n_ep[r]) for r in range(world)]
            e = 0
            for r in range(world):
                for i in range(n_ep[r]):
                    o, l = int(merged["ep_offset"][e]), int(merged["ep_len"][e])
                    lo_, ll = int(all_local[r]["ep_offset"][i]), int(all_local[r]["ep_len"][i])
                    assert l == ll
                    assert torch.equal(merged["s_bb0"][o:o + l], all_local[r]["s_bb0"][lo_:lo_ + ll])
                    assert torch.equal(merged["s_counts"][o:o + l], all_local[r]["s_counts"][lo_:lo_ + ll])
                    e += 1
        out.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        out.put((rank, repr(exc)))
        raise
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_all_games():
    from alphazero_implementation_b200.distributed import shard_range

    for total, world in ((65536, 8), (65536, 4), (10, 3), (7, 8)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(65536, 3, 8) == (24576, 32768)
    # weighted split: the trainer rank plays a smaller share, the rest is spread evenly; still a partition
    for total, world, share in ((65536, 8, 0.08), (100, 4, 0.1), (7, 3, 0.0), (10, 2, 1.0)):
        spans = [shard_range(total, r, world, share) for r in range(world)]
        assert spans[0] == (0, min(total, int(round(total * share))))
        assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1)) and spans[-1][1] == total
        others = [hi - lo for lo, hi in spans[1:]]
        assert max(others) - min(others) <= 1
    assert shard_range(65536, 0, 8, 0.078125) == (0, 5120) and shard_range(65536, 1, 8, 0.078125) == (5120, 5120 + 8631)


@pytest.mark.timeout(180)
def test_broadcast_and_allgather_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(150)
    results = sorted(out.get(timeout=5) for _ in range(2))
    assert results == [(0, "ok"), (1, "ok")], results
    assert all(p.exitcode == 0 for p in procs)
