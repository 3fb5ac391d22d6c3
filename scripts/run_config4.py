"""BASELINE configs[3]: one whole self-play round of G games x S sims/move with a network in the loop, games sharded
over the ranks, finished episodes all-gathered over NCCL (NVLink) into every rank.  Launch with torchrun (one rank per
GPU) or plain python (1 GPU).  The round follows the reference's stopping rule per rank (episode_generator.py:78-81): a
rank stops once its shard's quota of finished games is reached.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_config4.py [--games G] ...
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.distributed import all_gather_episodes, shard_range  # noqa: E402
from alphazero_implementation_b200.trainer import _concat, _to_device  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=65536, help="games (= episode quota) of the round, all ranks together")
ap.add_argument("--sims", type=int, default=800)
ap.add_argument("--net", default="resnet4x64")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
np.random.seed(1234 + rank)
if args.net == "basic":
    model, kw = az.BasicNN(), dict(inference_dtype=torch.bfloat16)
else:
    b, c = args.net.replace("resnet", "").split("x")
    model, kw = az.ResNet(int(b), int(c)), {}
lo, hi = shard_range(args.games, rank, world)
gen = az.EpisodeGenerator(model=model, num_simulations=args.sims, num_episodes=hi - lo,
                          game_initial_state=az.Config(6, 7, 4).sample_initial_state(), device=local, **kw)
eng = gen.search.engine_for(hi - lo)
for _ in gen.iter_steps(max_steps=1):  # warm-up: one move step (graph capture, allocator)
    pass
torch.cuda.synchronize()
if world > 1:
    all_gather_episodes(eng.drain_episodes_device())  # NCCL channel set-up
    dist.barrier()
st0 = eng.stats()
t0 = time.perf_counter()
mine = None
for batch in gen.generate_batches(quota=hi - lo):
    mine = batch if mine is None else _concat(mine, batch)
torch.cuda.synchronize()
t_play = time.perf_counter() - t0
st = {k: v - st0[k] for k, v in eng.stats().items()}
if world > 1:
    dist.barrier()
t1 = time.perf_counter()
merged = all_gather_episodes(_to_device(mine, eng.device), slot_offset=lo)
torch.cuda.synchronize()
t_gather = time.perf_counter() - t1
t = torch.tensor([t_play, time.perf_counter() - t0], dtype=torch.float64, device=eng.device)
c = torch.tensor([float(st["simulations"]), float(st["moves"])], dtype=torch.float64, device=eng.device)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
if rank == 0:
    n_ep, n_s = int(merged["ep_len"].numel()), int(merged["s_bb0"].numel())
    print(json.dumps({"workload": f"connect4_selfplay_round_{args.net}_{args.games}x{args.sims}", "n_gpus": world, "games_per_gpu": hi - lo,
                      "selfplay_s_max_over_ranks": float(t[0]), "round_s_incl_allgather": float(t[1]), "episode_allgather_s": t_gather,
                      "simulations": float(c[0]), "moves": float(c[1]), "sims_per_s": float(c[0]) / float(t[0]),
                      "games_per_s": n_ep / float(t[1]), "episodes_gathered": n_ep, "samples_gathered": n_s,
                      "mean_game_length": n_s / max(1, n_ep)}))
if world > 1:
    dist.destroy_process_group()
