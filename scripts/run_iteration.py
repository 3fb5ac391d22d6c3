"""BASELINE configs[4]: full AlphaZero iterations — sharded self-play, episode all-gather over NCCL, replay-buffer training
on rank 0, NCCL weight broadcast — timed end to end.  Launch with torchrun (one rank per GPU) or plain python (1 GPU).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/run_iteration.py [--games G] ...
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
from alphazero_implementation_b200.trainer import Trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096, help="episodes per iteration, all ranks together")
ap.add_argument("--sims", type=int, default=100)
ap.add_argument("--iterations", type=int, default=2)
ap.add_argument("--epochs", type=int, default=1)
ap.add_argument("--batch-size", type=int, default=1024)
ap.add_argument("--net", default="basic")
ap.add_argument("--precision", default="32-true", choices=["32-true", "bf16-mixed"])
ap.add_argument("--trainer-share", type=float, default=None, help="fraction of the games played by rank 0, which also trains (default: equal shards)")
ap.add_argument("--eager", action="store_true", help="optimiser steps as an eager loop instead of CUDA-graph replays")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
np.random.seed(1234 + rank)
model = az.BasicNN() if args.net == "basic" else az.ResNet(int(args.net.split("x")[0][6:]), int(args.net.split("x")[1]))
tr = Trainer(model, device=local)
t0 = time.perf_counter()
hist = tr.train(num_iterations=args.iterations, episodes_per_iter=args.games, simulations_per_episode=args.sims,
                epochs_per_iter=args.epochs, initial_state=az.Config(6, 7, 4).sample_initial_state(), buffer_size=args.games * 2,
                batch_size=args.batch_size, precision=args.precision, trainer_share=args.trainer_share, cuda_graph=not args.eager)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
# every rank must end with identical weights (rank 0 trained, everyone received the broadcast)
flat = torch.cat([p.detach().float().reshape(-1) for p in model.parameters()])
if world > 1:
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(ref, flat))], device=flat.device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    weights_identical = bool(same.item())
else:
    weights_identical = True
if rank == 0:
    print(json.dumps({"workload": "alphazero_iteration", "n_gpus": world, "net": args.net, "episodes_per_iter": args.games,
                      "simulations": args.sims, "iterations": args.iterations, "precision": args.precision, "trainer_share": args.trainer_share, "cuda_graph": not args.eager, "wall_s": wall, "weights_identical_on_all_ranks": weights_identical,
                      "history": hist}))
if world > 1:
    dist.destroy_process_group()
