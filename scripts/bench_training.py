"""Row (f1): optimiser-step throughput of `Trainer` on one GPU, eager loop vs CUDA-graph replay (`trainer._GraphedTraining`).
The replay set is real self-play data (uniform evaluator, so that producing it takes a second); the nets are random-init.

  python scripts/bench_training.py [--games 16384] [--sims 50] [--batch-size 2048] [--nets resnet4x64,basic]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphazero_implementation_b200 as az  # noqa: E402
from alphazero_implementation_b200.replay import ReplayBuffer  # noqa: E402
from alphazero_implementation_b200.trainer import _GraphedTraining  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=16384)
ap.add_argument("--sims", type=int, default=50)
ap.add_argument("--batch-size", type=int, default=2048)
ap.add_argument("--nets", default="resnet4x64,basic")
ap.add_argument("--precisions", default="32-true,bf16-mixed")
args = ap.parse_args()

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
np.random.seed(0)
gen = az.EpisodeGenerator(model=az.UniformEvaluator(), num_simulations=args.sims, num_episodes=args.games,
                          game_initial_state=az.Config().sample_initial_state())
rb = ReplayBuffer(buffer_size=args.games, num_simulations=args.sims)
for batch in gen.generate_batches(quota=args.games):
    rb.extend(batch)
out = {"samples": rb.num_samples, "batch_size": args.batch_size, "runs": []}
for net in args.nets.split(","):
    for precision in args.precisions.split(","):
        for use_graph, cl in ((False, False), (False, True), (True, True)):
            torch.manual_seed(0)
            model = (az.BasicNN() if net == "basic" else az.ResNet(int(net.split("x")[0][6:]), int(net.split("x")[1]))).to(dev).train()
            opt = model.configure_optimizers()
            for group in opt.param_groups:
                group["capturable"] = True
            steps = _GraphedTraining(model, opt, args.batch_size, precision, dev, channels_last=cl)
            g = torch.Generator().manual_seed(1)
            steps.run(rb, 1, g, use_graph=use_graph)  # warm-up epoch (cuDNN algorithm selection, optimiser state)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            loss_sum, n = steps.run(rb, 1, g, use_graph=use_graph)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out["runs"].append(dict(net=net, precision=precision, cuda_graph=use_graph, channels_last=steps.channels_last, optimizer_steps=n, epoch_s=dt,
                                    ms_per_step=dt / n * 1e3, samples_per_s=rb.num_samples / dt, mean_loss=float(loss_sum) / n))
print(json.dumps(out, indent=1))
