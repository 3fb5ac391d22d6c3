"""A/B: lanes per tree (8 / 16 / 32 = 4 / 2 / 1 trees per warp) of the fused tree kernel at BASELINE configs[1] (4096 x 200) and at
800 simulations - same steps, same uniforms; prints ms per move step and simulations / s."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200.engine import EVAL_UNIFORM

res = []
for E, S in ((4096, 200), (4096, 800), (16384, 800)):
    for G in (8, 16, 32):
        eng = az.Engine(num_games=E, num_simulations=S, lanes_per_tree=G)
        eng.reset_games()
        K = 12
        u = torch.from_numpy(np.random.RandomState(1000).random_sample((24 + K, E))).to(eng.device)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
        for i in range(24):
            eng.run_move_step(S, EVAL_UNIFORM, u[i])
            if i % 8 == 7:
                eng.drain_episodes_device()
        torch.cuda.synchronize()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
        for i in range(K):
            flush.zero_()
            ev[i][0].record()
            eng.run_move_step(S, EVAL_UNIFORM, u[24 + i])
            ev[i][1].record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / K
        st = eng.root_stats()
        chk = int(st["child_N"].sum().item())
        res.append(dict(trees=E, sims=S, lanes_per_tree=G, ms_per_step=ms, sims_per_s=E * S / ms * 1e3, checksum=chk))
        print(res[-1], flush=True)
        eng.close()
        del flush
print(json.dumps(res))
