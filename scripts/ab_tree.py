"""A/B: fused tree kernel throughput at several tree counts (library chosen with AZ_ENGINE_LIB)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
S = 200
out = []
for E in (4096, 8192, 16384, 65536):
    eng = az.Engine(num_games=E, num_simulations=S, device=0)
    eng.reset_games()
    u = torch.from_numpy(np.random.RandomState(5).random_sample((48, E))).cuda()
    for i in range(5):
        eng.run_simulations(S, 1); eng.sample_moves(u[i])
    torch.cuda.synchronize()
    ms = 0.0
    n = 40 if E == 4096 else 10
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.run_simulations(S, 1); b.record(); eng.sample_moves(u[5 + i]); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    out.append(f"{E}: {E * S * n / ms * 1e3:.3e}")
    eng.close()
print("  ".join(out))
