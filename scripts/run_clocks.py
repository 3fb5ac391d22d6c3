"""Debug: where one warp of k_run_sims spends its cycles (library built with -DAZ_TRUNK_CLOCKS)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200 import _lib

E, S = 4096, 200
eng = az.Engine(num_games=E, num_simulations=S, device=0)
eng.reset_games()
u = torch.from_numpy(np.random.RandomState(0).random_sample((12, E))).cuda()
lib = _lib.load()
lib.az_debug_run_clocks.argtypes = [C.c_void_p, C.c_int]
for i in range(8):
    eng.run_simulations(S, 1)
    eng.sample_moves(u[i])
torch.cuda.synchronize()
lib.az_debug_run_clocks(None, 1)
eng.run_simulations(S, 1)
torch.cuda.synchronize()
buf = np.zeros(8, np.int64)
lib.az_debug_run_clocks(buf.ctypes.data, 0)
print("per simulation (cycles): descend %.0f  leaf/expand %.0f  backup %.0f   total %.0f ; mean depth of lane-0 tree %.2f" % (
    buf[0] / S, buf[1] / S, buf[2] / S, (buf[0] + buf[1] + buf[2]) / S, buf[3] / S))
