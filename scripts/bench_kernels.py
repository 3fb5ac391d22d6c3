"""Achieved HBM GB/s of the bandwidth-bound kernels (rules, plane encoders, softmax epilogue) at sizes far above L2,
and of the per-step tree kernels at config-3/4 sizes.  Prints one JSON object; CUDA events, 3 warm-ups, best of 10."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
from alphazero_implementation_b200.engine import (LAYOUT_GRID_F32, LAYOUT_PLANES_BF16, LAYOUT_PLANES_BF16_NHWC, LAYOUT_PLANES_F32,
                                                  POLICY_PRIORS, _ptr, _stream)

peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
eng = az.Engine(num_games=8, num_simulations=4)
dev = eng.device
N = 1 << 24


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


# random mid-game positions: play 12 random plies from the empty board on the GPU itself
g = torch.Generator(device="cpu").manual_seed(0)
bb0 = torch.zeros(N, dtype=torch.int64, device=dev); bb1 = torch.zeros_like(bb0); pl = torch.zeros(N, dtype=torch.uint8, device=dev)
for _ in range(12):
    col = torch.randint(0, 7, (N,), generator=g, dtype=torch.uint8).to(dev)
    r = eng.env_step(bb0, bb1, pl, col)
    bb0, bb1, pl = r["bb0"], r["bb1"], r["player"]
col = torch.randint(0, 7, (N,), generator=g, dtype=torch.uint8).to(dev)
out = {k: torch.empty_like(v) for k, v in eng.env_step(bb0, bb1, pl, col).items()}
res = {}


def env():
    eng.lib.az_env_step(eng.h, _ptr(bb0), _ptr(bb1), _ptr(pl), _ptr(col), N, _ptr(out["bb0"]), _ptr(out["bb1"]), _ptr(out["player"]),
                        _ptr(out["legal"]), _ptr(out["ended"]), _ptr(out["reward"]), _ptr(out["status"]), _stream())


ms = timeit(env)
b = N * (8 + 8 + 1 + 1 + 8 + 8 + 1 + 1 + 1 + 2 + 1)
res["k_env_step"] = dict(items=N, bytes_per_item=40, ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
info = eng.state_info(bb0, bb1)


def sinfo():
    eng.lib.az_state_info(eng.h, _ptr(bb0), _ptr(bb1), None, N, _ptr(info["legal"]), _ptr(info["ended"]), _ptr(info["reward"]), _stream())


ms = timeit(sinfo)
b = N * (16 + 4)
res["k_state_info"] = dict(items=N, bytes_per_item=20, ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
if "rules" in sys.argv[1:]:  # only the two rules kernels (AZ_RULES_MODE selects the formulation)
    import os
    res["rules_mode"] = os.environ.get("AZ_RULES_MODE", "default")
    print(json.dumps(res))
    eng.close()
    sys.exit(0)
M = 1 << 22
for name, layout, per in (("grid_f32", LAYOUT_GRID_F32, 168), ("planes_f32", LAYOUT_PLANES_F32, 504), ("planes_bf16", LAYOUT_PLANES_BF16, 252),
                          ("planes_bf16_nhwc8", LAYOUT_PLANES_BF16_NHWC, 672)):
    x = eng.encode_states(bb0[:M], bb1[:M], pl[:M], layout)
    ms = timeit(lambda: eng.lib.az_encode_states(eng.h, _ptr(bb0), _ptr(bb1), _ptr(pl), M, _ptr(x), layout, _stream()))
    b = M * (17 + per)
    res[f"k_encode[{name}]"] = dict(items=M, bytes_per_item=17 + per, ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
    del x
logits = torch.randn(N, 7, device=dev); legal = info["legal"]; pri = torch.empty_like(logits)
ms = timeit(lambda: eng.lib.az_masked_softmax(eng.h, _ptr(logits), _ptr(legal), N, _ptr(pri), _stream()))
b = N * 57
res["k_masked_softmax"] = dict(items=N, bytes_per_item=57, ms=ms, gbs=b / ms / 1e6, frac=b / ms / 1e6 / peak)
eng.close()
del bb0, bb1, pl, col, out, logits, pri

# per-step tree kernels with an external evaluator (uniform priors supplied from the device), config-3 size
for E, S in ((16384, 800), (65536, 800)):
    e = az.Engine(num_games=E, num_simulations=S)
    e.reset_games()
    pri = torch.full((E, 7), 1.0 / 7, device=e.device); val = torch.zeros((E, 2), device=e.device)
    x = e.gather_leaves(LAYOUT_PLANES_BF16)
    for _ in range(60):  # grow the trees a bit so the descent has depth
        e.select_leaves(); e.expand_backup(pri, val, POLICY_PRIORS)
    st0 = e.stats()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = [0.0, 0.0, 0.0]
    n = 100
    for _ in range(n):
        evs[0].record(); e.select_leaves(); evs[1].record(); e.gather_leaves(LAYOUT_PLANES_BF16, x); evs[2].record()
        e.expand_backup(pri, val, POLICY_PRIORS); evs[3].record(); torch.cuda.synchronize()
        for k in range(3):
            t[k] += evs[k].elapsed_time(evs[k + 1])
    st = {k: e.stats()[k] - st0[k] for k in st0}
    sel_b = st["children_scanned"] * 24 + st["simulations"] * (24 + 17 + 4 * 4) + st["levels"] * 4
    exp_b = st["children_created"] * 24 + st["evaluations"] * (36 + 17 + 8) + st["backup_nodes"] * 28
    res[f"k_select[{E}x{S}]"] = dict(us=t[0] / n * 1e3, gbs=sel_b / n / (t[0] / n) / 1e6, frac=sel_b / n / (t[0] / n) / 1e6 / peak)
    res[f"k_encode_leaves[{E}]"] = dict(us=t[1] / n * 1e3, gbs=E * (18 + 252) / (t[1] / n) / 1e6, frac=E * (18 + 252) / (t[1] / n) / 1e6 / peak)
    res[f"k_expand_backup[{E}x{S}]"] = dict(us=t[2] / n * 1e3, gbs=exp_b / n / (t[2] / n) / 1e6, frac=exp_b / n / (t[2] / n) / 1e6 / peak)
    res[f"tree_kernels_sims_per_s[{E}x{S}]"] = E / (sum(t) / n) * 1e3
    e.close()
    del e, x, pri, val
print(json.dumps(dict(hbm_peak_gbs=peak, kernels=res), indent=1))
