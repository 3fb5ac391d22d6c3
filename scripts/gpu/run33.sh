#!/bin/bash
# GPU session: biases loaded per block + skip values prefetched (libaz_jit.so) against the shipped kernel; bf16 = the headline format
for lib in alphazero-implementation_b200/libaz_engine.so _ab/libaz_jit.so; do echo $lib
AZ_ENGINE_LIB=$PWD/$lib python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az
for dt in (torch.bfloat16, torch.float16):
    for blocks in (4, 8):
        search = az.AlphaZeroSearch(model=az.ResNet(blocks, 64), num_simulations=64, use_cuda_graph=False, inference_dtype=dt)
        eng = search.engine_for(16384); eng.reset_games(); net = search._net
        eng.select_leaves()
        for _ in range(5): net.forward_leaves(eng)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(40): net.forward_leaves(eng)
        b.record(); torch.cuda.synchronize()
        print(dt, blocks, round(a.elapsed_time(b) / 40 * 1e3, 1), 'us', flush=True)
        eng.close()
PY
done
