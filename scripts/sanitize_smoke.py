"""Small end-to-end run of every kernel family at tiny sizes, meant for checked builds: compute-sanitizer where it is
available, otherwise the library compiled with -DAZ_DEBUG_BOUNDS (device asserts on every tree-node index):
  (cd alphazero-implementation_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -shared \
     -Xcompiler -fPIC -DAZ_DEBUG_BOUNDS -o ../libaz_engine.so az_engine.cu az_mlp.cu az_conv.cu az_resnet_pipe.cu az_resnet_wide.cu az_cnn.cu)
  compute-sanitizer --tool memcheck python scripts/sanitize_smoke.py      (compute-sanitizer is closed on the round-2 GPU pool)"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import alphazero_implementation_b200 as az

torch.manual_seed(0)
E, S = 37, 24
for lanes in (8, 16, 32):
    eng = az.Engine(num_games=E, num_simulations=S, device=0, lanes_per_tree=lanes)
    eng.reset_games()
    u = torch.from_numpy(np.random.RandomState(0).random_sample((12, E))).cuda()
    for i in range(6):
        eng.run_simulations(S, 2)
        eng.sample_moves(u[i])
    for i in range(6, 12):
        eng.run_move_step(S, 1 + (i & 1), u[i])
    eng.root_stats()
    eng.drain_episodes()
    eng.close()
# every evaluator kernel: BasicNN (library fp32, tcgen05 bf16 / fp16), ResNet 64 channels (all five kernel variants), ResNet 128
# channels, CNNModel (conv + FC kernels); 9 roots = ragged batches for every tile size
for model, kw in ((az.BasicNN(), {}), (az.BasicNN(), dict(inference_dtype=torch.bfloat16)), (az.BasicNN(), dict(inference_dtype=torch.float16)),
                  (az.ResNet(1, 64), dict(trunk_variant=0)), (az.ResNet(1, 64), dict(trunk_variant=1, inference_dtype=torch.bfloat16)),
                  (az.ResNet(1, 64), dict(trunk_variant=2)), (az.ResNet(1, 64), dict(trunk_variant=3)), (az.ResNet(1, 64), dict(trunk_variant=4)),
                  (az.ResNet(2, 64), dict(trunk_variant=4, inference_dtype=torch.bfloat16)), (az.ResNet(1, 128), {}),
                  (az.CNNModel(), {}), (az.CNNModel(), dict(inference_dtype=torch.bfloat16))):
    s = az.AlphaZeroSearch(model=model, num_simulations=6, device=0, use_cuda_graph=False, **kw)
    nodes = [az.Node(az.Config(6, 7, 4).sample_initial_state()) for _ in range(9)]
    s.run_simulations(nodes)
    assert all(n.visit_count == 6 for n in nodes)
    s.close()
torch.cuda.synchronize()
print("sanitize_smoke ok")
