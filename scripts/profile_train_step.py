"""Where a ResNet training step spends its GPU time: ms per optimiser step (batch 2048, synthetic planes) for a few cuDNN
settings, and the top kernels of one step from torch.profiler.

  python scripts/profile_train_step.py [--net resnet4x64] [--batch-size 2048]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphazero_implementation_b200 as az  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="resnet4x64")
ap.add_argument("--batch-size", type=int, default=2048)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B = args.batch_size


def make():
    torch.manual_seed(0)
    return az.ResNet(int(args.net.split("x")[0][6:]), int(args.net.split("x")[1])).to(dev).train()


x = (torch.rand(B, 3, 6, 7, device=dev) < 0.3).float()
pt = torch.softmax(torch.randn(B, 7, device=dev), 1)
vt = torch.randn(B, 2, device=dev).sign()


def run(model, opt, xx, autocast, steps=20):
    def one():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = model.training_step((xx, pt, vt), 0)
        loss.backward()
        opt.step()
    for _ in range(5):
        one()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(steps):
        one()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, one


res = {}
for name, bench, cl, ac in (("fp32_nchw", False, False, False), ("fp32_nchw_benchmark", True, False, False),
                            ("fp32_channels_last_benchmark", True, True, False), ("bf16_nchw_benchmark", True, False, True),
                            ("bf16_channels_last_benchmark", True, True, True)):
    torch.backends.cudnn.benchmark = bench
    model = make()
    if cl:
        model = model.to(memory_format=torch.channels_last)
    xx = x.contiguous(memory_format=torch.channels_last) if cl else x
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    ms, one = run(model, opt, xx, ac)
    res[name] = ms
    if name in ("fp32_nchw", "bf16_channels_last_benchmark"):
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            one()
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]
        res[name + "_top_kernels"] = [dict(name=e.key[:90], calls=e.count, us=round(e.device_time_total, 1)) for e in rows]
print(json.dumps(res, indent=1))
