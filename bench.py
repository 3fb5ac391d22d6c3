#!/usr/bin/env python
"""Headline benchmark: MCTS simulations/s (and self-play games/s) for Connect4 self-play on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (N=1): BASELINE.json configs[1] — 4096 concurrent games x 200 simulations/move with the
deterministic uniform-prior evaluator (the bit-exact parity configuration).  One STEP = one move step of
the self-play loop for every game: `az_run_move_step` = 200 simulations per tree (`az_run_simulations`) and the move
(`az_sample_moves`: record samples, draw moves, recycle finished games) in ONE launch of `k_run_sims`.  With N > 1 every rank runs its own
4096 games (weak scaling, no collective on the data path); finished episodes are all-gathered over NCCL
after the timed region and that time is reported separately.

`value`   device-resident: uniforms already in HBM, CUDA events around each step, L2 flushed between steps.  Both GPU legs
          run 48 untimed burn-in move steps first, so that what is timed is the stationary mix of a running self-play loop
          (games at every stage, finishing and restarting) and not the cheaper opening moves all games share after a reset.
`e2e`     the public API (`EpisodeGenerator.generate_batches`): per step the uniforms come from pinned host
          memory and finished episodes + counters are read back to the host.
`--impl reference`  the CPU arm: the reference algorithm (oracle/c4_oracle.c, the C restatement pinned against
          the reference's own outputs; the reference itself is pure Python that cannot travel to the GPU box)
          on all host cores via OpenMP, same workload, bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NODE_BYTES = 20  # W f64 + N u32 + P f32 + CB u32


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=4096)
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--evaluator", default="uniform", choices=["uniform", "hash"])
    ap.add_argument("--lanes", type=int, default=0, help="lanes per tree: 8, 32 or 0 = engine default")
    ap.add_argument("--hot-nodes", type=int, default=None, help="nodes per tree kept in shared memory by the fused kernel (default: automatic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--burn-in", type=int, default=48, help="untimed move steps before the warm-up (games reach their stationary mix)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-tree-scaling", action="store_true")
    ap.add_argument("--net", default="basic_tc,resnet4x64", help="network-in-the-loop side measurements, comma separated: basic_tc | basic | resnetBxC | none")
    ap.add_argument("--net-games", type=int, default=16384, help="network-in-the-loop games on one GPU (BASELINE configs[2])")
    ap.add_argument("--net-games-sharded", type=int, default=65536, help="total games sharded over the ranks when N > 1 (configs[3])")
    ap.add_argument("--net-sims", type=int, default=800)
    ap.add_argument("--net-steps", type=int, default=2)
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread (the timed
    region of this workload lasts only milliseconds, far below nvidia-smi's own start-up time)."""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index: int):
        import threading

        self.samples, self.reason_bits, self.power = [], 0, []
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as exc:  # pragma: no cover
            self.err = repr(exc)
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv, i = self.nv, 0
        while not self.stop_flag:
            try:  # clock and throttle reasons every pass, power every fourth (each NVML call costs about a millisecond)
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                if i % 4 == 0:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            i += 1
            time.sleep(0.001)

    def stop(self) -> dict:
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "")]}
        self.stop_flag = True
        self.t.join(timeout=2)
        sm = sorted(self.samples)
        reasons = sorted(k for k, bit in self.REASONS.items() if self.reason_bits & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "power_w_max": max(self.power) if self.power else None,
                "samples": len(sm), "reasons": reasons, "source": "NVML polled every 2 ms during the timed region"}


def algorithmic_bytes(st: dict) -> int:
    """Bytes the tree kernels must move for the work counted in `st` ('algorithmic bytes'):
    select reads one 20-byte record per child scanned plus the root's N and CB (8 B) per simulation;
    expansion writes one 20-byte record per child created plus the leaf's CB (4 B);
    backup reads and writes W (f64) and N (u32) of every node on the path (24 B)."""
    return (st["children_scanned"] * NODE_BYTES + st["simulations"] * 8 + st["children_created"] * NODE_BYTES
            + st["evaluations"] * 4 + st["backup_nodes"] * 24)


def diff(a: dict, b: dict) -> dict:
    return {k: b[k] - a[k] for k in a}


# --------------------------------------------------------------------------------------------------
def cpu_arm(E: int, S: int, kind: int, target_s: float = 12.0):
    """The reference algorithm on the host cores (C restatement, OpenMP over trees).  A step = one move
    step of the E-game self-play loop; the sample is however many steps fit in ~target_s."""
    import numpy as np

    from oracle import c4oracle

    c4oracle.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = c4oracle.set_threads(cores)  # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers
    rng = np.random.RandomState(0)
    c4oracle.selfplay(E, S, rng.random_sample((2, E)), quota=10**9, eval_kind=kind)  # thread pool + page faults
    t0 = time.perf_counter()
    c4oracle.selfplay(E, S, rng.random_sample((5, E)), quota=10**9, eval_kind=kind)
    t1 = (time.perf_counter() - t0) / 5
    steps = max(1, min(3000, int(target_s / max(t1, 1e-4))))
    t0 = time.perf_counter()
    r = c4oracle.selfplay(E, S, rng.random_sample((steps, E)), quota=10**9, eval_kind=kind)
    dt = time.perf_counter() - t0
    return dict(sims_per_s=r.n_sims / dt, games_per_s=len(r.ep_slot) / dt, steps=steps, seconds=dt, cores=cores,
                ms_per_step=dt / steps * 1e3, sims=r.n_sims)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind = 1 if args.evaluator == "uniform" else 2
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_arm(args.games, args.sims, kind, target_s=1.0)
    res = cpu_arm(args.games, args.sims, kind, target_s=max(4.0, min(60.0, 0.5 * args.steps)))
    sample = f"{res['steps']} move steps of {args.games} games x {args.sims} sims from the initial position ({res['seconds']:.1f} s)"
    line = {
        "impl": "reference", "metric": "mcts_simulations_per_sec", "value": res["sims_per_s"], "unit": "sims/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "games_per_sec": res["games_per_s"],
        "config": {"workload": f"connect4_selfplay_{args.evaluator}_{args.games}x{args.sims}", "num_games": args.games,
                   "num_simulations": args.sims, "evaluator": args.evaluator, "c_puct": 1.0},
        "cpu_baseline": {"value": res["sims_per_s"], "unit": "sims/s", "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": res["sims_per_s"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference algorithm as the C restatement oracle/c4_oracle.c (pinned to the reference's own outputs), OpenMP over trees; "
                "the reference itself is pure Python (~2e3 sims/s/core, SURVEY.md §6) and its tree is not on the GPU box",
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def net_in_loop(args, device_index: int, peaks: dict, world: int = 1, rank: int = 0):
    """Side measurements with a network in the loop: self-play move steps of E games x S sims, one evaluator call per
    simulation step.  One GPU: BASELINE configs[2] (16384 games x 800 sims).  N > 1 GPUs: configs[3] — 65536 games sharded
    over the ranks (65536 / N per GPU), every rank times its own shard between barriers and the job's time is the max over
    ranks (the episode all-gather is timed in the main section; scripts/run_config4.py plays whole rounds).
    `resnetBx64`: bf16 ResNet-style net on the hand-written tcgen05 kernel (csrc/az_conv.cu: trunk + heads, leaf gather
    fused); other widths: conv/linear layers = cuDNN/cuBLAS tensor-core GEMMs.  `basic_tc`: the reference's BasicNN on the
    hand-written tcgen05 kernel (csrc/az_mlp.cu, leaf gather fused)."""
    import torch
    import torch.distributed as dist

    import alphazero_implementation_b200 as az

    out = []
    for spec in [s for s in args.net.split(",") if s and s != "none"]:
        kw = {}
        if spec == "basic_tc":
            model, name, flops, kw = az.BasicNN(), "BasicNN (tcgen05 fused MLP, bf16)", 577_024, dict(inference_dtype=torch.bfloat16)
        elif spec == "basic":
            model, name, flops = az.BasicNN(), "BasicNN (fp32, cuBLAS)", 577_024
        else:
            b, c = spec.replace("resnet", "").split("x")
            model = az.ResNet(num_res_blocks=int(b), num_channels=int(c))
            how = "tcgen05 fused kernel" if int(c) == 64 else "cuDNN"
            name, flops = f"ResNet {b}x{c} (bf16, {how})", model.flops_per_position()
        total_games = args.net_games if world == 1 else args.net_games_sharded
        E, S = total_games // world, args.net_sims
        torch.manual_seed(0)
        search = az.AlphaZeroSearch(model=model, num_simulations=S, device=device_index, **kw)
        eng = search.engine_for(E)
        eng.reset_games()
        g = torch.Generator(device="cpu").manual_seed(1 + rank)
        u = torch.rand((args.net_steps + 1, E), dtype=torch.float64, generator=g).to(eng.device)
        search.simulate(eng)  # warm-up move step (includes graph capture)
        eng.sample_moves(u[0])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        st0 = eng.stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.net_steps):
            search.simulate(eng)
            eng.sample_moves(u[i + 1])
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        st = diff(st0, eng.stats())
        sims, evals = float(st["simulations"]), float(st["evaluations"])
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=eng.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c = torch.tensor([sims, evals], dtype=torch.float64, device=eng.device)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            ms, sims, evals = float(t.item()), float(c[0].item()), float(c[1].item())
            eng.drain_episodes_device()
        else:
            eng.drain_episodes()
        sims_s = sims / ms * 1e3
        evals_s = E * world * S * args.net_steps / ms * 1e3  # every slot occupies a batch row each simulation
        rec = {"workload": f"connect4_selfplay_{spec}_{total_games}x{S}", "net": name, "num_games": total_games, "games_per_gpu": E,
               "n_gpus": world, "num_simulations": S, "move_steps": args.net_steps, "sims_per_s": sims_s,
               "us_per_sim_step": ms * 1e3 / (args.net_steps * S), "flops_per_position": flops, "tensor_tflops": evals_s * flops / 1e12,
               "tensor_pipe_frac_of_measured_bf16": evals_s * flops / 1e12 / (peaks["bf16_tflops"] * world),
               "leaf_eval_fraction": evals / max(1.0, sims)}
        out.append(rec)
        search._engine.close()
        del search, eng
        torch.cuda.empty_cache()
    return out or None


def tree_scaling(args, device_index: int, peaks: dict, kind: int):
    """The fused tree kernel at larger tree counts (same S, same evaluator): the kernel is latency-bound, so its fraction
    of the HBM roofline grows with the number of independent trees until the issue slots fill."""
    import numpy as np
    import torch

    import alphazero_implementation_b200 as az

    out = []
    for E in (16384, 65536):
        eng = az.Engine(num_games=E, num_simulations=args.sims, device=device_index, lanes_per_tree=args.lanes, hot_nodes=args.hot_nodes)
        eng.reset_games()
        u = torch.from_numpy(np.random.RandomState(5).random_sample((8, E))).to(eng.device)
        for i in range(3):
            eng.run_simulations(args.sims, kind)
            eng.sample_moves(u[i])
        torch.cuda.synchronize()
        st0 = eng.stats()
        ms = 0.0
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.run_simulations(args.sims, kind)
            b.record()
            eng.sample_moves(u[3 + i])
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        st = diff(st0, eng.stats())
        gbs = algorithmic_bytes(st) / (ms * 1e-3) / 1e9
        out.append({"trees": E, "num_simulations": args.sims, "kernel": "k_run_sims", "kernel_ms": ms / 5, "sims_per_s": st["simulations"] / ms * 1e3,
                    "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"]})
        eng.close()
        del eng
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import alphazero_implementation_b200 as az
    from alphazero_implementation_b200.engine import EVAL_HASH, EVAL_UNIFORM

    os.environ["NCCL_DEBUG"] = "WARN"  # NCCL_DEBUG=VERSION/INFO prints to stdout; stdout carries exactly one JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, K = max(3, args.warmup), max(1, args.steps)
    # Untimed burn-in before the warm-up: the games start together from the empty board, and the first ~40 move steps (opening
    # positions, no game finished yet) are cheaper than the stationary mix of a running self-play loop that the metric is about.
    # The CPU arm measures hundreds of steps, i.e. the same stationary mix.
    W += args.burn_in
    E, S = args.games, args.sims
    kind = EVAL_UNIFORM if args.evaluator == "uniform" else EVAL_HASH
    peaks = load_peaks()

    eng = az.Engine(num_games=E, num_simulations=S, device=local, lanes_per_tree=args.lanes, hot_nodes=args.hot_nodes)
    eng.reset_games()
    rng = np.random.RandomState(1000 + rank)
    u_all = torch.from_numpy(rng.random_sample((W + K, E))).to(eng.device)  # inputs resident in HBM
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        eng.run_move_step(S, kind, u_all[i])
        if eng.episode_counts()[0]:
            eng.drain_episodes_device()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    barrier()
    st0, l0 = eng.stats(), eng.launch_count
    sampler = ClockSampler(local)
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()  # L2 flush between timed iterations (outside the step's event pair)
        ev[i][0].record()
        eng.run_move_step(S, kind, u_all[W + i])  # az_run_move_step: 200 simulations per tree + the move, one launch
        ev[i][1].record()
        if (i + 1) % 16 == 0:  # keep the device ring from filling; not part of the device-resident step
            eng.drain_episodes_device()
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = eng.launch_count - l0 - K // 16 * 0
    st = diff(st0, eng.stats())
    sim_ms = [a.elapsed_time(b) for a, b in ev]
    step_ms = sim_ms
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=eng.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        cnt = torch.tensor([st["simulations"], st["episodes"]], dtype=torch.float64, device=eng.device)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        tot_sims, tot_eps = float(cnt[0].item()), float(cnt[1].item())
    else:
        tot_sims, tot_eps = float(st["simulations"]), float(st["episodes"])
    value = tot_sims / total_ms * 1e3
    eng.drain_episodes_device()

    # roofline of the dominant kernel (k_run_sims), this rank
    alg_bytes = algorithmic_bytes(st) / K
    k_ms = float(np.mean(sim_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"kernel": "k_run_sims", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (total_ms / K) if world == 1 else None,
                "bytes_per_sim": alg_bytes * K / max(1, st["simulations"]),
                "note": "latency-bound pointer chasing: one dependent load round per tree level; see DESIGN.md"}
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("k_run_sims_dram_bytes_per_launch")
        except Exception:
            pass

    # end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        gen = az.EpisodeGenerator(model=az.UniformEvaluator() if kind == EVAL_UNIFORM else az.HashEvaluator(), num_simulations=S,
                                  num_episodes=E, game_initial_state=az.Config(6, 7, 4).sample_initial_state(), device=local,
                                  lanes_per_tree=args.lanes)
        np.random.seed(7 + rank)
        g_eng = gen.search.engine_for(E)
        eps, e2e_t0, sims_before, bytes0 = 0, None, None, (0, 0)
        for step, batch, _ in gen.iter_steps(max_steps=W + K + 1):
            if step == W - 1:  # warm-up done: open the timed region on a quiet device (step W is already enqueued: it is
                barrier()      # finished by this barrier and counted as warm-up by the statistics snapshot below)
                sims_before = g_eng.stats()
                bytes0 = (gen.h2d_bytes, gen.d2h_bytes)
                e2e_t0 = time.perf_counter()
                continue
            if e2e_t0 is None:
                continue
            if batch is not None:
                eps += len(batch)
        barrier()
        e2e_s = time.perf_counter() - e2e_t0
        d = diff(sims_before, g_eng.stats())
        n_steps = d["moves"] // E
        h2d, d2h = gen.h2d_bytes - bytes0[0], gen.d2h_bytes - bytes0[1]
        t = torch.tensor([e2e_s], dtype=torch.float64, device=eng.device)
        c = torch.tensor([float(d["simulations"])], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
        e2e = {"value": float(c.item()) / float(t.item()), "unit": "sims/s", "h2d_bytes_per_step": int(h2d / max(1, n_steps)),
               "d2h_bytes_per_step": int(d2h / max(1, n_steps)), "steps": int(n_steps), "games_per_sec": eps / e2e_s * world,
               "api": "EpisodeGenerator.iter_steps: pinned-host uniforms in, finished episodes + ring counters out to pinned host memory every step; readback of step k overlaps step k+1"}

    # multi-GPU: all-gather of finished episodes over NCCL (config 4), timed apart from the data path
    allgather = None
    if world > 1:
        from alphazero_implementation_b200.distributed import all_gather_episodes

        all_gather_episodes(eng.drain_episodes_device())  # first call: NCCL channel set-up
        for _ in range(6):
            eng.run_move_step(S, kind, u_all[0])
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        drained = eng.drain_episodes_device()
        all_gather_episodes(drained)  # same payload once untimed: NCCL sizes its buffers for the message on first use
        barrier()  # the collective is timed from a common start, not from the slowest rank's arrival
        a0.record()
        merged = all_gather_episodes(drained)
        a1.record()
        torch.cuda.synchronize()
        allgather = {"ms": a0.elapsed_time(a1), "episodes": int(merged["ep_len"].numel()), "samples": int(merged["s_bb0"].numel())}

    arena_bytes = eng.device_bytes
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_arm(E, S, kind)
        cpu = {"value": r["sims_per_s"], "unit": "sims/s", "cores": r["cores"], "kind": "port",
               "sample": f"{r['steps']} move steps of {E} games x {S} sims from the initial position ({r['seconds']:.1f} s), "
                         "oracle/c4_oracle.c with OpenMP over trees"}
    extra, scaling = None, None
    eng.close()
    if rank == 0 and world == 1 and not args.no_tree_scaling:
        try:
            scaling = tree_scaling(args, local, peaks, kind)
        except Exception as exc:  # side measurements must not sink the headline
            scaling = {"error": repr(exc)}
    if args.net != "none":
        try:
            extra = net_in_loop(args, local, peaks, world, rank)
        except Exception as exc:
            if world > 1:
                raise  # a rank that skipped the collectives would hang the others
            extra = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": K, "warmup": W - args.burn_in,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "games_per_sec": tot_eps / total_ms * 1e3,
            "config": {"workload": f"connect4_selfplay_{args.evaluator}_{E}x{S}", "num_games_per_gpu": E, "num_simulations": S,
                       "evaluator": args.evaluator, "c_puct": 1.0, "lanes_per_tree": args.lanes or 8, "parallelism": f"games sharded x{world}",
                       "burn_in_steps": args.burn_in,
                       "l2": "flushed between timed steps (256 MiB memset outside the per-step event pair)",
                       "tree_arena_bytes": arena_bytes},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": wall_s, "work": {k: v / K for k, v in st.items()},
        }
        if allgather:
            line["episode_allgather"] = allgather
        if scaling:
            line["tree_scaling"] = scaling
        if extra:
            line["net_in_loop"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
