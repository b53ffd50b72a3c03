#!/usr/bin/env python
"""bench.py -- self-play MCTS simulations/s (and positions/s) of the batched self-play search.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA library)
  python bench.py --impl reference [...]                       restated reference on the host cores

A "step" is one self-play move of every concurrent game: Gumbel sequential halving
(1 + budget lock-step simulations), targets, action selection, step, restart
(takzero/src/search/node/batched.rs:207-409, selfplay/src/main.rs:138-153,238-329).
Workload (BASELINE.json configs[2]): 6x6 Tak, half komi 4, 8192 concurrent games per GPU, k = 16
sampled actions, 256 simulations/move, random-init 16x256 ResNet (torch seed 123), synthetic openings.
For N > 1 (configs[3]) every rank runs its own 8192 games (weak scaling, no data-path collective);
NCCL broadcasts the weights once per generation and sums the counters.

`value`  : simulations/s with everything resident in HBM (tz_selfplay_move, no host buffers).
`e2e`    : the same loop through the host-buffer C ABI (injected Gumbel noise H2D from pinned memory,
           moves / improved-policy targets / terminals D2H every step).
`roofline`: the convolution kernel (tcgen05; one fused launch per network pass), sampled with CUDA events inside
           the timed region.
`cpu_baseline`: the restated reference (oracle C search + libtorch-CPU f32 forward, 128 games in
           lock-step like selfplay/src/main.rs:37) on this box's host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BOARD_N, HALF_KOMI = 6, 4
GAMES_PER_GPU = int(os.environ.get("TZ_BENCH_GAMES", "8192"))
SAMPLED_ACTIONS, SEARCH_BUDGET = 16, 256
WEIGHTED_RANDOM_PLIES, SAMPLE_THRESHOLD, ALLOWED_DROP, TARGET_BETA = 10, 32, 0.5, 0.25
CPU_GAMES = 128  # BATCH_SIZE of the reference process (selfplay/src/main.rs:37)
METRIC = "self-play MCTS simulations/sec, 6x6 Tak"
UNIT = "simulations/s"


def target_visitations(k: int, budget: int) -> float:
    # IMPROVED_POLICY_VISITATIONS = budget / log2(k) / k * (k - 1)   (selfplay/src/main.rs:47-52)
    steps = k.bit_length() - 1
    return float(budget // steps // k * (k - 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return (float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1590.0))), "measured (bf16_tflops_sustained)",
                float(p.get("bf16_tflops", 0.0)) or None)
    return 1590.0, "fallback", None


# ---------------------------------------------------------------------------------- restated reference (CPU)

def cpu_reference_setup():
    import torch

    from oracle import net_ref
    from oracle import oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = net_ref.Net(BOARD_N, seed=123)
    agent = net.as_oracle_agent()
    games = [O.new_opening(BOARD_N, HALF_KOMI, (1000 + g) % 8, (1000 + g) // 8 % 2) for g in range(CPU_GAMES)]
    batched = O.Batched(games)
    betas = [0.0] * CPU_GAMES
    return batched, agent, betas, cores


def cpu_lockstep_rate(batched, agent, betas, min_seconds: float, min_locksteps: int):
    """Lock-step simulations of CPU_GAMES games (BatchedMCTS::simulate) until the time bound."""
    t0 = time.perf_counter()
    done = 0
    while done < min_locksteps or time.perf_counter() - t0 < min_seconds:
        batched.simulate(agent, betas)
        done += 1
    dt = time.perf_counter() - t0
    return CPU_GAMES * done / dt, done, dt


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batched, agent, betas, cores = cpu_reference_setup()
    per_step = 2  # lock-step simulations per step: a bounded sample of one move's 257
    for _ in range(args.warmup):
        for _ in range(per_step):
            batched.simulate(agent, betas)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(per_step):
            batched.simulate(agent, betas)
    dt = time.perf_counter() - t0
    value = CPU_GAMES * per_step * args.steps / dt
    sample = (f"{per_step} lock-step simulations of {CPU_GAMES} games per step (of the {1 + SEARCH_BUDGET} one move "
              f"needs), oracle C search + libtorch-CPU f32 forward")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated CPU baseline (oracle port), not the reference binary: no Rust toolchain in the image",
    })


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": f"6x6 Tak (half komi {HALF_KOMI}) self-play, {GAMES_PER_GPU} concurrent games per GPU, "
                    f"Gumbel sequential halving k={SAMPLED_ACTIONS}, {SEARCH_BUDGET} sims/move, 16x256 ResNet",
        "games_per_gpu": GAMES_PER_GPU, "games_total": GAMES_PER_GPU * n_gpus,
        "sampled_actions": SAMPLED_ACTIONS, "search_budget": SEARCH_BUDGET,
        "weights": "random init, torch seed 123, BN mean 0 / var 1, empty SimHash set",
        "sharding": f"games sharded over {n_gpus} GPU(s), no data-path collective",
        "l2": "inputs of every lock-step (node arenas, states, input planes, 302 MB of logits) far larger than the "
              "126 MB L2, no explicit flush; the network's chunk activation sets are L2-resident by design",
    }


# ---------------------------------------------------------------------------------------------- our arm

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON result: anything a library prints on the way (NCCL's version
    # banner, torchrun notices) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(json_fd, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    if args.impl == "reference":
        run_reference(args, emit)
        return

    import numpy as np
    import torch

    from takzero_b200 import build as tz_build
    from takzero_b200 import capi, network, weights
    from takzero_b200 import distributed as tzd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (takzero_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tz_build.build()

    G, k, budget = GAMES_PER_GPU, SAMPLED_ACTIONS, SEARCH_BUDGET
    game_base, _ = tzd.shard(rank, world, G)
    m = capi.BatchedMCTS(BOARD_N, HALF_KOMI, G, device=local_rank, game_base=game_base)
    M = m.move_stride
    cuda_dev = torch.device("cuda", local_rank)

    # weights: every rank builds the tensor list, rank 0's values are broadcast over NCCL (one "generation")
    t_w = time.perf_counter()
    tensors = weights.random_init(BOARD_N, seed=123 + rank)
    if world > 1:
        tensors = tzd.broadcast_weights(tensors, src=0, device=cuda_dev)
    net_dtype = network.DTYPE_BF16 if os.environ.get("TZ_BENCH_DTYPE", "f16") == "bf16" else network.DTYPE_F16
    network.set_weights(m, tensors, net_dtype)
    weight_load_s = time.perf_counter() - t_w
    m.set_agent(capi.AGENT_NETWORK)
    m.new_openings(seed=1000)

    vis = target_visitations(k, budget)
    params = capi.SelfplayParams(k, budget, 0.0, WEIGHTED_RANDOM_PLIES, SAMPLE_THRESHOLD, ALLOWED_DROP, vis,
                                 TARGET_BETA, 20261018)

    def barrier():
        m.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident path: warm-up, then EXACTLY --steps timed moves ------------------------------------
    for _ in range(args.warmup):
        m.selfplay_move(params)
    barrier()
    c0 = m.counters()
    l0 = m.launch_count()
    clocks = ClockSampler(local_rank)
    clocks.start()
    m.profile_begin(8)
    m.timer_start()
    for _ in range(args.steps):
        m.selfplay_move(params)
    ms = m.timer_stop()
    barrier()
    prof = m.profile_end()
    clock_info = clocks.stop()
    status = m.status()
    if status:
        raise SystemExit(f"device search error bits 0x{status:x} during the timed region")
    c1 = m.counters()
    launches = m.launch_count() - l0
    sims = c1.simulations - c0.simulations
    evals = c1.evaluations - c0.evaluations
    known = c1.known - c0.known
    positions = G * args.steps
    if dist is not None:
        ms = tzd.max_over_ranks([ms], cuda_dev)[0]
        sims, evals, known, positions, launches = (
            int(x) for x in tzd.sum_counters([sims, evals, known, positions, launches], cuda_dev))
    value = sims / (ms / 1000.0)

    # ---- roofline of the dominant kernel (tower convolution, tcgen05) ---------------------------------
    peak, peak_kind, peak_burst = measured_peaks()
    cat = dict(zip(capi.PROFILE_CATEGORIES, range(8)))
    conv_ms = prof.ms[cat["conv_input"]] + prof.ms[cat["conv_tower"]] + prof.ms[cat["conv_policy"]]
    conv_launches = prof.launches[cat["conv_input"]] + prof.launches[cat["conv_tower"]] + prof.launches[cat["conv_policy"]]
    flops_pos = weights.flops_per_position(BOARD_N)
    achieved = (prof.positions * flops_pos / (conv_ms / 1000.0) / 1e12) if conv_ms > 0 else 0.0
    traffic = None
    ncu_path = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(ncu_path):
        with open(ncu_path) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {
        "kernel": "conv::k_conv3x3_pair (bf16 implicit GEMM, tcgen05 cta_group::2 / TMEM; one fused launch = the 34 "
                  "convolutions of a network pass)", "bound": "tensor", "achieved": achieved,
        "peak": peak, "peak_kind": peak_kind, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
        # the kernel runs inside a long step, so the sustained cuBLAS rate is the denominator; against cuBLAS's
        # burst rate (a GEMM timed alone, before the power cap bites) the same number is:
        "peak_burst": peak_burst, "frac_of_burst": achieved / peak_burst if peak_burst else None,
        "traffic": traffic,
        "sampled": {"locksteps": int(prof.locksteps), "positions": int(prof.positions),
                    "conv_launches": int(conv_launches), "conv_ms": conv_ms,
                    "flop_per_position": flops_pos},
        "kernel_ms_sampled": {name: prof.ms[i] for name, i in cat.items() if prof.launches[i]},
    }

    # ---- end to end through the host-buffer C ABI -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        rng = np.random.default_rng(7 + rank)
        pool = 2
        gumbel_pool = capi.pinned_array((pool, G, M), np.float32)
        gumbel_pool[:] = rng.gumbel(size=(pool, G, M)).astype(np.float32)
        betas = capi.pinned_array((G,), np.float32)
        betas[:] = 0.0
        randoms = capi.pinned_array((G,), np.uint64)
        sym = capi.pinned_array((G,), np.int32)
        adj = capi.pinned_array((G,), np.int32)
        plies = m.positions()["ply"].astype(np.int64)

        def host_move(i):
            nonlocal plies
            moves = m.gumbel_sequential_halving(betas, k, budget, gumbel_pool[i % pool])
            pol, ube, cnt = m.targets(vis, TARGET_BETA)
            randoms[:] = rng.integers(0, 1 << 62, size=G, dtype=np.uint64)
            sel = m.select_actions_in_selfplay(WEIGHTED_RANDOM_PLIES, SAMPLE_THRESHOLD, ALLOWED_DROP, randoms)
            play = np.where(plies < WEIGHTED_RANDOM_PLIES, sel, moves).astype(np.uint16)
            m.step(play)
            sym[:] = rng.integers(0, 8, size=G)
            adj[:] = rng.integers(0, 2, size=G)
            term = m.restart_terminal_envs(sym, adj)
            plies = np.where(term != 0, 2, plies + 1)
            return float(pol[0, 0]) + float(ube[0])

        h2d = 4 * G + 4 * G * M + 8 * G + 2 * G + 8 * G
        d2h = 2 * G + 4 * G * M + 4 * G + 4 * G + 2 * G + 4 * G + 5 * 4
        e2e_warm = max(1, min(args.warmup, 2))
        for i in range(e2e_warm):
            host_move(i)
        barrier()
        ce0 = m.counters()
        t0 = time.perf_counter()
        m.timer_start()
        for i in range(args.steps):
            host_move(e2e_warm + i)
        ems = m.timer_stop()
        barrier()
        wall_ms = 1000.0 * (time.perf_counter() - t0)
        ce1 = m.counters()
        esims = ce1.simulations - ce0.simulations
        if dist is not None:
            ems, wall_ms = tzd.max_over_ranks([ems, wall_ms], cuda_dev)
            esims = int(tzd.sum_counters([esims], cuda_dev)[0])
        e2e = {"value": esims / (ems / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": ems / args.steps, "wall_ms_per_step": wall_ms / args.steps,
               "api": "tz_gumbel_sequential_halving + tz_targets + tz_select_selfplay + tz_step + tz_restart_terminal "
                      "with pinned host buffers"}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        batched, agent, cbetas, cores = cpu_reference_setup()
        batched.simulate(agent, cbetas)  # warm-up (root expansion, thread pool)
        rate, n_lock, dt = cpu_lockstep_rate(batched, agent, cbetas, 15.0, 4)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_lock} lock-step simulations of {CPU_GAMES} games in {dt:.1f} s: oracle C search + "
                         f"libtorch-CPU f32 forward (torch threads = {cores}); restated reference, not the Rust binary"}

    if rank == 0:
        cfg = workload_config(world)
        cfg["arena_slots_per_game"] = m.arena_slots
        cfg["weight_broadcast_and_upload_s"] = weight_load_s
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if net_dtype == network.DTYPE_F16 else "bf16", "data": "synthetic",
            "config": cfg,
            "positions_per_s": positions / (ms / 1000.0), "nn_evals_per_s": evals / (ms / 1000.0),
            "known_fraction": known / sims if sims else 0.0,
            "pipeline_tensor_frac": (evals / (ms / 1000.0)) * flops_pos / 1e12 / (peak * world),
            "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "e2e": e2e,
            "cpu_baseline": cpu,
        }
        emit(out)
    m.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
