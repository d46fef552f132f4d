#!/usr/bin/env python
"""bench.py -- env-steps/sec of IDQN on Foraging-8x8-2p-3f-v3 (BASELINE.json metric), B200 path vs the CPU reference loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one training iteration of the hot path on every GPU: E envs collect one episode each (<= 25 env steps,
fused forward + epsilon-greedy + transition + replay write per env step) followed by `updates_per_iteration` IDQN updates
(replay sample + target forward + fused forward/TD/backward + reduce + clip/Adam/target), default E updates = the
reference's one update per collected episode (marlbase/dqn/train.py:299-311).  Workload = BASELINE.json configs[1]:
4096 envs per GPU, batch_size 1024.  `value` counts real env transitions (sum of episode lengths) per second with the
whole loop device resident; `e2e` runs the same iteration with every env step's observations / rewards / flags crossing
pinned HOST buffers (the gym-style plugin boundary), copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENV_NAME = "lbforaging:Foraging-8x8-2p-3f-v3"
TIME_LIMIT = 25
LBF_KW = dict(rows=8, cols=8, n_agents=2, max_num_food=3, sight=8, time_limit=TIME_LIMIT)
N_AGENTS, OBS_DIM, N_ACTIONS, HIDDEN = 2, 15, 6, 128
FWD_FLOP_PER_ROW = 2 * (OBS_DIM * HIDDEN + HIDDEN * HIDDEN + HIDDEN * N_ACTIONS)  # 38 144
# dram__bytes_read.sum + dram__bytes_write.sum of one launch at this workload, from the ncu --set full capture summarised in
# profiles/r1_tc_pipeline.md (cold caches: ncu flushes L2 before every kernel)
TRAFFIC_NCU = {"tc_dqn_fwd_kernel": 9303552, "tc_dh1_kernel": 3705344, "tc_dw_kernel": 92672512}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="env instances per GPU")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--buffer", type=int, default=65536, help="replay ring capacity in episodes (per GPU)")
    ap.add_argument("--updates-per-iter", type=int, default=0, help="0 = one update per collected episode (= --envs)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gradient exchange inside the fused reduce + Adam kernel over NVLink peer memory (default), or one NCCL all-reduce per update")
    ap.add_argument("--tc-backward", type=int, default=1, help="1 = three-kernel tcgen05 training pipeline (default), 0 = fused FP32 FFMA training kernel")
    return ap.parse_args()


# ---- clocks ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(gpu_index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ---- reference arm: the CPU loop on all host cores ----------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_loop

    cores = os.cpu_count() or 1
    # each "step": every core runs `eps` iterations of (collect one 25-step episode with one env; one update at batch_size)
    eps = 3
    rounds = cpu_loop.run_parallel(LBF_KW, args.batch, cores, prefill=args.batch, n_episodes=eps, n_rounds=args.warmup + args.steps)
    timed = rounds[args.warmup:]
    # independent copies: whole-machine throughput = sum of the copies' own rates (a straggler does not stall the others)
    value = sum(r[2] for r in timed) / len(timed)
    secs = sum(r[1] for r in timed)
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} independent single-thread copies (torch.set_num_threads(1), marlbase/run.py:29) of the reference loop: per step "
                                   f"{eps} x (one 25-step episode with one env + one IDQN update at batch_size={args.batch}); pure-Python LBF restatement + PyTorch-CPU learner"},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, world):
    upi = args.updates_per_iter or args.envs
    return {"workload": f"IDQN {ENV_NAME} time_limit={TIME_LIMIT}, {args.envs} vectorised envs/GPU, batch_size={args.batch}, "
                        f"{upi} updates per iteration of {args.envs} episodes (BASELINE.json configs[1])",
            "envs_per_gpu": args.envs, "batch_size": args.batch, "updates_per_iteration": upi, "buffer_episodes": args.buffer,
            "parallelism": f"dp{world}" if world > 1 else "single", "global_batch": args.batch * world,
            "collective": (("gradient sum over NVLink peer memory inside the fused reduce + Adam kernel" if args.collective == "peer"
                            else "one NCCL all-reduce of 155 KB per update") if world > 1 else None),
            "l2": "inputs larger than L2: each update gathers 1024 random episodes from a %.0f MB replay ring" % (args.buffer * 3421 / 1e6)}


# ---- B200 arm --------------------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from codebase_b200 import _native as nat
    from codebase_b200.config import Config
    from codebase_b200.dqn.model import QNetwork
    from codebase_b200.dqn.train import Collector
    from codebase_b200.lbf import TrajStore
    from codebase_b200.utils.envs import make_env

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, B, T = args.envs, args.batch, TIME_LIMIT
    U = args.updates_per_iter or E
    env = make_env(args.seed, name=ENV_NAME, time_limit=T, parallel_envs=E, env_gid0=rank * E)
    cfg = Config(dict(optimizer="Adam", lr=3e-4, gamma=0.99, grad_clip=1.0, double_q=True, target_update_interval_or_tau=200, standardise_returns=False))
    model = QNetwork(env.single_observation_space, env.single_action_space, cfg, [128, 128], False, False, True, "cuda", max_batch=B, max_episode_length=T)
    if world > 1 and args.collective == "peer":
        try:
            model.attach_peers()
        except Exception as e:  # e.g. no peer access between the GPUs of this box: fall back to the NCCL exchange, and say so
            print(f"[bench] peer-memory exchange unavailable ({e}); using one NCCL all-reduce per update", file=sys.stderr)
            args.collective = "nccl"
        flag = torch.tensor([1 if args.collective == "peer" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)   # all ranks or none
        if int(flag.item()) == 0:
            args.collective = "nccl"
    if world > 1:
        dist.broadcast(model.theta, 0)
        model.hard_update()
    rb = TrajStore(args.buffer, env.n_agents, T, env.cfg.obs_dim, dev)
    coll = Collector(env, model, T)
    lib = nat.lib()
    nat.check(lib.marl_set_option(b"tensor_core_backward", C.c_int32(int(args.tc_backward))), "marl_set_option")
    state = dict(pos=0, updates=0)
    steps_dev = torch.zeros((), dtype=torch.int64, device=dev)
    # pinned host mirrors for the e2e leg
    nat_env = env.native
    h_obs = torch.empty_like(nat_env.obs, device="cpu").pin_memory()
    h_rew = torch.empty_like(nat_env.rew, device="cpu").pin_memory()
    h_done = torch.empty_like(nat_env.done, device="cpu").pin_memory()
    h_trunc = torch.empty_like(nat_env.trunc, device="cpu").pin_memory()
    h_loss = torch.empty(6, dtype=torch.float32).pin_memory()
    d_obs_in = torch.empty_like(nat_env.obs)

    def do_updates():
        n_valid = min(state["pos"], args.buffer)
        if world == 1 or args.collective == "peer":   # peer: the gradient sum of all ranks happens inside the fused reduce + Adam kernel
            model.update_n(rb, B, n_valid, args.seed + 7919 * rank, state["updates"], U)
        else:
            for u in range(U):
                nat.check(lib.marl_replay_sample(C.c_uint64(args.seed + 7919 * rank), C.c_uint64(state["updates"] + u), C.c_int32(B), C.c_int32(n_valid),
                                                 nat.ptr(model._idx), nat.stream_ptr()), "marl_replay_sample")
                model.update_grads(rb, model._idx[:B])
                dist.all_reduce(model.grad)  # [sum-gradients | loss numerator | filled count], one exchange per update
                model.update_apply()
        state["updates"] += U

    def iteration(host_boundary: bool):
        slot0 = state["pos"] % args.buffer
        eps = 0.5
        if not host_boundary:
            final_len, _ = coll.collect(rb, slot0, eps)
        else:  # every env step's observations enter through pinned host memory and its results return to it
            nat_env.reset(traj=rb, slot0=slot0)
            h_obs.copy_(nat_env.obs, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            for _ in range(T):
                d_obs_in.copy_(h_obs, non_blocking=True)                                  # H2D: this step's inputs
                model.q_values(d_obs_in, out=coll.q)
                nat_env.rollout_step(coll.q, policy=1, epsilon=eps, traj=rb, slot0=slot0)
                h_obs.copy_(nat_env.obs, non_blocking=True); h_rew.copy_(nat_env.rew, non_blocking=True)      # D2H: this step's results
                h_done.copy_(nat_env.done, non_blocking=True); h_trunc.copy_(nat_env.trunc, non_blocking=True)
                torch.cuda.current_stream().synchronize()                                 # the host "sees" the step before the next one
            final_len = nat_env.final_len
        steps_dev.add_(final_len.sum())
        state["pos"] += E
        do_updates()
        if host_boundary:
            h_loss.copy_(model._metrics, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(k, host_boundary):
        steps_dev.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            iteration(host_boundary)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        n = steps_dev.clone()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(n, op=dist.ReduceOp.SUM)
        barrier()
        return float(ms.item()), int(n.item())

    for _ in range(max(args.warmup, 3)):
        iteration(False)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, n_steps = timed(args.steps, False)
    clocks = sampler.stop() if sampler else None
    # roofline leg: one more iteration with the library's CUDA events around (and inside) the training pass of its first 1024
    # updates -- outside the headline region, because in-stream events serialise launches that otherwise overlap (PDL)
    model.timing(True)
    iteration(False)
    torch.cuda.synchronize()
    train_ms, train_n = model.timing(False)
    kernel_ms, kernel_n = model.timing_kernels()
    e2e = None
    if not args.no_e2e:
        iteration(True)
        ms_e, n_e = timed(args.steps, True)
        e2e = {"value": n_e / (ms_e / 1e3), "unit": "env-steps/s",
               "h2d_bytes_per_step": T * h_obs.numel() * 4,
               "d2h_bytes_per_step": T * (h_obs.numel() * 4 + h_rew.numel() * 4 + h_done.numel() + h_trunc.numel()) + 8}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n_steps / (ms / 1e3)
    rows = N_AGENTS * (T + 1) * B
    train_flop = 3 * rows * FWD_FLOP_PER_ROW             # online forward (1x) + backward (2x) of one update
    train_avg_s = (train_ms / max(train_n, 1)) / 1e3
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = n_sm * 128 * 2 * sm_mhz * 1e6 / 1e12    # FFMA lanes x 2 flop x clock actually seen during the run
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tensor_peak = peaks.get("bf16_tflops_sustained", 1443.2)   # sustained figure: the kernel is timed inside a long step
    if kernel_n:
        # tensor-core training pass: three kernels; the roofline object describes the slowest one
        names = ["tc_dqn_fwd_kernel", "tc_dh1_kernel", "tc_dw_kernel"]
        D_in, A_out = 15, 6
        flops = [rows * FWD_FLOP_PER_ROW,                                  # online forward
                 rows * 2 * 128 * 128,                                     # dH1 = dH2 x W2
                 rows * 2 * (128 * 128 + 128 * D_in + A_out * 128 + 128 + 128 + A_out)]   # dW2, dW1, dW3 and the bias sums
        # algorithmic HBM bytes per update: H1, H2 written + read, dH1 written + read (FP32), 64-byte row records written + read twice,
        # gathered observations, target outputs, per-CTA gradient partials
        inter = [rows * (2 * 512 + 64) + rows * D_in * 4 + rows * A_out * 4, rows * (512 + 64), rows * (3 * 512 + 64 + D_in * 4) + n_sm * 4 * (model.n_params // 2)]
        us = [1e3 * m / kernel_n for m in kernel_ms]
        k = max(range(3), key=lambda i: us[i])
        achieved = flops[k] / (us[k] * 1e-6) / 1e12
        roofline = {"bound": "tensor", "kernel": names[k], "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s", "frac": achieved / tensor_peak,
                    "traffic": TRAFFIC_NCU.get(names[k]),   # dram bytes of one launch, ncu --set full (profiles/r1_tc_pipeline.md)
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (dense bf16; TF32 runs at half of it and 3xTF32 needs three MMAs per "
                                   "FP32-accurate product: the FP32-equivalent ceiling of this arithmetic is peak / 6)",
                    "fp32_equivalent_peak": tensor_peak / 6, "frac_of_fp32_equivalent_peak": achieved / (tensor_peak / 6),
                    "launch_us": us[k], "launches_timed": kernel_n, "flop_per_launch": flops[k],
                    "kernels": {names[i]: {"launch_us": us[i], "flop": flops[i], "tflops": flops[i] / (us[i] * 1e-6) / 1e12,
                                           "algorithmic_bytes": inter[i], "gbs": inter[i] / (us[i] * 1e-6) / 1e9, "hbm_frac": inter[i] / (us[i] * 1e-6) / 1e9 / hbm_peak}
                                for i in range(3)},
                    "training_pass": {"launch_us": 1e6 * train_avg_s, "flop": train_flop, "tflops": train_flop / train_avg_s / 1e12,
                                      "frac_of_fp32_cuda_core_peak": train_flop / train_avg_s / 1e12 / fp32_peak, "fp32_cuda_core_peak": fp32_peak,
                                      "note": "events between the three kernels serialise them; the headline run overlaps their heads and tails (PDL)"},
                    "hbm_peak_gbs": hbm_peak, "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"}
    else:
        train_bytes = B * 3421 + 7 * 4 * model.n_params       # gathered episodes + parameter / Adam traffic
        achieved = train_flop / train_avg_s / 1e12 if train_n else None
        roofline = {"bound": "fp32-fma", "kernel": "train_kernel<16, kHeadDqn>", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": (achieved / fp32_peak) if achieved else None,
                    "traffic": 5524224,  # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r1_train_kernel_final.md)
                    "peak_source": f"{n_sm} SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz median SM clock sampled during the run (MEASURED_PEAKS.json holds no FP32 figure)",
                    "launch_us": 1e6 * train_avg_s, "launches_timed": train_n, "flop_per_launch": train_flop,
                    "hbm": {"achieved_gbs": (train_bytes / train_avg_s / 1e9) if train_n else None, "peak_gbs": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                            "frac": (train_bytes / train_avg_s / 1e9 / hbm_peak) if train_n else None, "bytes_per_launch": train_bytes,
                            "note": "the fused learner is compute-bound (~1 750 FLOP/B, SURVEY F7): the HBM fraction is small by construction"},
                    "bf16_tensor_peak_tflops": peaks.get("bf16_tflops_sustained")}
    # reset; per env step: forward, env; per update: target forward, 3-kernel pass | fused FP32 kernel, reduce + Adam (which also draws the next
    # update's replay indices); one sample kernel per iteration
    launches_per_step = 1 + 2 * T + U * (5 if kernel_n else 3) + 1
    line = {"metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "updates_per_sec": U * args.steps * world / (ms / 1e3), "env_steps_timed": n_steps}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """The reference loop (oracle port) on ONE core -- the reference pins torch to one thread (run.py:29) -- bounded sample."""
    from oracle import cpu_loop
    from oracle.lbf_ref import LBFConfig

    loop = cpu_loop.CpuIdqn(LBFConfig(**LBF_KW), args.batch, seed=args.seed)
    loop.prefill(args.batch)
    loop.run(1)
    n = 30
    steps, secs = loop.run(n)
    return {"value": steps / secs, "unit": "env-steps/s", "cores": 1, "kind": "port",
            "sample": f"{n} iterations of the reference loop (one 25-step episode with one env + one IDQN update at batch_size={args.batch}), 1 thread, "
                      f"pure-Python LBF restatement + PyTorch-CPU learner, {secs:.1f} s"}


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library prints) was redirected to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # stray prints of other libraries must not pollute the JSON contract
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
