#!/usr/bin/env python
"""bench.py -- env-steps/sec of the classic-control hot path on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: ONE mgym_step launch over 2^24 CartPole-v1
envs per GPU with device-side auto-reset (BASELINE.json configs[1]).  Envs shard over ranks as
independent contiguous slices (no data-path collective; "scaling": "weak"); the only collective is
the all-reduce of 5 statistics doubles after the timed region.

--impl reference times the CPU restatement of the reference's step loop (oracle/, a C port: the
Rust crate cannot be built in this image) on the host cores.
"""
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

METRIC = "env-steps/sec (CartPole-v1, 2^24 envs)"
UNIT = "env-steps/s"
KIND_NAME = "CartPole-v1"
NUM_ENVS_PER_GPU = 1 << 24
# Algorithmic bytes per env-step of the per-call step kernel:
# state read 16 + state/obs write 16 + action u8 1 + reward f32 4 + flags u8 1 + u16 counter 2+2
BYTES_PER_ENV_STEP = 42


def workload_name(n):
    return (f"{KIND_NAME}, {n} envs per GPU, per-call step kernel with device-side auto-reset "
            "(BASELINE.json configs[1])")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--num-envs", type=int, default=NUM_ENVS_PER_GPU, help="envs per GPU")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    return ap.parse_args()


# -------------------------------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", os.environ.get("MGYM_SMI_MS", "100")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [l for (ts, l) in self.lines if t0 - 0.05 <= ts <= t1 + 0.15] or [l for (_, l) in self.lines]
        sm, sm_max, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in inside:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                sm_max.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(sm_max) if sm_max else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, from the committed ncu capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[f"{KIND_NAME}/{n}/step"]
        return t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]
    except Exception:
        return None, None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPUs next to its GPU (NVML's ideal affinity) so that the pinned host buffers of the
    end-to-end leg are allocated on the local NUMA node; matters when 8 ranks stream over PCIe at once."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus ({allowed[0]}-{allowed[-1]})"
    except Exception as e:  # affinity is an optimisation only
        return f"unbound ({type(e).__name__})"
    return "unbound"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle's scalar step loop, one env loop per thread
# -------------------------------------------------------------------------------------------------
def cpu_step_loop(seconds, threads):
    """Runs the reference's caller loop (cartpole.rs:460-471, restated in oracle/mgym_oracle.c) for
    about `seconds` on `threads` host threads; returns (env_steps_per_sec, total_steps, wall)."""
    from oracle import oracle as o

    total, wall, _ = o.baseline_loop(o.CARTPOLE, 1_000_000, threads, seed=0x5EED)  # calibrate
    rate = total / max(wall, 1e-9)
    per_thread = max(1_000_000, int(rate * seconds / threads))
    total, wall, _ = o.baseline_loop(o.CARTPOLE, per_thread, threads, seed=0x5EED)
    return total / wall, total, wall


def run_reference(args, rank):
    """--impl reference: every 'step' is a bounded sample -- each host thread advances its own
    CartPole env by S steps with random actions and reset-on-done (the reference's loop)."""
    if rank != 0:
        return
    from oracle import oracle as o

    threads = os.cpu_count() or 1
    S = 1 << 20  # env-steps per thread per 'step' (~80 ms): K=2000 ends within about three minutes
    for _ in range(args.warmup):
        o.baseline_loop(o.CARTPOLE, S, threads, seed=0x5EED)
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        n, _, _ = o.baseline_loop(o.CARTPOLE, S, threads, seed=0x5EED)
        total += n
    wall = time.perf_counter() - t0
    value = total / wall
    sample = f"{threads} threads x {S} CartPole-v1 env-steps per step (one env loop per thread, reset on done)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.num_envs), "num_envs_per_gpu": args.num_envs, "mode": "step",
                   "arm": "CPU restatement of cartpole.rs:251-348 (oracle/ C port; the Rust crate cannot be built "
                          "here: no cargo/rustc); each step is a bounded sample of the workload", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# native arm
# -------------------------------------------------------------------------------------------------
def run_native(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import modurl_gym_b200 as m

    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    m.load_library()
    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        # NCCL prints its version banner (and, with NCCL_DEBUG=INFO, its topology log) on stdout when the
        # communicator comes up; stdout is reserved for the ONE JSON line, so fd 1 points at stderr meanwhile.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()  # eager communicator creation
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from modurl_gym_b200.distributed import max_over_ranks, shard_range

    n = args.num_envs  # weak scaling: a fixed slice per GPU
    begin, end = shard_range(n * world, rank, world)
    assert end - begin == n
    env = m.GpuVecEnv(KIND_NAME, n, device=local_rank, seed=0x5EED, env_index_base=begin)
    env.reset()
    # rotating pool of pre-generated random actions keeps RNG out of the timed kernel
    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + rank)
    pool = [torch.randint(0, 2, (n,), dtype=torch.uint8, device=device, generator=gen) for _ in range(16)]

    def one_step(i):
        # the public call: Gym::step.  For CartPole the returned observation is a view of the resident state rows
        # (obs_out = NULL in the C call), reward and flags go to the env's own buffers.
        env.step(pool[i & 15])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        one_step(i)
    env.reset_stats()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for i in range(args.steps):
        one_step(i)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_max = max_over_ranks(ms, device)
    stats = env.all_reduce_stats()  # NCCL all-reduce of the episode statistics (5 doubles)

    # ---- end to end through the host-buffer entry point: H2D actions, step, D2H obs/reward/flags ----
    h_act = torch.randint(0, 2, (n,), dtype=torch.uint8).pin_memory()
    h_obs = torch.empty((4, n), dtype=torch.float32).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_flg = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        env.step_host(h_act, h_obs, h_rew, h_flg)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        env.step_host(h_act, h_obs, h_rew, h_flg)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), device)
    h2d = h_act.numel() * h_act.element_size()
    d2h = sum(x.numel() * x.element_size() for x in (h_obs, h_rew, h_flg))

    if rank == 0:
        total_env_steps = float(n) * world * args.steps
        value = total_env_steps / (ms_max * 1e-3)
        launch_s = ms_max * 1e-3 / args.steps
        peak, peak_src = measured_peak_gbs()
        achieved = BYTES_PER_ENV_STEP * n / launch_s / 1e9
        traffic, traffic_src = ncu_traffic(n)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(n),
                "num_envs_per_gpu": n, "mode": "step", "obs": "zero-copy (obs aliases the resident state rows)",
                "actions": "rotating pool of 16 pre-generated uint8[N] buffers",
                "l2": f"working set {BYTES_PER_ENV_STEP * n / 1e6:.0f} MB per step > 126 MB L2 (inputs larger than L2)",
                "parallelism": f"dp{world} (independent env slices, no data-path collective)",
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n, "kernel": "step_kernel_tma<CartPole-v1, u16 counter>",
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "peak_source": peak_src},
            "e2e": {"value": float(n) * world * args.e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
                    "path": "mgym_step_host: pinned host actions -> device, step, obs/reward/flags -> pinned host",
                    # what bounds it: bytes over each GPU's PCIe link per second of the timed region
                    "pcie_GBps_per_gpu": (h2d + d2h) * args.e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "bound": "pcie (22 B per env-step cross the link; the kernel itself takes 2 % of the call)",
                    "host_affinity_rank0": numa},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "episode_stats": {"episodes": stats.episodes, "mean_length": stats.length_sum / max(stats.episodes, 1),
                              "mean_return": stats.return_sum / max(stats.episodes, 1)},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, total, wall = cpu_step_loop(args.cpu_seconds, threads)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"{total} CartPole-v1 env-steps in {wall:.1f} s: oracle C port of cartpole.rs:251-348, one env "
                          f"loop per thread, random actions, reset on done; an upper bound on the Rust reference, "
                          f"which also allocates 3 tensors per step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
