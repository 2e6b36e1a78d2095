#!/usr/bin/env python
"""Secondary measurements (not the driver's bench.py contract): every kind in both kernel modes.

Prints one JSON line per (kind, mode): env-steps/s with everything resident in HBM, algorithmic GB/s and
the fraction of the measured HBM peak.  Used for BASELINE.json configs[2] (MountainCar rollout) and
configs[3] (Acrobot / Pendulum at 2^22 envs) and for the step-vs-rollout evidence in profiles/.

    python tools/bench_suite.py [--kinds 0,1,2,3,4] [--modes step,rollout,rollout_policy] [--chunk 32] [--reps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import modurl_gym_b200 as m  # noqa: E402

import bench  # noqa: E402  (the repo-root bench.py: one table of contract bytes for both tools)

NAMES = bench.NAMES
DEFAULT_N = [1 << 24, 1 << 24, 1 << 24, 1 << 22, 1 << 22]
# SURVEY 8(d) contract bytes per env-step (what bench.py quotes fractions against)
STEP_BYTES = bench.STEP_CONTRACT
OBS_BYTES = bench.OBS_BYTES
ACT_BYTES = bench.ACT_BYTES


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kinds", default="0,1,2,3,4")
    ap.add_argument("--modes", default="step,rollout,rollout_policy")
    ap.add_argument("--chunk", type=int, default=32, help="steps per rollout launch (trajectory ring depth)")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--num-envs", type=int, default=0)
    ap.add_argument("--track-stats", type=int, default=1)
    args = ap.parse_args()
    pk = peak()
    for kind in [int(k) for k in args.kinds.split(",")]:
        n = args.num_envs or DEFAULT_N[kind]
        env = m.GpuVecEnv(kind, n, seed=0x5EED, track_stats=bool(args.track_stats))
        env.reset()
        gen = torch.Generator(device="cuda")
        gen.manual_seed(kind)
        K = args.chunk
        if env.continuous:
            lim = 1.0 if kind == 2 else 2.0
            acts = (torch.rand((K, n), device="cuda", generator=gen) * 2 - 1) * lim
        else:
            acts = torch.randint(0, env.action_space().n, (K, n), dtype=torch.uint8, device="cuda", generator=gen)
        reward = torch.empty((K, n), device="cuda")
        flags = torch.empty((K, n), dtype=torch.uint8, device="cuda")
        obs = torch.empty((K, env.obs_dim, n), device="cuda")
        zero_copy = kind in (0, 1, 2)
        for mode in args.modes.split(","):
            if mode == "manual":
                continue
            if mode == "step":
                i = [0]

                def fn():
                    k = i[0] % K
                    env.step_raw(acts[k], None if zero_copy else obs[0], reward[0], flags[0])
                    i[0] += 1

                sec = timed(fn, args.reps * 4)
                steps, bytes_per = 1, STEP_BYTES[kind]
            else:
                policy = mode == "rollout_policy"

                def fn():
                    env.rollout(K, None if policy else acts, obs=obs, reward=reward, flags=flags, count_done=True)

                sec = timed(fn, max(2, args.reps // 4))
                steps = K
                bytes_per = OBS_BYTES[kind] + 4 + 1 + (0 if policy else ACT_BYTES[kind])
            rate = n * steps / sec
            gbs = rate * bytes_per / 1e9
            s = env.stats()
            print(json.dumps({"kind": NAMES[kind], "mode": mode, "num_envs": n, "steps_per_launch": steps,
                              "ms_per_launch": sec * 1e3, "env_steps_per_s": rate, "bytes_per_env_step": bytes_per,
                              "algorithmic_GBps": gbs, "frac_of_measured_hbm_peak": gbs / pk,
                              "mean_episode_length": s.length_sum / max(s.episodes, 1)}), flush=True)
        if "manual" in args.modes.split(","):
            # the reference's own protocol, batched: Gym::step without auto-reset, then the caller resets the envs
            # that returned done or truncated (cartpole.rs:468-470) -- two launches per step
            menv = m.GpuVecEnv(kind, n, seed=0x5EED, auto_reset=False)
            menv.reset()
            i = [0]

            def fn():
                k = i[0] % K
                menv.step_raw(acts[k], None if zero_copy else obs[0], reward[0], flags[0])
                menv.reset(mask=flags[0])
                i[0] += 1

            sec = timed(fn, args.reps * 4)
            rate = n / sec
            print(json.dumps({"kind": NAMES[kind], "mode": "manual step + masked reset", "num_envs": n,
                              "steps_per_launch": 1, "ms_per_launch": sec * 1e3, "env_steps_per_s": rate}), flush=True)
            menv.close()
            del menv
        env.close()
        del env, acts, reward, flags, obs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
