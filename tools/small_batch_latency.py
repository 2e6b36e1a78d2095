#!/usr/bin/env python
"""Small-batch latency of one [sample_actions, step] pair: eager calls through Python/ctypes against a CUDA graph
of 16 such pairs replayed (handles created with graph_capturable=True keep the step index on the device).

    python tools/small_batch_latency.py [--kind 0]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import modurl_gym_b200 as m  # noqa: E402


def timed(fn, reps):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", type=int, default=0)
    args = ap.parse_args()
    m.load_library()
    print("| envs | eager, default handle | eager, capturable handle | graph of 16 x (sample + step), per step |")
    print("|---|---|---|---|")
    for n in (1024, 16384, 262144, 1 << 20):
        row = []
        for cap in (False, True):
            env = m.GpuVecEnv(args.kind, n, seed=1, graph_capturable=cap)
            env.reset()
            acts = torch.empty(n, dtype=env.action_dtype, device="cuda")
            env.step(env.sample_actions(out=acts))
            torch.cuda.synchronize()
            row.append(timed(lambda: env.step(env.sample_actions(out=acts)), 2000))
            if cap:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(16):
                        env.step(env.sample_actions(out=acts))
                row.append(timed(g.replay, 300) / 16)
            env.close()
        print(f"| {n} | {row[0]:.2f} us | {row[1]:.2f} us | {row[2]:.2f} us |")


if __name__ == "__main__":
    main()
