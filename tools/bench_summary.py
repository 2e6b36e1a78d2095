#!/usr/bin/env python
"""Prints a compact table of a bench.py JSON line (headline, e2e, every configs[] entry with its clock record)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e, open(f).read()[:300])
        continue
    print(f"== {f}  N={d['n_gpus']}")
    c = d["clocks"] or {}
    print(f" value {d['value']:.4e}  {d['ms_per_step']:.5f} ms/step  frac {d['roofline']['frac']:.4f}  sm {c.get('sm_mhz')} {c.get('reasons')}")
    e = d["e2e"]
    r = e.get("roofline", {})
    print(f" e2e {e['value']:.3e}  d2h {r.get('achieved', 0):.1f} of {r.get('peak', 0):.1f} GB/s per GPU = {r.get('frac', 0):.3f}")
    if "host_link" in e:
        print("  link", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in e["host_link"].items() if k != "how"})
    print(" native_nccl_allreduce:", d.get("native_nccl_allreduce"), " launches", d["gpu_launches"])
    for k, c in d.get("configs", {}).items():
        if "error" in c:
            print(f"  {k:38s} ERROR {c['error']}")
            continue
        ck = c.get("clocks") or {}
        if k == "mixed_suite":
            print(f"  mixed_suite {c['suite_env_steps_per_s']:.3e}  {c['ms_per_sweep']:.3f} ms/sweep  allreduce {c['allreduce_ms']:.3f} ms  "
                  f"native {c['native_nccl_allreduce']}  sm {ck.get('sm_mhz')} {ck.get('reasons')}")
            continue
        print(f"  {k:38s} {c['env_steps_per_s']:.3e}  {c['ms_per_env_step_batch']:.4f} ms  {c['bytes_per_env_step_contract']:2d} B  "
              f"{c['achieved_GBps_per_gpu']:.0f} GB/s  frac {c['frac']:.3f}  sm {ck.get('sm_mhz')} {ck.get('reasons')} n={ck.get('samples')}")
    if "cpu_baseline" in d:
        print(f" cpu_baseline {d['cpu_baseline']['value']:.3e} on {d['cpu_baseline']['cores']} cores")
