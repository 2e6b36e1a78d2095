#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): ncu launch list of the bench command and one `ncu --set full`
# capture per hot kernel.  Everything lands in gpurun_out/; tools/ncu_summary.py condenses it into profiles/.
set -u
R=${1:-r02}
mkdir -p gpurun_out
# launch list of the bench command (a number printed under ncu is never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$R.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 --no-configs > gpurun_out/ncu_launches_$R.log 2>&1
S="python tools/bench_suite.py --reps 3"
tools/ncu_export.sh prof_step_tma_cartpole_$R step_kernel_tma 3 -- $S --kinds 0 --modes step > /dev/null
tools/ncu_export.sh prof_rollout_cartpole_$R rollout_kernel 3 -- $S --kinds 0 --modes rollout > /dev/null
tools/ncu_export.sh prof_step_tma_mountaincar_$R step_kernel_tma 3 -- $S --kinds 1 --modes step > /dev/null
tools/ncu_export.sh prof_rollout_mountaincar_$R rollout_kernel 3 -- $S --kinds 1 --modes rollout > /dev/null
tools/ncu_export.sh prof_rollout_mountaincarcont_$R rollout_kernel 3 -- $S --kinds 2 --modes rollout > /dev/null
tools/ncu_export.sh prof_step_tma_pendulum_$R step_kernel_tma 3 -- $S --kinds 3 --modes step > /dev/null
tools/ncu_export.sh prof_rollout_pendulum_$R rollout_kernel 3 -- $S --kinds 3 --modes rollout > /dev/null
tools/ncu_export.sh prof_step_tma_acrobot_$R step_kernel_tma 3 -- $S --kinds 4 --modes step > /dev/null
ls gpurun_out | grep $R | head -60
python tools/bench_suite.py --modes step,rollout,rollout_policy,manual > gpurun_out/suite_$R.jsonl 2> gpurun_out/suite_$R.err
python tools/small_batch_latency.py > gpurun_out/small_batch_$R.md 2>&1
