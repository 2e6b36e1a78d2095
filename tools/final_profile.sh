#!/bin/bash
# Round-end evidence run on ONE B200 (under gpurun): bench lines, suite, ncu launch list of the bench command and
# one `ncu --set full` capture per hot kernel.  Everything lands in gpurun_out/; copy what is to be judged to profiles/.
set -u
R=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/bench_${R}_n1.json 2> gpurun_out/bench_err.log; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_${R}_reference.json 2>> gpurun_out/bench_err.log; echo "reference rc=$?"
python tools/bench_suite.py --modes step,rollout,rollout_policy,manual > gpurun_out/suite_$R.jsonl 2>&1
python tools/small_batch_latency.py > gpurun_out/small_batch_$R.md 2>&1
# launch list of the bench command (a number printed under ncu is never a bench value)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches.log 2>&1
tools/ncu_export.sh prof_step_tma_cartpole_$R step_kernel_tma 3 -- python tools/bench_suite.py --kinds 0 --modes step --reps 3 > /dev/null
tools/ncu_export.sh prof_rollout_cartpole_$R rollout_kernel 3 -- python tools/bench_suite.py --kinds 0 --modes rollout --reps 3 > /dev/null
tools/ncu_export.sh prof_step_tma_mountaincar_$R step_kernel_tma 3 -- python tools/bench_suite.py --kinds 1 --modes step --reps 3 > /dev/null
tools/ncu_export.sh prof_rollout_mountaincar_$R rollout_kernel 3 -- python tools/bench_suite.py --kinds 1 --modes rollout --reps 3 > /dev/null
tools/ncu_export.sh prof_step_tma_pendulum_$R step_kernel_tma 3 -- python tools/bench_suite.py --kinds 3 --modes step --reps 3 > /dev/null
tools/ncu_export.sh prof_step_tma_acrobot_$R step_kernel_tma 3 -- python tools/bench_suite.py --kinds 4 --modes step --reps 3 > /dev/null
ls gpurun_out | head -50
