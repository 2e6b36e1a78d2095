#!/bin/bash
# usage: tools/ncu_export.sh <name> <kernel-regex> <skip> -- <command...>
# Runs <command> under `ncu --set full` for ONE launch of the matching kernel, exports the raw and source
# pages as CSV into gpurun_out/ and deletes the (25 MB) report so that gpurun_out stays under its size cap.
set -u
name=$1; regex=$2; skip=$3; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page source --csv --print-source sass > gpurun_out/$name.source.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page details > gpurun_out/$name.details.txt 2>/dev/null
[ "${KEEP_REP:-0}" = "1" ] || rm -f gpurun_out/$name.ncu-rep
ls -la gpurun_out/$name.*
