#!/bin/bash
# ncu evidence (run under gpurun, one GPU):  TAG=<name> KERNEL=<regex> BENCH_ARGS="..." bash profiles/gpu_profile.sh
# Each ncu pass runs only after the same command has exited 0 without ncu.  Outputs land in gpurun_out/;
# profiles/summarize_ncu.py turns them into the committed summaries.
set -u
export HJB_BENCH_NO_CLOCK_LOOP=1   # the 1 s clock-sampling loop would put hundreds of launches under ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vhjb --no-extra --no-parity --no-verify ${BENCH_ARGS:-}"
TAG=${TAG:-rollout}
KERNEL=${KERNEL:-rollout_kernel}
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL} -s ${SKIP:-3} -c 1 \
    -o gpurun_out/${TAG}_full -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
# the raw metric page travels back as CSV (small); the report itself (8-13 MB with sources) only when KEEP_REP=1 —
# gpurun merges at most 64 MiB of gpurun_out/ back
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
[ "${KEEP_REP:-0}" = "1" ] || rm -f gpurun_out/${TAG}_full.ncu-rep
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
tail -2 gpurun_out/${TAG}_ncu_full.log
