#!/bin/bash
# One parameterised profiling recipe for the GPU box (replaces the per-call scratch scripts of round 1).
#   tools/profile.sh launches <tag> [bench args]          launch list (gpu__time_duration) of a bench.py run
#   tools/profile.sh kernel <tag> <kernel regex> <skip> <count> [bench args]
#                                                         ncu --set full of <count> launches of the kernel after <skip> ones
# Reports land in gpurun_out/ (scratch); summarise with tools/ncu_summary.py / tools/agg_launches.py into profiles/.
set -e
mode=$1; tag=$2; shift 2
case $mode in
launches)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --frames-in-flight 1 "$@" > gpurun_out/launches_$tag.log 2>&1 || tail -3 gpurun_out/launches_$tag.log
  ;;
kernel)
  regex=$1; skip=$2; count=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -f -o gpurun_out/ncu_$tag \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs --frames-in-flight 1 "$@" > gpurun_out/ncu_$tag.log 2>&1 || tail -3 gpurun_out/ncu_$tag.log
  python tools/ncu_summary.py gpurun_out/ncu_$tag.ncu-rep > gpurun_out/ncu_$tag.txt 2>/dev/null || true
  ;;
esac
