#!/bin/bash
# ncu --set full captures of the per-step kernels of the workloads named on the command line
# (each preceded by the same command without ncu, as the profiling recipe requires)
set -u
mkdir -p gpurun_out
for w in "$@"; do
  case $w in
    spatial) pat='regex:spatial_(prepare|interact)'; skip=24;;
    air) pat='regex:air_(index|step|bits|candidates)'; skip=24;;
    traders) pat='regex:traders_step'; skip=12;;
    pandemic) pat='regex:pandemic_(move|interact)'; skip=24;;
    forest) pat='regex:forest_step'; skip=12;;
    coin) pat='regex:coin_'; skip=24;;
  esac
  python tools/prof_workload.py $w > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$pat" -s $skip -c 2 -f -o gpurun_out/r01_$w python tools/prof_workload.py $w > gpurun_out/ncu_$w.log 2>&1
  echo "$w rc=$?"
done
