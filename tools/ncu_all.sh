#!/bin/bash
# ncu --set full captures of the per-step kernels of the workloads named on the command line
# (each preceded by the same command without ncu, as the profiling recipe requires).
# "forest4096" = the bench.py N=1 configuration; "launches" = the launch list of `bench.py --steps 2 --warmup 3`.
set -u
mkdir -p gpurun_out
for w in "$@"; do
  extra=""
  name=$w
  case $w in
    spatial) pat='regex:spatial_(prepare|interact)'; skip=24;;
    air) pat='regex:air_(candidates|step)'; skip=24;;
    traders) pat='regex:traders_step'; skip=12;;
    pandemic) pat='regex:pandemic_(move|interact)'; skip=24;;
    forest) pat='regex:forest_step'; skip=12;;
    forest4096) pat='regex:forest_step'; skip=12; name=forest; extra="--scale 0.0625";;
    coin) pat='regex:coin_'; skip=24;;
    forestbench)
      # forest_step exactly as bench.py times it: every launch preceded by the (untimed) L2 flush
      python bench.py --steps 2 --warmup 3 --only-headline > gpurun_out/plain_forestbench.log 2>&1 &&
      ncu --set full --clock-control none --import-source on -k regex:forest_step -s 3 -c 2 -f -o gpurun_out/r02_forestbench \
          python bench.py --steps 2 --warmup 3 --only-headline > gpurun_out/ncu_forestbench.log 2>&1
      echo "forestbench rc=$?"; continue;;
    launches)
      python bench.py --steps 2 --warmup 3 --only-headline > gpurun_out/plain_bench.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
          python bench.py --steps 2 --warmup 3 --only-headline > gpurun_out/ncu_bench.log 2>&1
      echo "launches rc=$?"; continue;;
  esac
  python tools/prof_workload.py $name $extra > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$pat" -s $skip -c 2 -f -o gpurun_out/r02_$w python tools/prof_workload.py $name $extra > gpurun_out/ncu_$w.log 2>&1
  echo "$w rc=$?"
done
