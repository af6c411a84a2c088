#!/bin/bash
# GPU-box profiling recipe (run through gpurun): launch list + ncu --set full captures of the epoch kernels at the
# coarse and at the finest mini-epoch level of the graded schedule.
# usage: bash tools_profile.sh <tag> [nodes]
set -x
TAG=${1:-r01}
NODES=${2:-11000000}
mkdir -p gpurun_out
# 4 batches: graded schedule 9 + 17 + 34 + 34 mini-epochs = 94 launches of each epoch kernel
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|DeviceRadixSort|Onesweep' -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --nodes $NODES --steps 1 --warmup 0 --batches 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# launches 10-11 of k_epoch_*: second batch (17 mini-epochs); launches 60-61: a finest-level (34) mini-epoch
ncu --set full --clock-control none --import-source on -k regex:k_epoch -s 10 -c 2 -f -o gpurun_out/k4_${TAG}_coarse \
    python bench.py --nodes $NODES --steps 1 --warmup 0 --batches 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_epoch -s 60 -c 2 -f -o gpurun_out/k4_${TAG}_fine \
    python bench.py --nodes $NODES --steps 1 --warmup 0 --batches 4 --no-e2e --no-cpu-baseline >> gpurun_out/bench_under_ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -5
