set -x
mkdir -p gpurun_out
bash tools/gpu_ncu_sweep.sh r02_sweep_events_v2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_async.csv python bench.py --steps 1 --warmup 0 --batches 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_r02_async.log 2>&1
tail -2 gpurun_out/bench_under_ncu_r02_async.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 20 > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err; cat gpurun_out/bench_r02l.json; tail -3 gpurun_out/bench_r02l.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 > gpurun_out/bench_r02l_hub.json 2> gpurun_out/bench_r02l_hub.err; cut -c1-250 gpurun_out/bench_r02l_hub.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --dim 15 --no-e2e > gpurun_out/bench_r02l_d15.json 2> gpurun_out/bench_r02l_d15.err; cut -c1-250 gpurun_out/bench_r02l_d15.json
