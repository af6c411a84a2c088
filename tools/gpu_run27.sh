set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu_r02x.log 2>&1; tail -4 gpurun_out/pytest_gpu_r02x.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02x.json 2> gpurun_out/bench_r02x.err; cut -c1-200 gpurun_out/bench_r02x.json; tail -3 gpurun_out/bench_r02x.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 > gpurun_out/bench_r02x_hub.json 2> gpurun_out/bench_r02x_hub.err; cut -c1-200 gpurun_out/bench_r02x_hub.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|DeviceRadixSort|Onesweep' -c 400 --csv --log-file gpurun_out/launches_r02x.csv python bench.py --steps 1 --warmup 0 --batches 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_r02x.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o gpurun_out/r02_sweep_events_v4 python bench.py --steps 1 --warmup 0 --batches 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r02_sweep_events_v4.log 2>&1
tail -3 gpurun_out/ncu_r02_sweep_events_v4.log
