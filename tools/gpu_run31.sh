set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 3 -c 1 -f -o /tmp/r02_sweep_events_cp_v1 python bench.py --steps 1 --warmup 0 --batches 2 --no-e2e --no-cpu-baseline --flags 128 > gpurun_out/ncu_r02_sweep_events_cp_v1.log 2>&1
tail -3 gpurun_out/ncu_r02_sweep_events_cp_v1.log
ls -la /tmp/r02_sweep_events_cp_v1.ncu-rep
ncu -i /tmp/r02_sweep_events_cp_v1.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_sweep_events_cp_v1_source.csv 2>/dev/null
ncu -i /tmp/r02_sweep_events_cp_v1.ncu-rep --page raw --csv > gpurun_out/r02_sweep_events_cp_v1_raw.csv 2>/dev/null
ls -la gpurun_out
