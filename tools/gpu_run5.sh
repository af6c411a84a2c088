set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cells.py tests/test_gpu_parity.py -q > gpurun_out/pytest_cells.log 2>&1
tail -15 gpurun_out/pytest_cells.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err
cut -c1-300 gpurun_out/bench_r02d.json; grep -o '"roofline.*breakdown_ms_per_step[^}]*}' gpurun_out/bench_r02d.json; tail -5 gpurun_out/bench_r02d.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 67 > gpurun_out/bench_r02d_m67.json 2> gpurun_out/bench_r02d_m67.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02d_m67.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02d_m67.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 134 > gpurun_out/bench_r02d_m134.json 2> gpurun_out/bench_r02d_m134.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02d_m134.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02d_m134.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 17 > gpurun_out/bench_r02d_m17.json 2> gpurun_out/bench_r02d_m17.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02d_m17.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02d_m17.json
