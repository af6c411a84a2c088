set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cells.py tests/test_gpu_parity.py -q -x > gpurun_out/pytest_cells.log 2>&1
tail -25 gpurun_out/pytest_cells.log
timeout 600 python tools/gpu_fidelity_probe.py c3s 3 0:0:0 0:67:0 0:67:1 0:67:4 0:67:67 16:67:0 > gpurun_out/probe_c3s_cells.log 2>&1
cat gpurun_out/probe_c3s_cells.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
cat gpurun_out/bench_r02b.json; tail -5 gpurun_out/bench_r02b.err
