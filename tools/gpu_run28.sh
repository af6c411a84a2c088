set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=memory.total --format=csv,noheader; free -g | head -2
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "hubness or uniform or draws or async_sweeps" > gpurun_out/pytest_r02y.log 2>&1; tail -3 gpurun_out/pytest_r02y.log
timeout 1500 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02y_c5_1gpu.json 2> gpurun_out/bench_r02y_c5_1gpu.err; cut -c1-400 gpurun_out/bench_r02y_c5_1gpu.json; tail -5 gpurun_out/bench_r02y_c5_1gpu.err
