set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cells.py tests/test_gpu_parity.py -q > gpurun_out/pytest_cells.log 2>&1
tail -15 gpurun_out/pytest_cells.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
cat gpurun_out/bench_r02c.json; tail -5 gpurun_out/bench_r02c.err
ANNEMBED_CUDA_LIB=$PWD/annembed_b200/libannembed_cuda_t1024.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02c_t1024.json 2> gpurun_out/bench_r02c_t1024.err
cat gpurun_out/bench_r02c_t1024.json; tail -5 gpurun_out/bench_r02c_t1024.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 67 > gpurun_out/bench_r02c_m67.json 2> gpurun_out/bench_r02c_m67.err
cat gpurun_out/bench_r02c_m67.json
