set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "async_sweeps" > gpurun_out/pytest_r02zc.log 2>&1; tail -3 gpurun_out/pytest_r02zc.log
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --flags 128 > gpurun_out/bench_r02zc_cp.json 2> gpurun_out/bench_r02zc_cp.err; cut -c1-200 gpurun_out/bench_r02zc_cp.json; tail -2 gpurun_out/bench_r02zc_cp.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --flags 128 --hubness 1 > gpurun_out/bench_r02zc_cp_hub.json 2> gpurun_out/bench_r02zc_cp_hub.err; cut -c1-200 gpurun_out/bench_r02zc_cp_hub.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02zc_reg.json 2> gpurun_out/bench_r02zc_reg.err; cut -c1-200 gpurun_out/bench_r02zc_reg.json
