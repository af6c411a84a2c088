set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fidelity.py -q -x -k "async or thinned or fidelity or layout or mirror" > gpurun_out/pytest_r02ze.log 2>&1; tail -3 gpurun_out/pytest_r02ze.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02ze.json 2> gpurun_out/bench_r02ze.err; cut -c1-200 gpurun_out/bench_r02ze.json; tail -2 gpurun_out/bench_r02ze.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --hubness 1 > gpurun_out/bench_r02ze_hub.json 2> gpurun_out/bench_r02ze_hub.err; cut -c1-200 gpurun_out/bench_r02ze_hub.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --dim 15 > gpurun_out/bench_r02ze_d15.json 2> gpurun_out/bench_r02ze_d15.err; cut -c1-200 gpurun_out/bench_r02ze_d15.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --flags 128 > gpurun_out/bench_r02ze_cp.json 2> gpurun_out/bench_r02ze_cp.err; cut -c1-200 gpurun_out/bench_r02ze_cp.json
