set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02zb.json 2> gpurun_out/bench_r02zb.err; cut -c1-200 gpurun_out/bench_r02zb.json; tail -3 gpurun_out/bench_r02zb.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --hubness 1 > gpurun_out/bench_r02zb_hub.json 2> gpurun_out/bench_r02zb_hub.err; cut -c1-200 gpurun_out/bench_r02zb_hub.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --flags 128 > gpurun_out/bench_r02zb_cp.json 2> gpurun_out/bench_r02zb_cp.err; cut -c1-200 gpurun_out/bench_r02zb_cp.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --dim 15 > gpurun_out/bench_r02zb_d15.json 2> gpurun_out/bench_r02zb_d15.err; cut -c1-200 gpurun_out/bench_r02zb_d15.json
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "async" > gpurun_out/pytest_r02zb.log 2>&1; tail -3 gpurun_out/pytest_r02zb.log
