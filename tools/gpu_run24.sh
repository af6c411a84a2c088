set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "hubness or draws or mirror" > gpurun_out/pytest_hub_r02u.log 2>&1; tail -6 gpurun_out/pytest_hub_r02u.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 --no-e2e > gpurun_out/bench_r02u_hub.json 2> gpurun_out/bench_r02u_hub.err; cut -c1-230 gpurun_out/bench_r02u_hub.json; tail -3 gpurun_out/bench_r02u_hub.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 --flags 64 --no-e2e > gpurun_out/bench_r02u_hub_node.json 2> gpurun_out/bench_r02u_hub_node.err; cut -c1-230 gpurun_out/bench_r02u_hub_node.json
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02u_uni.json 2> gpurun_out/bench_r02u_uni.err; cut -c1-230 gpurun_out/bench_r02u_uni.json
timeout 600 python tools/gpu_fidelity_probe.py c3s_hub 3 0:0 > gpurun_out/probe_async11_c3s_hub.log 2>&1; cat gpurun_out/probe_async11_c3s_hub.log | cut -c1-400
