set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "hubness or uniform or draws or async_sweeps or thinned" > gpurun_out/pytest_r02za.log 2>&1; tail -8 gpurun_out/pytest_r02za.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 > gpurun_out/bench_r02za_hub.json 2> gpurun_out/bench_r02za_hub.err; cut -c1-200 gpurun_out/bench_r02za_hub.json; tail -3 gpurun_out/bench_r02za_hub.err
timeout 600 python tools/gpu_fidelity_probe.py c3s_hub 3 0:0 > gpurun_out/probe_line3_c3s_hub.log 2>&1; cat gpurun_out/probe_line3_c3s_hub.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --hubness 1 --flags 128 > gpurun_out/bench_r02za_hub_cp.json 2> gpurun_out/bench_r02za_hub_cp.err; cut -c1-200 gpurun_out/bench_r02za_hub_cp.json
