set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cells.py -q -x > gpurun_out/pytest_parity_r02h.log 2>&1; tail -25 gpurun_out/pytest_parity_r02h.log
timeout 600 python tools/gpu_fidelity_probe.py c3s 3 0:0 > gpurun_out/probe_async3_c3s.log 2>&1; cat gpurun_out/probe_async3_c3s.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c1 5 0:0 > gpurun_out/probe_async3_c1.log 2>&1; cat gpurun_out/probe_async3_c1.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c2 5 0:0 > gpurun_out/probe_async3_c2.log 2>&1; cat gpurun_out/probe_async3_c2.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c4s 3 0:0 > gpurun_out/probe_async3_c4s.log 2>&1; cat gpurun_out/probe_async3_c4s.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err
cut -c1-200 gpurun_out/bench_r02h.json; grep -o '"roofline.*breakdown_ms_per_step[^}]*}' gpurun_out/bench_r02h.json; tail -5 gpurun_out/bench_r02h.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --flags 4 > gpurun_out/bench_r02h_f4.json 2> gpurun_out/bench_r02h_f4.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02h_f4.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02h_f4.json
