set -x
mkdir -p gpurun_out
timeout 600 python tools/gpu_fidelity_probe.py c3s 3 4:0 4:60 4:15 > gpurun_out/probe_async2_c3s.log 2>&1; cat gpurun_out/probe_async2_c3s.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c1 5 4:0 4:100 4:25 > gpurun_out/probe_async2_c1.log 2>&1; cat gpurun_out/probe_async2_c1.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c4s 3 4:0 > gpurun_out/probe_async2_c4s.log 2>&1; cat gpurun_out/probe_async2_c4s.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --flags 4 > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err
cut -c1-200 gpurun_out/bench_r02g.json; grep -o '"roofline.*breakdown_ms_per_step[^}]*}' gpurun_out/bench_r02g.json; tail -5 gpurun_out/bench_r02g.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 15 --flags 4 > gpurun_out/bench_r02g_m15.json 2> gpurun_out/bench_r02g_m15.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02g_m15.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02g_m15.json
