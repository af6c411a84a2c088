set -x
mkdir -p gpurun_out
# async sweep: fidelity on the 1M Higgs-shape case, kappa 2 / 1 / 4, then the 70k cases and d=15
timeout 600 python tools/gpu_fidelity_probe.py c3s 3 0:0 0:60 0:15 > gpurun_out/probe_async_c3s.log 2>&1; cat gpurun_out/probe_async_c3s.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c1 5 0:0 0:100 0:25 > gpurun_out/probe_async_c1.log 2>&1; cat gpurun_out/probe_async_c1.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c2 5 0:0 > gpurun_out/probe_async_c2.log 2>&1; cat gpurun_out/probe_async_c2.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c4s 3 0:0 > gpurun_out/probe_async_c4s.log 2>&1; cat gpurun_out/probe_async_c4s.log | cut -c1-400
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err
cut -c1-200 gpurun_out/bench_r02f.json; grep -o '"roofline.*breakdown_ms_per_step[^}]*}' gpurun_out/bench_r02f.json; tail -5 gpurun_out/bench_r02f.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 15 > gpurun_out/bench_r02f_m15.json 2> gpurun_out/bench_r02f_m15.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02f_m15.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02f_m15.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 60 > gpurun_out/bench_r02f_m60.json 2> gpurun_out/bench_r02f_m60.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02f_m60.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02f_m60.json
