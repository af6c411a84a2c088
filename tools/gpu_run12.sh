set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or thinned" > gpurun_out/pytest_async_r02j.log 2>&1; tail -25 gpurun_out/pytest_async_r02j.log
for c in c4s c3s c1; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_async6_$c.log 2>&1; cat gpurun_out/probe_async6_$c.log | cut -c1-400
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err
cut -c1-200 gpurun_out/bench_r02j.json; grep -o '"roofline.*breakdown_ms_per_step[^}]*}' gpurun_out/bench_r02j.json; tail -5 gpurun_out/bench_r02j.err
for c in c2 c3s_hub; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_async6_$c.log 2>&1; cat gpurun_out/probe_async6_$c.log | cut -c1-400
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --mini-epochs 60 > gpurun_out/bench_r02j_m60.json 2> gpurun_out/bench_r02j_m60.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02j_m60.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02j_m60.json
