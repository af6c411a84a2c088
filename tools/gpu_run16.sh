set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or thinned or relabel or abi or mirror" > gpurun_out/pytest_async_r02n.log 2>&1; tail -8 gpurun_out/pytest_async_r02n.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err
cut -c1-200 gpurun_out/bench_r02n.json; grep -o '"roofline.*' gpurun_out/bench_r02n.json | cut -c1-1200; tail -5 gpurun_out/bench_r02n.err
ANNEMBED_CUDA_LIB=$PWD/annembed_b200/libannembed_cuda_m6.so timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02n_m6.json 2> gpurun_out/bench_r02n_m6.err
grep -o '"ms_per_step[^,]*' gpurun_out/bench_r02n_m6.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_r02n_m6.json;  grep -o '"ms_steps[^]]*' gpurun_out/bench_r02n_m6.json
for c in c3s c4s c1; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_async7_$c.log 2>&1; cat gpurun_out/probe_async7_$c.log | cut -c1-400
done
