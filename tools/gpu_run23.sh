set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "hubness or async or thinned or draws or mirror" > gpurun_out/pytest_hub_r02t.log 2>&1; tail -6 gpurun_out/pytest_hub_r02t.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 > gpurun_out/bench_r02t_hub.json 2> gpurun_out/bench_r02t_hub.err; cut -c1-230 gpurun_out/bench_r02t_hub.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/bench_r02t_hub.json; tail -3 gpurun_out/bench_r02t_hub.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --hubness 1 --flags 64 --no-e2e > gpurun_out/bench_r02t_hub_node.json 2> gpurun_out/bench_r02t_hub_node.err; cut -c1-230 gpurun_out/bench_r02t_hub_node.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --dim 15 --no-e2e > gpurun_out/bench_r02t_d15.json 2> gpurun_out/bench_r02t_d15.err; cut -c1-230 gpurun_out/bench_r02t_d15.json
for c in c3s_hub c4s; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_async10_$c.log 2>&1; cat gpurun_out/probe_async10_$c.log | cut -c1-400
done
