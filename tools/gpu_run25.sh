set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or thinned or hubness or events" > gpurun_out/pytest_cp_r02v.log 2>&1; tail -5 gpurun_out/pytest_cp_r02v.log
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02v_cp.json 2> gpurun_out/bench_r02v_cp.err; cut -c1-230 gpurun_out/bench_r02v_cp.json; tail -3 gpurun_out/bench_r02v_cp.err
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --flags 128 > gpurun_out/bench_r02v_reg.json 2> gpurun_out/bench_r02v_reg.err; cut -c1-230 gpurun_out/bench_r02v_reg.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --hubness 1 > gpurun_out/bench_r02v_cp_hub.json 2> gpurun_out/bench_r02v_cp_hub.err; cut -c1-230 gpurun_out/bench_r02v_cp_hub.json
for c in c3s c1; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_cp1_$c.log 2>&1; cat gpurun_out/probe_cp1_$c.log | cut -c1-400
done
