set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or thinned or hubness or events or uniform or draws" > gpurun_out/pytest_r02w.log 2>&1; tail -5 gpurun_out/pytest_r02w.log
for fl in 0 128 256 384; do
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-e2e --flags $fl > gpurun_out/bench_r02w_f$fl.json 2> gpurun_out/bench_r02w_f$fl.err; cut -c1-160 gpurun_out/bench_r02w_f$fl.json
done
for c in c3s c1 c2 c4s; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_line1_$c.log 2>&1; cat gpurun_out/probe_line1_$c.log | cut -c1-400
done
