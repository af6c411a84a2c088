set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi_r02q.log 2>&1; tail -5 gpurun_out/pytest_multi_r02q.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_r02q_2gpu.json 2> gpurun_out/bench_r02q_2gpu.err; cat gpurun_out/bench_r02q_2gpu.json | cut -c1-2400; tail -3 gpurun_out/bench_r02q_2gpu.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02q.json 2> gpurun_out/bench_r02q.err
cut -c1-200 gpurun_out/bench_r02q.json; grep -o '"roofline.*' gpurun_out/bench_r02q.json | cut -c1-1200; tail -5 gpurun_out/bench_r02q.err
for c in c3s c1; do
timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_async8_$c.log 2>&1; cat gpurun_out/probe_async8_$c.log | cut -c1-400
done
