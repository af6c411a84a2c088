set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi_r02m.log 2>&1; tail -30 gpurun_out/pytest_multi_r02m.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_r02m_2gpu.json 2> gpurun_out/bench_r02m_2gpu.err; cat gpurun_out/bench_r02m_2gpu.json | cut -c1-3000; tail -5 gpurun_out/bench_r02m_2gpu.err
