set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/pytest_multi_r02zf.log 2>&1; tail -4 gpurun_out/pytest_multi_r02zf.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r02zf_2gpu.json 2> gpurun_out/bench_r02zf_2gpu.err; cut -c1-200 gpurun_out/bench_r02zf_2gpu.json; tail -3 gpurun_out/bench_r02zf_2gpu.err
