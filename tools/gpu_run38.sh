set -x
mkdir -p gpurun_out
free -g | head -2; df -h /dev/shm | tail -1
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --config c5 --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r02zi_c5_4gpu.json 2> gpurun_out/bench_r02zi_c5_4gpu.err; cut -c1-300 gpurun_out/bench_r02zi_c5_4gpu.json; tail -5 gpurun_out/bench_r02zi_c5_4gpu.err; rm -f /dev/shm/annembed_bench_*
