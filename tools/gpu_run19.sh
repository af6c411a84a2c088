set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench_r02p_${N}gpu.json 2> gpurun_out/bench_r02p_${N}gpu.err; cat gpurun_out/bench_r02p_${N}gpu.json | cut -c1-2600; tail -5 gpurun_out/bench_r02p_${N}gpu.err
