set -x
mkdir -p gpurun_out
TAG=${1:-r02_sweep}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 20 -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --steps 1 --warmup 0 --batches 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_${TAG}.log
ls -la gpurun_out/${TAG}*
