set -x
mkdir -p gpurun_out
for i in 1 2; do
for V in head ballot m4; do
L=$PWD/annembed_b200/libannembed_cuda_$V.so; [ $V = ballot ] && L=$PWD/annembed_b200/libannembed_cuda.so
ANNEMBED_CUDA_LIB=$L timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab2_${V}_$i.json 2> gpurun_out/bench_ab2_${V}_$i.err
echo $V $i; grep -o '"ms_per_step[^,]*' gpurun_out/bench_ab2_${V}_$i.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_ab2_${V}_$i.json
done
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or thinned" > gpurun_out/pytest_async_r02o.log 2>&1; tail -4 gpurun_out/pytest_async_r02o.log
