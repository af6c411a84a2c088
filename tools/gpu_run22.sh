set -x
mkdir -p gpurun_out
for i in 1 2; do
for V in head trims new; do
L=$PWD/annembed_b200/libannembed_cuda_$V.so; [ $V = new ] && L=$PWD/annembed_b200/libannembed_cuda.so
ANNEMBED_CUDA_LIB=$L timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab4_${V}_$i.json 2> gpurun_out/bench_ab4_${V}_$i.err
echo $V $i; grep -o '"ms_per_step[^,]*' gpurun_out/bench_ab4_${V}_$i.json
done
done
