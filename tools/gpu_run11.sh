set -x
mkdir -p gpurun_out
for c in c1 c2 c3s c3s_hub c4s; do
timeout 600 python tools/gpu_fidelity_probe.py $c 5 0:0 > gpurun_out/probe_async5_$c.log 2>&1; cat gpurun_out/probe_async5_$c.log | cut -c1-400
done
for W in w64 w64p; do
for c in c4s c3s c1; do
ANNEMBED_CUDA_LIB=$PWD/annembed_b200/libannembed_cuda_$W.so timeout 600 python tools/gpu_fidelity_probe.py $c 3 0:0 > gpurun_out/probe_async5_${c}_$W.log 2>&1; cat gpurun_out/probe_async5_${c}_$W.log | cut -c1-400
done
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
cut -c1-200 gpurun_out/bench_r02i.json; grep -o '"roofline.*breakdown_ms_per_step[^}]*}' gpurun_out/bench_r02i.json; tail -5 gpurun_out/bench_r02i.err
