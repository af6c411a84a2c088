set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "async or thinned or relabel or mirror" > gpurun_out/pytest_async_r02s.log 2>&1; tail -4 gpurun_out/pytest_async_r02s.log
for i in 1 2; do
for V in head new; do
L=$PWD/annembed_b200/libannembed_cuda_head.so; [ $V = new ] && L=$PWD/annembed_b200/libannembed_cuda.so
ANNEMBED_CUDA_LIB=$L timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab3_${V}_$i.json 2> gpurun_out/bench_ab3_${V}_$i.err
echo $V $i; grep -o '"ms_per_step[^,]*' gpurun_out/bench_ab3_${V}_$i.json; grep -o '"avg_launch_ms[^,]*' gpurun_out/bench_ab3_${V}_$i.json
done
done
timeout 600 python tools/gpu_fidelity_probe.py c3s 3 0:0 > gpurun_out/probe_async9_c3s.log 2>&1; cat gpurun_out/probe_async9_c3s.log | cut -c1-400
TAG=r02_sweep_events_v3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 5 -c 1 -f -o gpurun_out/${TAG} python bench.py --steps 1 --warmup 0 --batches 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_${TAG}.log 2>&1
ls -la gpurun_out/${TAG}*
