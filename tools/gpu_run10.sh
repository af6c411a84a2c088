set -x
mkdir -p gpurun_out
for W in w64 w4; do
ANNEMBED_CUDA_LIB=$PWD/annembed_b200/libannembed_cuda_$W.so timeout 600 python tools/gpu_fidelity_probe.py c4s 3 0:0 > gpurun_out/probe_async4_c4s_$W.log 2>&1; cat gpurun_out/probe_async4_c4s_$W.log | cut -c1-400
ANNEMBED_CUDA_LIB=$PWD/annembed_b200/libannembed_cuda_$W.so timeout 600 python tools/gpu_fidelity_probe.py c1 5 0:0 > gpurun_out/probe_async4_c1_$W.log 2>&1; cat gpurun_out/probe_async4_c1_$W.log | cut -c1-400
done
timeout 600 python tools/gpu_fidelity_probe.py c4s 3 4:0 0:120 0:30 > gpurun_out/probe_async4_c4s.log 2>&1; cat gpurun_out/probe_async4_c4s.log | cut -c1-400
timeout 600 python tools/gpu_fidelity_probe.py c3s_hub 3 0:0 > gpurun_out/probe_async4_c3s_hub.log 2>&1; cat gpurun_out/probe_async4_c3s_hub.log | cut -c1-400
bash tools/gpu_ncu_sweep.sh r02_sweep_events_v1
