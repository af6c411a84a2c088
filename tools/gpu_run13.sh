set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/pytest_gpu_r02k.log 2>&1; tail -60 gpurun_out/pytest_gpu_r02k.log
