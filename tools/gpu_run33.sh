set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu_r02zd.log 2>&1; tail -4 gpurun_out/pytest_gpu_r02zd.log
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --hubness 1 > gpurun_out/bench_r02zd_hub.json 2> gpurun_out/bench_r02zd_hub.err; cut -c1-200 gpurun_out/bench_r02zd_hub.json; tail -2 gpurun_out/bench_r02zd_hub.err
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02zd.json 2> gpurun_out/bench_r02zd.err; cut -c1-200 gpurun_out/bench_r02zd.json; tail -2 gpurun_out/bench_r02zd.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02zd.log 2>&1; tail -2 gpurun_out/smoke_r02zd.log
