set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q --deselect "tests/test_gpu_fidelity.py::test_layout_statistics_within_one_percent_of_the_reference_loop[c3s-True]" --deselect "tests/test_gpu_fidelity.py::test_layout_statistics_within_one_percent_of_the_reference_loop[c4s-False]" -s > gpurun_out/pytest_gpu_r02a.log 2>&1
tail -30 gpurun_out/pytest_gpu_r02a.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
cat gpurun_out/bench_r02a.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|DeviceRadixSort|Onesweep' -c 700 --csv --log-file gpurun_out/launches_r02a.csv python bench.py --steps 1 --warmup 0 --batches 4 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_r02a.log 2>&1
