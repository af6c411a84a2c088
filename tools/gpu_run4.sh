set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fidelity.py -q -s > gpurun_out/pytest_fidelity.log 2>&1
grep -E "cuda, oracle|passed|failed|Error" gpurun_out/pytest_fidelity.log | cut -c1-600
timeout 600 python tools/gpu_schedule_probe.py c3s 3 17x26,100x14 17x30,100x10 17x20,34x10,134x10 34x30,134x10 17x32,134x8 > gpurun_out/probe_sched2.log 2>&1
cat gpurun_out/probe_sched2.log
timeout 600 python tools/gpu_schedule_probe.py c1 3 67x30 17x20,134x10 17x20,67x10 34x30 17x30 > gpurun_out/probe_sched_c1.log 2>&1
cat gpurun_out/probe_sched_c1.log
timeout 600 python tools/gpu_schedule_probe.py c2 3 67x25 17x17,134x8 17x17,67x8 34x25 17x25 > gpurun_out/probe_sched_c2.log 2>&1
cat gpurun_out/probe_sched_c2.log
