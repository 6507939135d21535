python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_k_tests.log 2>&1; tail -25 gpurun_out/r02_k_tests.log
python tools/gpu_probe.py workloads > gpurun_out/r02_k_workloads.txt 2>&1; cat gpurun_out/r02_k_workloads.txt
