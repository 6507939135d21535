timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_ai_tests.log 2>&1; tail -8 gpurun_out/r02_ai_tests.log
timeout 900 python bench.py > gpurun_out/r02_ai_bench.json 2> gpurun_out/r02_ai_bench.err; head -c 600 gpurun_out/r02_ai_bench.json; echo
python tools/gpu_probe.py workloads > gpurun_out/r02_ai_workloads.txt 2>&1; cat gpurun_out/r02_ai_workloads.txt
