timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_q_tests.log 2>&1; tail -4 gpurun_out/r02_q_tests.log
timeout 300 python tools/gpu_probe.py workloads > gpurun_out/r02_q_workloads.txt 2>&1; cat gpurun_out/r02_q_workloads.txt
timeout 300 python tools/gpu_probe.py tune radar 4000000 '{}' >> gpurun_out/r02_q_workloads.txt 2>&1
timeout 300 python tools/gpu_probe.py tune step 8000000 '{}' >> gpurun_out/r02_q_workloads.txt 2>&1
timeout 300 python tools/gpu_probe.py tune les 2000000 '{}' >> gpurun_out/r02_q_workloads.txt 2>&1
tail -3 gpurun_out/r02_q_workloads.txt
timeout 900 python bench.py > gpurun_out/r02_q_bench.json 2> gpurun_out/r02_q_bench.err; head -c 250 gpurun_out/r02_q_bench.json; echo
timeout 900 python bench.py --workload les --photons 2000000 > gpurun_out/r02_q_bench_les.json 2> gpurun_out/r02_q_bench_les.err; head -c 250 gpurun_out/r02_q_bench_les.json; echo
tools/ncu_full.sh r02_final2 les 1000000
