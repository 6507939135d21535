timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_t_tests.log 2>&1; tail -6 gpurun_out/r02_t_tests.log
rm -f gpurun_out/r02_t_rep.txt
timeout 300 python tools/gpu_probe.py tune step 8000000 '{}' '{"replica_columns":0}' >> gpurun_out/r02_t_rep.txt 2>&1
timeout 300 python tools/gpu_probe.py tune radar 4000000 '{}' '{"replica_columns":0}' >> gpurun_out/r02_t_rep.txt 2>&1
timeout 300 python tools/gpu_probe.py tune les 2000000 '{}' '{"pool_shape":4}' >> gpurun_out/r02_t_rep.txt 2>&1
cat gpurun_out/r02_t_rep.txt
