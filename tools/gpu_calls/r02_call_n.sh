python tools/gpu_probe.py tune les 2000000 '{}' '{"l2_persist":1}' '{}' '{"l2_persist":1}' > gpurun_out/r02_n_l2persist.txt 2>&1; cat gpurun_out/r02_n_l2persist.txt
