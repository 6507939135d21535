timeout 300 python tools/gpu_probe.py tune les 2000000 '{}' '{"pool_shape":6}' '{"pool_shape":6,"birth_min":24}' '{}' '{"pool_shape":6}' > gpurun_out/r02_av_tune.txt 2>&1
cat gpurun_out/r02_av_tune.txt
