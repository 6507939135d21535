rm -f gpurun_out/r02_au_tune.txt
timeout 300 python tools/gpu_probe.py tune landsat 16000000 '{}' '{"resident_blocks":7,"pool_shape":2}' '{"resident_blocks":7,"pool_shape":2,"birth_min":8}' '{"resident_blocks":7,"pool_shape":2,"birth_min":8,"min_running":20}' '{"resident_blocks":7,"pool_shape":2,"birth_min":4,"birth_low":8,"event_threshold":8}' >> gpurun_out/r02_au_tune.txt 2>&1
cat gpurun_out/r02_au_tune.txt
