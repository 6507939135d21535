rm -f gpurun_out/r02_ag_tune.txt
timeout 400 python tools/gpu_probe.py tune landsat 16000000 '{}' '{"birth_low":24}' '{"birth_low":32}' '{"birth_low":8}' '{"min_running":20}' '{"min_running":22}' '{"min_running":20,"event_threshold":8}' '{"min_running":20,"event_threshold":12}' '{"resident_blocks":5,"pool_shape":5}' '{"resident_blocks":5,"pool_shape":6}' '{"resident_blocks":6,"pool_shape":5}' '{"resident_blocks":6,"pool_shape":5,"min_running":20}' '{"resident_blocks":5,"pool_shape":5,"birth_min":32}' >> gpurun_out/r02_ag_tune.txt 2>&1
timeout 200 python tools/gpu_probe.py tune les-small 2000000 '{}' '{"birth_low":32}' '{"min_running":20}' '{"resident_blocks":6,"pool_shape":5}' '{"resident_blocks":5,"pool_shape":6}' >> gpurun_out/r02_ag_tune.txt 2>&1
timeout 200 python tools/gpu_probe.py tune les 2000000 '{}' '{"birth_low":32}' '{"min_running":20}' '{"event_threshold":8}' >> gpurun_out/r02_ag_tune.txt 2>&1
timeout 200 python tools/gpu_probe.py tune step 8000000 '{}' '{"birth_low":32}' '{"min_running":20}' '{"event_threshold":8}' >> gpurun_out/r02_ag_tune.txt 2>&1
cat gpurun_out/r02_ag_tune.txt
