rm -f gpurun_out/r02_am_tune.txt
timeout 300 python tools/gpu_probe.py tune les 2000000 '{}' '{"resident_blocks":6}' '{"pool_shape":4}' '{"min_running":12}' '{"min_running":20}' '{"steps_per_event_phase":8}' '{"steps_per_event_phase":32}' '{"birth_min":8}' '{"birth_min":24}' >> gpurun_out/r02_am_tune.txt 2>&1
timeout 300 python tools/gpu_probe.py tune landsat 16000000 '{}' '{"resident_blocks":5}' '{"steps_per_event_phase":8}' '{"steps_per_event_phase":32}' '{"min_running":18}' '{"min_running":20}' '{"min_running":20,"steps_per_event_phase":32}' >> gpurun_out/r02_am_tune.txt 2>&1
timeout 300 python tools/gpu_probe.py tune planeparallel 16000000 '{}' '{"resident_blocks":4}' '{"birth_min":1}' '{"birth_min":32}' '{"event_threshold":16,"birth_low":16}' >> gpurun_out/r02_am_tune.txt 2>&1
cat gpurun_out/r02_am_tune.txt
