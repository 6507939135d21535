rm -f gpurun_out/r02_f_tune.txt
python tools/gpu_probe.py tune les 2000000 '{}' '{"resident_blocks":6}' >> gpurun_out/r02_f_tune.txt 2>&1
python tools/gpu_probe.py tune landsat 16000000 '{}' '{"min_running":12}' '{"min_running":20}' '{"min_running":24}' '{"event_threshold":8}' '{"event_threshold":24}' '{"event_threshold":32}' '{"steps_per_event_phase":8}' '{"steps_per_event_phase":32}' '{"min_running":20,"event_threshold":24}' '{"resident_blocks":5}' >> gpurun_out/r02_f_tune.txt 2>&1
python tools/gpu_probe.py tune step 4000000 '{}' '{"min_running":12}' '{"min_running":24}' '{"event_threshold":8}' '{"event_threshold":32}' '{"stage_tallies":0}' >> gpurun_out/r02_f_tune.txt 2>&1
python tools/gpu_probe.py locality >> gpurun_out/r02_f_tune.txt 2>&1
cat gpurun_out/r02_f_tune.txt
