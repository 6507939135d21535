rm -f gpurun_out/r02_j_tune.txt
python tools/gpu_probe.py tune step 8000000 '{}' '{"steps_per_event_phase":8}' '{"steps_per_event_phase":8,"min_running":12}' '{"min_running":8}' '{"steps_per_event_phase":32}' >> gpurun_out/r02_j_tune.txt 2>&1
python tools/gpu_probe.py tune radar 4000000 '{}' '{"steps_per_event_phase":8}' '{"steps_per_event_phase":8,"min_running":12}' '{"min_running":8}' '{"resident_blocks":5}' >> gpurun_out/r02_j_tune.txt 2>&1
python tools/gpu_probe.py tune planeparallel 16000000 '{}' '{"steps_per_event_phase":8}' '{"min_running":8}' >> gpurun_out/r02_j_tune.txt 2>&1
cat gpurun_out/r02_j_tune.txt
