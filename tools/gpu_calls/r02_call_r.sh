timeout 300 python tools/gpu_probe.py tune step 8000000 '{}' '{"resident_blocks":4}' '{"stage_tallies":0}' > gpurun_out/r02_r_tsm.txt 2>&1
timeout 300 python tools/gpu_probe.py tune planeparallel 16000000 '{}' '{"resident_blocks":4}' >> gpurun_out/r02_r_tsm.txt 2>&1
cat gpurun_out/r02_r_tsm.txt
