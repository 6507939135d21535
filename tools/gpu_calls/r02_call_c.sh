python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_c_tests.log 2>&1; tail -15 gpurun_out/r02_c_tests.log
rm -f gpurun_out/r02_c_jump.txt
for jm in 3 6 10; do echo "I3RC_JUMP_MIN=$jm" >> gpurun_out/r02_c_jump.txt; I3RC_JUMP_MIN=$jm python tools/gpu_probe.py tune landsat 16000000 "{}" >> gpurun_out/r02_c_jump.txt 2>&1; done
python tools/gpu_probe.py tune landsat 16000000 '{"skip_empty":0}' '{"resident_blocks":5}' '{"resident_blocks":5,"skip_empty":0}' '{"skip_empty":0,"resident_blocks":7,"pool_shape":2}' '{"skip_empty":0,"resident_blocks":8,"pool_shape":3}' >> gpurun_out/r02_c_jump.txt 2>&1
python tools/gpu_probe.py tune step 4000000 '{}' '{"stage_tallies":0}' >> gpurun_out/r02_c_jump.txt 2>&1
cat gpurun_out/r02_c_jump.txt
python bench.py --scaling strong --steps 8 --no-e2e --no-cpu-baseline --no-ncu > gpurun_out/r02_c_strong1.json 2> gpurun_out/r02_c_strong1.err; head -c 1500 gpurun_out/r02_c_strong1.json
