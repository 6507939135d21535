timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_p_tests.log 2>&1; tail -25 gpurun_out/r02_p_tests.log
timeout 300 python tools/gpu_probe.py tune les 2000000 '{}' '{"resident_blocks":5}' > gpurun_out/r02_p_slab.txt 2>&1
I3RC_SPLIT_LAYERS=0 timeout 300 python tools/gpu_probe.py tune les-small 2000000 '{}' >> gpurun_out/r02_p_slab.txt 2>&1
timeout 300 python tools/gpu_probe.py tune les-small 2000000 '{}' '{"slab_jump":0}' '{"resident_blocks":5}' >> gpurun_out/r02_p_slab.txt 2>&1
I3RC_SPLIT_LAYERS=0 timeout 300 python tools/gpu_probe.py tune radar 4000000 '{}' >> gpurun_out/r02_p_slab.txt 2>&1
timeout 300 python tools/gpu_probe.py tune radar 4000000 '{}' '{"slab_jump":0}' '{"resident_blocks":5}' >> gpurun_out/r02_p_slab.txt 2>&1
cat gpurun_out/r02_p_slab.txt
