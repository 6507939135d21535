python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_d_tests.log 2>&1; tail -12 gpurun_out/r02_d_tests.log
rm -f gpurun_out/r02_d_probe.txt
python tools/gpu_probe.py tune landsat 16000000 '{}' '{"vertical_shortcut":0}' >> gpurun_out/r02_d_probe.txt 2>&1
python tools/gpu_probe.py tune step 4000000 '{}' '{"vertical_shortcut":0}' >> gpurun_out/r02_d_probe.txt 2>&1
python tools/gpu_probe.py tune radar 2000000 '{}' '{"vertical_shortcut":0}' >> gpurun_out/r02_d_probe.txt 2>&1
python tools/gpu_probe.py tune les 2000000 '{}' '{"vertical_shortcut":0}' >> gpurun_out/r02_d_probe.txt 2>&1
python tools/gpu_probe.py tune planeparallel 8000000 '{}' '{"vertical_shortcut":0}' >> gpurun_out/r02_d_probe.txt 2>&1
cat gpurun_out/r02_d_probe.txt
python bench.py > gpurun_out/r02_d_bench.json 2> gpurun_out/r02_d_bench.err; head -c 400 gpurun_out/r02_d_bench.json
