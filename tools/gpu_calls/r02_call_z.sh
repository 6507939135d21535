timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_z_tests.log 2>&1; tail -8 gpurun_out/r02_z_tests.log
rm -f gpurun_out/r02_z_ee.txt
for w in "landsat 16000000" "les 2000000" "step 8000000" "radar 4000000" "les-small 2000000"; do timeout 300 python tools/gpu_probe.py tune $w '{}' '{"le_early_exit":0}' '{"le_lower_bound":0}' >> gpurun_out/r02_z_ee.txt 2>&1; done
cat gpurun_out/r02_z_ee.txt
timeout 900 python bench.py > gpurun_out/r02_z_bench.json 2> gpurun_out/r02_z_bench.err; head -c 300 gpurun_out/r02_z_bench.json; echo
