timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_aa_tests.log 2>&1; tail -8 gpurun_out/r02_aa_tests.log
rm -f gpurun_out/r02_aa_ee.txt
for w in "landsat 16000000" "les 2000000" "les-small 2000000"; do timeout 300 python tools/gpu_probe.py tune $w '{}' '{"le_early_exit":0}' >> gpurun_out/r02_aa_ee.txt 2>&1; done
cat gpurun_out/r02_aa_ee.txt
bash tools/ncu_full.sh r02_aa landsat 4000000
ncu -i gpurun_out/r02_aa_landsat.ncu-rep --page source --csv > gpurun_out/r02_aa_landsat_sass.csv 2>/dev/null
