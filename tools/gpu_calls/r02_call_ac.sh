# grouped births: parity suite, then birth_min 1 (= births at once, as before) / 8 / 16 / 24 on four workloads
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02_ac_tests.log 2>&1; tail -5 gpurun_out/r02_ac_tests.log
rm -f gpurun_out/r02_ac_births.txt
for w in "landsat 16000000" "les 2000000" "step 8000000" "planeparallel 16000000"; do timeout 200 python tools/gpu_probe.py tune $w '{"birth_min":1}' '{"birth_min":8}' '{"birth_min":16}' '{"birth_min":24}' '{"birth_min":32}' >> gpurun_out/r02_ac_births.txt 2>&1; done
cat gpurun_out/r02_ac_births.txt
