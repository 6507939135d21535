timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_s_tests.log 2>&1; tail -4 gpurun_out/r02_s_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
rm -f gpurun_out/r02_s_workloads.txt
for w in "planeparallel 16000000" "step 8000000" "radar 4000000" "landsat 16000000" "les-small 2000000" "les 2000000"; do timeout 300 python tools/gpu_probe.py tune $w '{}' >> gpurun_out/r02_s_workloads.txt 2>&1; done
cat gpurun_out/r02_s_workloads.txt
timeout 600 python bench.py --workload step --photons 8000000 --no-ncu > gpurun_out/r02_s_bench_step.json 2> gpurun_out/r02_s_bench_step.err; head -c 200 gpurun_out/r02_s_bench_step.json; echo
timeout 600 python bench.py --workload planeparallel --photons 16000000 --no-ncu > gpurun_out/r02_s_bench_pp.json 2> gpurun_out/r02_s_bench_pp.err; head -c 200 gpurun_out/r02_s_bench_pp.json; echo
timeout 600 python bench.py --workload radar --photons 4000000 --no-ncu > gpurun_out/r02_s_bench_radar.json 2> gpurun_out/r02_s_bench_radar.err; head -c 200 gpurun_out/r02_s_bench_radar.json; echo
