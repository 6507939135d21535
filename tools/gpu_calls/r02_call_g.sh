python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_g_tests.log 2>&1; tail -6 gpurun_out/r02_g_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_g_smoke.log 2>&1; tail -3 gpurun_out/r02_g_smoke.log
python bench.py > gpurun_out/r02_g_bench.json 2> gpurun_out/r02_g_bench.err; head -c 300 gpurun_out/r02_g_bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_g_bench_reference.json 2> gpurun_out/r02_g_ref.err; head -c 200 gpurun_out/r02_g_bench_reference.json; echo
python tools/gpu_probe.py workloads > gpurun_out/r02_g_workloads.txt 2>&1; cat gpurun_out/r02_g_workloads.txt
tools/ncu_full.sh r02_final landsat 4000000; tools/ncu_full.sh r02_final planeparallel 4000000; tools/ncu_full.sh r02_final les 1000000
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > gpurun_out/r02_g_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > gpurun_out/r02_g_ncu_launch.log 2>&1
ls -la gpurun_out | tail -20
