# final records of round 2 (build with births in groups + direction-fastest bounds)
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_f3_tests.log 2>&1; tail -4 gpurun_out/r02_f3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_f3_smoke.log 2>&1; tail -2 gpurun_out/r02_f3_smoke.log
python bench.py > gpurun_out/r02_f3_bench.json 2> gpurun_out/r02_f3_bench.err; head -c 300 gpurun_out/r02_f3_bench.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_f3_bench_reference.json 2> gpurun_out/r02_f3_ref.err; head -c 200 gpurun_out/r02_f3_bench_reference.json; echo
for w in planeparallel step radar les; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02_f3_bench_$w.json 2> gpurun_out/r02_f3_bench_$w.err; head -c 200 gpurun_out/r02_f3_bench_$w.json; echo; done
python tools/gpu_probe.py workloads > gpurun_out/r02_f3_workloads.txt 2>&1; cat gpurun_out/r02_f3_workloads.txt
tools/ncu_full.sh r02_f3 landsat 4000000; tools/ncu_full.sh r02_f3 planeparallel 4000000; tools/ncu_full.sh r02_f3 les 1000000
for w in landsat les planeparallel; do ncu -i gpurun_out/r02_f3_$w.ncu-rep --page source --csv > gpurun_out/r02_f3_${w}_sass.csv 2>/dev/null; done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > gpurun_out/r02_f3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_f3_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > gpurun_out/r02_f3_ncu_launch.log 2>&1
ls gpurun_out | grep r02_f3 | wc -l
