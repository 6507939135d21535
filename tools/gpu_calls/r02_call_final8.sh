# last tree of the round (80-slot kernel for Landsat): GPU suite, smoke, default bench line, launch list
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_f8_tests.log 2>&1; tail -3 gpurun_out/r02_f8_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_f8_bench.json 2> gpurun_out/r02_f8_bench.err; head -c 250 gpurun_out/r02_f8_bench.json; echo
bash tools/ncu_full.sh r02_f8 landsat 4000000
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > gpurun_out/r02_f8_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_f8_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-ncu > gpurun_out/r02_f8_ncu_launch.log 2>&1
