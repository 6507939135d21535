# last tree of the round: GPU suite, smoke, the default bench line and the 512x512x256 one
python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_f4_tests.log 2>&1; tail -3 gpurun_out/r02_f4_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_f4_bench.json 2> gpurun_out/r02_f4_bench.err; head -c 250 gpurun_out/r02_f4_bench.json; echo
python bench.py --workload les --no-cpu-baseline > gpurun_out/r02_f4_bench_les.json 2> gpurun_out/r02_f4_bench_les.err; head -c 250 gpurun_out/r02_f4_bench_les.json; echo
bash tools/ncu_full.sh r02_f4 les 1000000
