# bench lines of the secondary workloads at batch sizes where the 2.4 ms drain of a launch is a few per cent (16 M / 8 M / 4 M photons per step)
for w in step radar les; do python bench.py --workload $w --no-cpu-baseline > gpurun_out/r02_f6_bench_$w.json 2> gpurun_out/r02_f6_bench_$w.err; head -c 220 gpurun_out/r02_f6_bench_$w.json; echo; done
