timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_v_tests.log 2>&1; tail -4 gpurun_out/r02_v_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_v_bench.json 2> gpurun_out/r02_v_bench.err; head -c 300 gpurun_out/r02_v_bench.json; echo; tail -2 gpurun_out/r02_v_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_v_bench_reference.json 2> gpurun_out/r02_v_ref.err; head -c 200 gpurun_out/r02_v_bench_reference.json; echo
timeout 600 python bench.py --scaling strong --steps 8 --workload les --report-volume --no-cpu-baseline --no-ncu --no-e2e > gpurun_out/r02_v_strong_les.json 2>/dev/null; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_v_strong_les.json")); print("strong les 1 gpu: value %.4g setup_ms %.1f"%(d["value"], d["config"]["setup_ms"]))
PY
