timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "arrays" 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-ncu > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_m_bench.json")); e=d["e2e"]; print("value %.4g ms/step %.2f | e2e %.4g ms/step %.2f transport_ms/step %.2f"%(d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["transport_ms_per_step"]), e["meanFluxUp_last"])
PY
tail -3 gpurun_out/r02_m_bench.err
