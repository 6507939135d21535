python bench.py --no-cpu-baseline --no-ncu > gpurun_out/r02_l_bench.json 2> gpurun_out/r02_l_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_l_bench.json")); e=d["e2e"]; print("value %.4g ms/step %.2f | e2e %.4g ms/step %.2f transport_ms/step %.2f"%(d["value"], d["ms_per_step"], e["value"], e["ms_per_step"], e["transport_ms_per_step"]))
PY
