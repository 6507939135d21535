python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "arrays or staged" 2>&1 | tail -5
python tools/gpu_probe.py tune landsat 16000000 '{}' '{"tables_in_smem":1}' '{"resident_blocks":4}' > gpurun_out/r02_h_tabsm.txt 2>&1; cat gpurun_out/r02_h_tabsm.txt
python bench.py --no-cpu-baseline --no-ncu > gpurun_out/r02_h_bench.json 2> gpurun_out/r02_h_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_h_bench.json")); print("value %.4g e2e %.4g ms/step e2e %.2f"%(d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
