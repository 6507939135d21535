rm -f gpurun_out/r02_ah_tune.txt
for w in "landsat 16000000" "les 2000000" "les-small 2000000" "step 8000000" "radar 4000000"; do
timeout 300 python tools/gpu_probe.py tune $w '{}' '{"event_threshold":8,"birth_low":8}' '{"event_threshold":4,"birth_low":4}' '{"event_threshold":2,"birth_low":2}' '{"event_threshold":1,"birth_low":0}' '{"event_threshold":4,"birth_low":0}' '{"event_threshold":8,"birth_low":0}' '{"event_threshold":4,"birth_low":4,"birth_min":24}' >> gpurun_out/r02_ah_tune.txt 2>&1
done
cat gpurun_out/r02_ah_tune.txt
