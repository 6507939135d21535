rm -f gpurun_out/r02_aq_tune.txt
for w in "landsat 16000000" "les 2000000" "step 8000000"; do
timeout 300 python tools/gpu_probe.py tune $w '{}' '{"min_running_full":20}' '{"min_running_full":24}' '{"min_running_full":20,"min_running_dry":8}' '{"min_running_dry":8}' '{"min_running_dry":4}' '{"min_running_dry":24}' '{"min_running_full":20,"min_running_full_at":8}' '{"min_running_full":24,"min_running_full_at":28}' >> gpurun_out/r02_aq_tune.txt 2>&1
done
cat gpurun_out/r02_aq_tune.txt
