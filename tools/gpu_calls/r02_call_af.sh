# scheduler statistics (diagnostic build, make STATS=1) and a re-tune of the round / batch thresholds with grouped births
L=i3rc_monte_carlo_model_b200/libi3rc_b200.so
cp $L /tmp/keep.so
cp tools/ab_libs/stats.so $L
rm -f gpurun_out/r02_af_sched.txt
for w in "landsat 4000000" "les 1000000" "step 4000000" "les-small 1000000"; do timeout 90 python tools/gpu_probe.py tune $w '{}' >> gpurun_out/r02_af_sched.txt 2>&1 || echo TIMEOUT >> gpurun_out/r02_af_sched.txt; done
cp /tmp/keep.so $L
cat gpurun_out/r02_af_sched.txt
timeout 300 python tools/gpu_probe.py tune landsat 16000000 '{}' '{"min_running":12}' '{"min_running":14}' '{"min_running":18}' '{"min_running":20}' '{"event_threshold":8}' '{"event_threshold":24}' '{"event_threshold":32}' '{"birth_min":12}' '{"birth_min":20}' > gpurun_out/r02_af_tune.txt 2>&1
cat gpurun_out/r02_af_tune.txt
