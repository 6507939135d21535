# the whole GPU suite on the checked build (make CHECK=1: index invariants count as bad photons), with and without the layer table forced
L=i3rc_monte_carlo_model_b200/libi3rc_b200.so
cp $L /tmp/keep.so; cp tools/ab_libs/check.so $L
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_check_tests.log 2>&1; tail -4 gpurun_out/r02_check_tests.log
I3RC_SPLIT_LAYERS=2 timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_check_tests_split.log 2>&1; tail -4 gpurun_out/r02_check_tests_split.log
timeout 200 python tools/gpu_probe.py workloads 2>&1 | tail -7
cp /tmp/keep.so $L
