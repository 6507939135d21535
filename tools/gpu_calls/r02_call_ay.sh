L=i3rc_monte_carlo_model_b200/libi3rc_b200.so
cp $L /tmp/keep.so; cp tools/ab_libs/fastmath.so $L
timeout 600 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -12
cp /tmp/keep.so $L
bash tools/gpu_calls/r02_call_ab.sh r02_ay base fastmath
