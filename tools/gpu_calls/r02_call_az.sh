L=i3rc_monte_carlo_model_b200/libi3rc_b200.so
cp $L /tmp/keep.so
for v in ftzsqrt ftz; do cp tools/ab_libs/$v.so $L; echo "== $v"; timeout 600 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -6; done
cp /tmp/keep.so $L
bash tools/gpu_calls/r02_call_ab2.sh r02_az "landsat 16000000;step 8000000;planeparallel 16000000;les 2000000" base ftz ftzsqrt
