# A/B of whole builds: tools/ab_libs/<name>.so are copied over the product library one after the other
# usage: r02_call_ab.sh <tag> <build> <build> ...
L=i3rc_monte_carlo_model_b200/libi3rc_b200.so
tag=$1; shift
cp $L /tmp/keep.so
rm -f gpurun_out/${tag}_builds.txt
for v in "$@"; do
  cp tools/ab_libs/$v.so $L
  for w in "landsat 16000000" "les 2000000" "les-small 2000000" "step 8000000" "radar 4000000" "planeparallel 16000000"; do
    echo -n "$v " >> gpurun_out/${tag}_builds.txt
    timeout 100 python tools/gpu_probe.py tune $w '{}' >> gpurun_out/${tag}_builds.txt 2>&1 || echo "FAILED/TIMEOUT" >> gpurun_out/${tag}_builds.txt
  done
done
cp /tmp/keep.so $L
cat gpurun_out/${tag}_builds.txt
