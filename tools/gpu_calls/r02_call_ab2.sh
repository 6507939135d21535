# A/B of whole builds on chosen workloads: r02_call_ab2.sh <tag> "<workload photons>;<workload photons>" <build> ...
L=i3rc_monte_carlo_model_b200/libi3rc_b200.so
tag=$1; wls=$2; shift; shift
cp $L /tmp/keep.so
rm -f gpurun_out/${tag}_builds.txt
for v in "$@"; do
  cp tools/ab_libs/$v.so $L
  IFS=';' read -ra WL <<< "$wls"
  for w in "${WL[@]}"; do
    echo -n "$v " >> gpurun_out/${tag}_builds.txt
    timeout 100 python tools/gpu_probe.py tune $w '{}' >> gpurun_out/${tag}_builds.txt 2>&1 || echo "FAILED/TIMEOUT" >> gpurun_out/${tag}_builds.txt
  done
done
cp /tmp/keep.so $L
cat gpurun_out/${tag}_builds.txt
