for w in "les 1000000" "landsat 4000000"; do set -- $w
bash tools/ncu_full.sh r02_an $1 $2
ncu -i gpurun_out/r02_an_$1.ncu-rep --page source --csv > gpurun_out/r02_an_$1_sass.csv 2>/dev/null
done
ls -la gpurun_out/r02_an_*
