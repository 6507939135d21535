bash tools/ncu_full.sh r02_ad landsat 4000000
ncu -i gpurun_out/r02_ad_landsat.ncu-rep --page source --csv > gpurun_out/r02_ad_landsat_sass.csv 2>/dev/null
tail -3 gpurun_out/r02_ad_landsat_plain.log
