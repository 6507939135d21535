#!/bin/bash
# One `ncu --set full` capture of the transport kernel per workload (run under gpurun, one GPU), after the same command
# has exited 0 without ncu.  Usage: tools/ncu_full.sh <tag> <workload> [photons]   -> gpurun_out/<tag>_<workload>.ncu-rep
# and the details page as CSV next to it.
set -u
tag=$1; wl=$2; nph=${3:-0}
cmd="python bench.py --probe-launch --workload $wl --photons $nph"
$cmd > gpurun_out/${tag}_${wl}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_transport -s 1 -c 1 -f -o gpurun_out/${tag}_${wl} $cmd > gpurun_out/${tag}_${wl}_ncu.log 2>&1
ncu -i gpurun_out/${tag}_${wl}.ncu-rep --page details --csv > gpurun_out/${tag}_${wl}_details.csv 2>/dev/null
