# 2-GPU call: weak default, strong Landsat, strong C5 with volume absorption (577 MB float64 moment buffer)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29533"
$TR --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_e_weak2.json 2> gpurun_out/r02_e_weak2.err; head -c 300 gpurun_out/r02_e_weak2.json; echo
python bench.py --scaling strong --steps 8 --no-cpu-baseline --no-ncu > gpurun_out/r02_e_strong_landsat_1.json 2> gpurun_out/r02_e_s1.err
$TR --nproc-per-node 2 bench.py --gpus 2 --scaling strong --steps 8 > gpurun_out/r02_e_strong_landsat_2.json 2> gpurun_out/r02_e_s2.err
python bench.py --scaling strong --steps 8 --workload les --report-volume --no-cpu-baseline --no-ncu > gpurun_out/r02_e_strong_les_1.json 2> gpurun_out/r02_e_l1.err
$TR --nproc-per-node 2 bench.py --gpus 2 --scaling strong --steps 8 --workload les --report-volume > gpurun_out/r02_e_strong_les_2.json 2> gpurun_out/r02_e_l2.err
for f in gpurun_out/r02_e_strong_*.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); c=d["config"]
print(sys.argv[1], "gpus",d["n_gpus"],"value %.4g"%d["value"],"ms/step %.2f"%d["ms_per_step"],"setup_ms %.1f"%c["setup_ms"],"allreduce_ms %.2f"%c["allreduce_ms"],"bytes",c["allreduce_bytes"],"batches",c["total_batches"])
PY
done
tail -3 gpurun_out/r02_e_*.err
