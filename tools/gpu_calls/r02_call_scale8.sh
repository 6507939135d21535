# the round's last build on one 8-GPU box: the driver's N = 8 command (weak scaling, Landsat) with the c5_strong leg
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
timeout 900 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_f7_scale8.json 2> gpurun_out/r02_f7_scale8.err
python - gpurun_out/r02_f7_scale8.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d.get("c5_strong") or {}
print("gpus",d["n_gpus"],"value %.4g"%d["value"],"e2e %.4g"%d["e2e"]["value"],"ms/step %.2f"%d["ms_per_step"],"allreduce_ms %.2f"%d["config"]["allreduce_ms"], "| c5_strong:", {k:(round(v,2) if isinstance(v,float) else v) for k,v in c.items() if k in ("value","ms_total","setup_ms","trace_ms","allreduce_ms")})
PY
tail -3 gpurun_out/r02_f7_scale8.err
