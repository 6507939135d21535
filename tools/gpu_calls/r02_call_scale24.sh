TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
for n in 2 4; do
timeout 600 $TR --nproc-per-node $n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r02_f3_scale$n.json 2> gpurun_out/r02_f3_scale$n.err
python - gpurun_out/r02_f3_scale$n.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); c=d.get("c5_strong") or {}
print("gpus",d["n_gpus"],"value %.4g"%d["value"],"e2e %.4g"%d["e2e"]["value"],"ms/step %.2f"%d["ms_per_step"],"allreduce_ms %.2f"%d["config"]["allreduce_ms"], "| c5_strong:", {k:(round(v,2) if isinstance(v,float) else v) for k,v in c.items() if k in ("value","ms_total","setup_ms","trace_ms","allreduce_ms")})
PY
done
