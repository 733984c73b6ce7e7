# usage: bash scripts/gpu_r2_scale2.sh N tag   -- the driver's own launch line for N ranks, full default bench
N=$1; TAG=$2
mkdir -p gpurun_out/$TAG
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 > gpurun_out/$TAG/bench_n$N.json 2> gpurun_out/$TAG/bench_n$N.err; echo "bench rc=$?"
else
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $EXTRA > gpurun_out/$TAG/bench_n$N.json 2> gpurun_out/$TAG/bench_n$N.err; echo "bench rc=$?"
fi
tail -3 gpurun_out/$TAG/bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/$TAG/bench_n$N.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("N",d["n_gpus"],"value",d["value"],"ms",d["ms_per_step"],"e2e",e["value"],e["ms_per_step"],"ratio",e["value"]/d["value"])
for k in ("dup_selfplay_65536","league_1M","c1_eval_match"):
    if k in d: print(k, {kk:vv for kk,vv in d[k].items() if kk not in ("workload","note","pool")})
PY
