# usage: bash scripts/gpu_r2_scale.sh N tag
N=$1; TAG=$2
mkdir -p gpurun_out/$TAG
nvidia-smi topo -m > gpurun_out/$TAG/topo.txt 2>&1
lscpu | head -30 > gpurun_out/$TAG/lscpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --no-policy --no-update --no-cpu > gpurun_out/$TAG/bench_n$N.json 2> gpurun_out/$TAG/bench_n$N.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/exp_pcie_ranks.py > gpurun_out/$TAG/pcie_n$N.jsonl 2> gpurun_out/$TAG/pcie_n$N.err; echo "pcie rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/$TAG/bench_n$N.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("value",d["value"],"ms",d["ms_per_step"])
print("e2e",e["value"],e["ms_per_step"])
for k in ("u16_uniforms","f32_payload","sync_per_call"): print(k,e[k]["value"],e[k]["ms_per_step"])
PY
cat gpurun_out/$TAG/pcie_n$N.jsonl
