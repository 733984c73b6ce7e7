set -x
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -30 gpurun_out/r2c/pytest.log
timeout 900 python bench.py > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2c/bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c/bench.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"])
for k in ("dup_selfplay_65536","c1_eval_match"): print(k, json.dumps(d.get(k))[:900])
print("policy", json.dumps(d.get("policy_rollout"))[:1200])
print("ppo", json.dumps(d.get("ppo_update"))[:1200])
PY
