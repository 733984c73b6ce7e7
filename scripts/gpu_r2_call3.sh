set -x
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -30 gpurun_out/r2c/pytest.log
timeout 900 python bench.py --no-policy --no-matches --no-cpu > gpurun_out/r2c/bench2.json 2> gpurun_out/r2c/bench2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c/bench2.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"])
print("ppo", json.dumps(d.get("ppo_update"))[:700])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c/update_launches.csv python scripts/prof_update.py > gpurun_out/r2c/prof_update.log 2>&1
tail -40 gpurun_out/r2c/update_launches.csv | cut -c1-200
