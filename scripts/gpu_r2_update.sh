#!/bin/bash
# PPO update kernels: parity tests, then the bench's ppo_update leg
O=gpurun_out/r2w; mkdir -p $O
timeout 900 python -m pytest tests/test_cuda_ppo.py tests/test_cuda_train.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 600 python bench.py --no-policy --no-matches --no-cpu > $O/bench_update.json 2> $O/bench_update.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_update.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"])
print(json.dumps(d["ppo_update"])[:1500])
PY
