#!/bin/bash
# end-of-round check on one GPU: the whole GPU suite, smoke(), the default bench and the reference arm
O=gpurun_out/r2F; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"],"clocks",d["clocks"])
for k in ("dup_selfplay_65536","c1_eval_match","policy_rollout","ppo_update"):
    print(k, json.dumps(d.get(k))[:260])
r=json.loads(open("$O/bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r.get("value"), r.get("cpu_baseline",{}).get("cores"))
PY
