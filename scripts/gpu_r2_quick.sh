O=gpurun_out/r2e; mkdir -p $O
timeout 1200 python -m pytest tests/test_cuda_ppo.py tests/test_cuda_train.py -m gpu -x -q 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/update_launches.csv python scripts/prof_update.py > $O/prof_update.log 2>&1
grep "k_bias_grad\|k_ppo_loss\|k_train_fused\|k_adam\|gram" $O/update_launches.csv | tail -6 | awk -F'","' '{print $5, $NF}'
timeout 900 python bench.py --no-policy --no-matches --no-cpu > $O/bench_update.json 2> $O/bench_update.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_update.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"], "ppo tc ms", d["ppo_update"]["tc"]["ms_per_optimizer_step"])
PY
