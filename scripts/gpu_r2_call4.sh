set -x
O=gpurun_out/r2d; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 900 python bench.py --no-policy --no-matches --no-cpu > $O/bench_update.json 2> $O/bench_update.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_update.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"], "ppo tc ms", d["ppo_update"]["tc"]["ms_per_optimizer_step"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/update_launches.csv python scripts/prof_update.py > $O/prof_update.log 2>&1
grep -c . $O/update_launches.csv; grep "k_bias_grad\|k_ppo_loss\|k_train_fused\|k_adam" $O/update_launches.csv | tail -8 | awk -F'","' '{print $5, $NF}'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-update --no-policy > $O/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rollout_ws' -s 3 -c 2 -f -o $O/prof_rollout python scripts/prof_kernels.py --only rollout --reps 2 > $O/prof_rollout.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_fused' -s 3 -c 2 -f -o $O/prof_mlp_fused python scripts/prof_mlp.py > $O/prof_mlp.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum,lts__t_bytes.sum --clock-control none -k regex:'k_legal_mask' -c 6 --csv --log-file $O/mask_8M.csv python scripts/prof_kernels.py --only mask --big 8388608 --reps 3 > $O/mask_8M.log 2>&1
cat $O/mask_8M.csv | tail -12 | cut -c1-260
timeout 300 python scripts/prof_kernels.py --only mask --big 8388608 --reps 10 > $O/mask_8M_timing.jsonl 2>&1; cat $O/mask_8M_timing.jsonl
ls -la $O
