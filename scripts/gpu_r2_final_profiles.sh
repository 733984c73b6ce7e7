O=gpurun_out/r2k; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_mlp_fused' -s 3 -c 2 -f -o $O/prof_mlp_fused_v2 python scripts/prof_mlp.py > $O/prof_mlp.log 2>&1; echo "ncu mlp rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-update > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"frac",d["roofline"]["frac"], "clocks", d["clocks"])
for k in ("dup_selfplay_65536","c1_eval_match"): print(k, {kk:vv for kk,vv in d[k].items() if kk not in ("workload","note")})
pr=d["policy_rollout"]; print("policy", pr["tc"]["ms_per_rollout_plus_gae"], pr["tc"]["forward_ms_8192"], "ppo", d["ppo_update"]["tc"]["ms_per_optimizer_step"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("build"))
PY
grep -c . $O/launches_bench.csv
