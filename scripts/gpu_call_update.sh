set -x
mkdir -p gpurun_out/m
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/m/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/m/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/m/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/m/smoke.log
timeout 600 python bench.py > gpurun_out/m/bench.json 2> gpurun_out/m/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/m/bench_ref.json 2> gpurun_out/m/bench_ref.err
timeout 300 python scripts/exp_update.py > gpurun_out/m/update_timing.json 2> gpurun_out/m/update_timing.err
timeout 200 python scripts/exp_train_trace.py 0 > gpurun_out/m/train_trace.txt 2>&1
timeout 200 python scripts/exp_grad_error.py 64 > gpurun_out/m/grad_error_b64.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m/update_launches.csv python scripts/exp_update.py --iters 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_train_fused -s 2 -c 2 -f -o gpurun_out/m/prof_train_fused python scripts/prof_update.py 0 2 > gpurun_out/m/prof_train_fused.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/m/launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/m/bench_under_ncu.log 2>&1
tail -3 gpurun_out/m/pytest.log; cat gpurun_out/m/smoke.log; cut -c1-400 gpurun_out/m/bench.json
