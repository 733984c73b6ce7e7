set -x
mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest.log
timeout 600 python bench.py > gpurun_out/c1/bench.json 2> gpurun_out/c1/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/c1/bench_ref.json 2> gpurun_out/c1/bench_ref.err
timeout 300 python scripts/prof_kernels.py > gpurun_out/c1/kernels.jsonl 2> gpurun_out/c1/kernels.err
timeout 300 python scripts/prof_kernels.py --graph --only step,observe,mask,dup,gae > gpurun_out/c1/kernels_graph.jsonl 2>> gpurun_out/c1/kernels.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c1/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/c1/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_step|k_produce|k_legal_mask|k_dup_step' -s 12 -c 10 -f -o gpurun_out/c1/prof_single python scripts/prof_kernels.py --only step,observe,mask,dup --reps 1 > gpurun_out/c1/prof_single.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rollout_ws' -s 3 -c 2 -f -o gpurun_out/c1/prof_rollout python scripts/prof_kernels.py --only rollout --reps 2 > gpurun_out/c1/prof_rollout.log 2>&1
tail -3 gpurun_out/c1/pytest.log; cat gpurun_out/c1/bench.json | cut -c1-600
