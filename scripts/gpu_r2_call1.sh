set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a/smi.txt
nvidia-smi topo -m > gpurun_out/r2a/topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 600 python bench.py --no-policy --no-update > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
timeout 300 python scripts/exp_pcie_ranks.py > gpurun_out/r2a/pcie_n1.jsonl 2> gpurun_out/r2a/pcie_n1.err
tail -3 gpurun_out/r2a/pytest.log; cut -c1-1500 gpurun_out/r2a/bench.json; cat gpurun_out/r2a/pcie_n1.jsonl
