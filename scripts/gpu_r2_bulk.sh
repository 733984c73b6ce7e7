#!/bin/bash
# bulk-copy observation path of the rollout kernel: parity, then the interleaved A/B
mkdir -p gpurun_out/r2r
timeout 400 python -m pytest tests/test_cuda_env.py tests/test_cuda_host_api.py -x -q -m gpu > gpurun_out/r2r/pytest.log 2>&1
echo "pytest rc $?" | tee -a gpurun_out/r2r/pytest.log
tail -5 gpurun_out/r2r/pytest.log
timeout 300 python scripts/exp_bulk.py > gpurun_out/r2r/exp_bulk.txt 2>&1; echo "exp rc $?"
cat gpurun_out/r2r/exp_bulk.txt
