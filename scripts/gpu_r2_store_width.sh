#!/bin/bash
# 256-bit observation stores: parity of every env kernel, then the interleaved A/B
mkdir -p gpurun_out/r2o
timeout 400 python -m pytest tests/test_cuda_env.py tests/test_cuda_host_api.py tests/test_cuda_api.py -x -q -m gpu > gpurun_out/r2o/pytest.log 2>&1
echo "pytest rc $?" | tee -a gpurun_out/r2o/pytest.log
tail -5 gpurun_out/r2o/pytest.log
timeout 300 python scripts/exp_store_width.py > gpurun_out/r2o/exp_store_width.txt 2>&1; echo "exp rc $?"
cat gpurun_out/r2o/exp_store_width.txt
