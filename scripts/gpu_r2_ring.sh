#!/bin/bash
# ring hand-over experiment: parity of the ring variants, interleaved A/B, per-role clocks
mkdir -p gpurun_out/r2l
timeout 300 python -m pytest tests/test_cuda_env.py tests/test_cuda_host_api.py -x -q -m gpu -k "rollout or uniforms" > gpurun_out/r2l/pytest_ring.log 2>&1
echo "pytest rc $?" | tee -a gpurun_out/r2l/pytest_ring.log
tail -5 gpurun_out/r2l/pytest_ring.log
timeout 200 python scripts/exp_ring.py > gpurun_out/r2l/exp_ring.txt 2>&1; echo "exp_ring rc $?"
cat gpurun_out/r2l/exp_ring.txt
for r in 0 3 4; do
  BRL_B200_LIB=brl_b200/lib/libbrl_roletiming.so timeout 100 python scripts/exp_role_cycles.py $r >> gpurun_out/r2l/role_cycles.txt 2>&1
done
cat gpurun_out/r2l/role_cycles.txt
