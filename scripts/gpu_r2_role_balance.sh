#!/bin/bash
# role-balance experiment: parity of the variants, interleaved A/B, per-role clocks
mkdir -p gpurun_out/r2m
timeout 300 python -m pytest tests/test_cuda_env.py tests/test_cuda_host_api.py -x -q -m gpu -k "rollout or uniforms" > gpurun_out/r2m/pytest_roles.log 2>&1
echo "pytest rc $?" | tee -a gpurun_out/r2m/pytest_roles.log
tail -5 gpurun_out/r2m/pytest_roles.log
timeout 200 python scripts/exp_role_balance.py > gpurun_out/r2m/exp_role_balance.txt 2>&1; echo "exp_role_balance rc $?"
cat gpurun_out/r2m/exp_role_balance.txt
for r in 0 1 2 3; do
  BRL_B200_LIB=brl_b200/lib/libbrl_roletiming.so timeout 100 python scripts/exp_role_cycles.py $r >> gpurun_out/r2m/role_cycles.txt 2>&1
done
cat gpurun_out/r2m/role_cycles.txt
