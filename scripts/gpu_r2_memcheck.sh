O=gpurun_out/r2i; mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cuda_host_api.py tests/test_cuda_xla.py tests/test_cuda_mlp.py tests/test_cuda_ppo.py tests/test_board_log.py -m gpu -x -q -k "not 20_000 and not 13000 and not 70_001 and not 4096 and not 2048 and not 8192 and not update_step" > $O/memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" $O/memcheck.txt | tail -8
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_cuda_api.py -m gpu -x -q -k "c1_eval or simple_evaluate or simple_duplicate or league" > $O/memcheck_eval.txt 2>&1; echo "memcheck eval rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" $O/memcheck_eval.txt | tail -5
