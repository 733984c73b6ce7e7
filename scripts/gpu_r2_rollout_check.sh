#!/bin/bash
# rollout kernel change: env parity tests, uniform-width and launch-shape timing, bench line without the secondary legs
O=gpurun_out/r2A; mkdir -p $O
timeout 400 python -m pytest tests/test_cuda_env.py tests/test_cuda_host_api.py tests/test_cuda_api.py -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
timeout 100 python scripts/exp_uniform_width.py > $O/uniform_width.txt 2>&1; cat $O/uniform_width.txt
timeout 100 python scripts/exp_lib_variants.py > $O/variants.txt 2>&1; cat $O/variants.txt
timeout 600 python bench.py --no-policy --no-matches --no-update > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("value",d["value"],"ms",d["ms_per_step"],"frac",d["roofline"]["frac"])
print("e2e u32",e["value"],e["ms_per_step"],"u16",e["u16_uniforms"]["ms_per_step"],"f32",e["f32_payload"]["ms_per_step"])
PY
