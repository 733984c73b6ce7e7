"""Build a debug variant of the library from the same sources with extra -D flags:
    python scripts/build_variant.py <name> <flag> [<flag> ...]     -> brl_b200/lib/libbrl_<name>.so
Load it with BRL_B200_LIB=<path> (brl_b200/_lib.py).  Used for -DBRL_PLAIN_EPISODE_LOAD (scripts/racecheck_prefetch.sh),
-DBRL_GRAM_TIMING (scripts/exp_gram_phases.py) and -DBRL_ROLE_TIMING (scripts/exp_role_timing.py)."""
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import build as b  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(b.LIB_DIR, f"libbrl_{name}.so")
os.makedirs(b.LIB_DIR, exist_ok=True)
with tempfile.TemporaryDirectory() as tmp:
    procs, objs = [], []
    for src in b.SOURCES:
        obj = os.path.join(tmp, os.path.splitext(src)[0] + ".o")
        procs.append(subprocess.Popen([b._nvcc(), *b.NVCC_FLAGS, *flags, "-c", os.path.join(b.CSRC, src), "-o", obj]))
        objs.append(obj)
    assert all(p.wait() == 0 for p in procs)
    subprocess.run([b._nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs], check=True)
print(out)
