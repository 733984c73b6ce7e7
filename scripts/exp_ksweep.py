"""Per-step slope and fixed overhead of the rollout kernel: time vs k_steps at 8192 envs."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from scripts.exp_quick import run
from brl_b200 import _lib
for k in (1, 2, 4, 8, 16, 32, 64, 128):
    run(8192, k, 0, reps=20, label="ws auto")
for k in (1, 4, 16, 32, 64):
    run(8192, k, _lib.tune(classic_rollout=True, epw=8), reps=20, label="tile8")
