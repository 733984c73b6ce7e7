"""Policy/value net of the rollout and evaluation loops: the "DeepMind" 4x1024 ReLU
actor-critic (src/models.py:23-33).  Not fused into the env kernels in this round,
so per the north star it is a plain library GEMM chain (cuBLAS through torch.addmm);
parameters keep haiku's layout (`actor_critic/linear{,_1..5}` -> {'w' [in,out], 'b'}).
"""
from __future__ import annotations

import io
import pickle
from typing import Dict

import numpy as np
import torch

LAYERS = ("actor_critic/linear", "actor_critic/linear_1", "actor_critic/linear_2", "actor_critic/linear_3",
          "actor_critic/linear_4", "actor_critic/linear_5")
SIZES = ((480, 1024), (1024, 1024), (1024, 1024), (1024, 1024), (1024, 38), (1024, 1))


class _NoJaxUnpickler(pickle.Unpickler):
    """brl pickles embed `jax._src.array._reconstruct_array`; rebuild plain NumPy instead
    (ppo.py:351-362 writes them, eval.py:56-57 reads them)."""

    def find_class(self, module, name):
        if module.startswith("jax") and name == "_reconstruct_array":
            def rebuild(fun, args, state, aval=None):
                arr = fun(*args)
                arr.__setstate__(state)
                return arr
            return rebuild
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core", 1)
        return super().find_class(module, name)


def load_params(path: str, device="cuda") -> Dict[str, Dict[str, torch.Tensor]]:
    with open(path, "rb") as fh:
        raw = _NoJaxUnpickler(io.BytesIO(fh.read())).load()
    return {k: {n: torch.as_tensor(np.asarray(a, dtype=np.float32), device=device) for n, a in v.items()}
            for k, v in raw.items()}


def init_params(seed: int, device="cuda") -> Dict[str, Dict[str, torch.Tensor]]:
    """Random-init weights of the architecture (haiku default: truncated-normal, stddev 1/sqrt(fan_in))."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, (fi, fo) in zip(LAYERS, SIZES):
        w = np.clip(rng.normal(0, 1, (fi, fo)), -2, 2).astype(np.float32) / np.sqrt(fi).astype(np.float32)
        out[name] = {"w": torch.as_tensor(w, device=device), "b": torch.zeros(fo, dtype=torch.float32, device=device)}
    return out


def params_to_numpy(params) -> Dict[str, np.ndarray]:
    """flat {'w0'..'w5','b0'..'b5'} for the oracle's NumPy MLP"""
    flat = {}
    for i, name in enumerate(LAYERS):
        flat[f"w{i}"] = params[name]["w"].detach().cpu().numpy()
        flat[f"b{i}"] = params[name]["b"].detach().cpu().numpy()
    return flat


class ForwardPass:
    """`hk.without_apply_rng(hk.transform(forward_fn))` look-alike: `.apply(params, x)`."""

    def __init__(self, activation: str = "relu", model_type: str = "DeepMind", precision: str = "fp32"):
        if model_type != "DeepMind":
            raise NotImplementedError("only the DeepMind 4x1024 net is on the hot path (SURVEY 8a a16)")
        self.act = torch.relu if activation == "relu" else torch.tanh
        self.precision = precision

    def apply(self, params, x: torch.Tensor):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.precision == "tf32"
        try:
            dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
            h = x.to(dt)
            for name in LAYERS[:4]:
                h = self.act(torch.addmm(params[name]["b"].to(dt), h, params[name]["w"].to(dt)))
            logits = torch.addmm(params[LAYERS[4]]["b"].to(dt), h, params[LAYERS[4]]["w"].to(dt)).float()
            value = torch.addmm(params[LAYERS[5]]["b"].to(dt), h, params[LAYERS[5]]["w"].to(dt)).float().squeeze(-1)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return logits, value


def make_forward_pass(activation: str = "relu", model_type: str = "DeepMind", precision: str = "fp32") -> ForwardPass:
    """src/models.py:73-83"""
    return ForwardPass(activation, model_type, precision)
