"""Host-side scalar PRNG keys for the loops that the reference drives with
`jax.random.PRNGKey/split` (src/roll_out.py:77,83; src/evaluation.py:34-35).  A key is
a 64-bit int; device kernels turn (key, global env index, step) into Philox streams."""
from __future__ import annotations

_MASK = (1 << 64) - 1


def PRNGKey(seed: int) -> int:
    return _mix(seed & _MASK)


def _mix(z: int) -> int:  # splitmix64 finaliser
    z = (z + 0x9E3779B97F4A7C15) & _MASK
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK
    return z ^ (z >> 31)


def split(key: int, num: int = 2):
    return tuple(_mix((key + (i + 1) * 0xD1B54A32D192ED03) & _MASK) for i in range(num))
