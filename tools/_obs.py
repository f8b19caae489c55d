"""Synthetic observations for the timing tools (the same recipe as bench.py / BASELINE.json:
README seed-42 sample for n = 50, default_rng(42) otherwise).  Stated here so that the tools
under tools/ never import the oracle package -- the ones that need it live in tests/tools/."""
import numpy as np


def observations(n: int = 50, seed: int = 42) -> np.ndarray:
    if n == 50:
        return np.random.RandomState(seed).normal(20.0, 2.0, n).astype(np.float32)
    return np.random.default_rng(seed).normal(20.0, 2.0, n).astype(np.float32)
