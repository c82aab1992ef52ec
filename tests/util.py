"""Shared helpers for the tests: synthetic workloads and the two checkers (oracle port, reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import port as oport  # noqa: E402
from oracle import ref as oref  # noqa: E402

PRIMES = oport.DEFAULT_PRIMES_128
T_FOR_N = {2048: 1 << 16, 4096: 1 << 18, 8192: 1 << 30, 16384: 1 << 30}


def have_ref():
    return oref.available()


def random_cts(rng, n, primes, count, size=2):
    """Uniform canonical residues in SEAL layout -- indistinguishable from ciphertexts for the evaluator."""
    K = len(primes)
    out = np.zeros((count, size, K, n + 1), dtype=np.uint64)
    for j, q in enumerate(primes):
        out[:, :, j, :n] = rng.integers(0, q, size=(count, size, n), dtype=np.uint64)
    return out


def random_evk(rng, n, primes, dbc=16):
    K = len(primes)
    sizes = [2 * ((int(q).bit_length() + dbc - 1) // dbc) for q in primes]
    parts = [random_cts(rng, n, primes, 1, s).ravel() for s in sizes]
    return np.concatenate(parts), sizes, dbc


def synthetic_image(seed, count=784):
    """i.i.d. uniform in the normalised MNIST range, seeded per image."""
    rng = np.random.default_rng(1000 + seed)
    return rng.uniform(-0.4242, 2.8215, size=count).astype(np.float32)
