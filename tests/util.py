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


# ---- version-independent deterministic data (splitmix64), used by the hash-pinned golden cases
def splitmix64(start, count):
    z = (np.arange(count, dtype=np.uint64) + np.uint64(start)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def det_cts(seed, n, primes, count, size=2):
    K = len(primes)
    out = np.zeros((count, size, K, n + 1), dtype=np.uint64)
    raw = splitmix64(seed * 1000003, count * size * K * n).reshape(count, size, K, n)
    for j, q in enumerate(primes):
        out[:, :, j, :n] = raw[:, :, j, :] % np.uint64(q)
    return out


def det_floats(seed, count):
    return ((splitmix64(seed * 7919 + 17, count) >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2 - 1).astype(np.float32)


def det_evk(seed, n, primes, dbc=16):
    sizes = [2 * ((int(q).bit_length() + dbc - 1) // dbc) for q in primes]
    return np.concatenate([det_cts(seed + 100 + i, n, primes, 1, s).ravel() for i, s in enumerate(sizes)]), sizes, dbc


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


# the hash-pinned chain: conv(2ch 4x4 -> 3 filters 2x2, stride 1) -> avgpool 2x2/1 -> bn -> square -> fc(12 -> 4)
CHAIN = dict(xd=4, yd=4, zd=2, nf=3)


def chain_params(seed):
    return dict(conv_w=det_floats(seed + 1, 3 * 2 * 2 * 2), conv_b=det_floats(seed + 2, 3), mean=det_floats(seed + 3, 3),
                invstd=np.abs(det_floats(seed + 4, 3)) + np.float32(0.5), fc_w=det_floats(seed + 5, 4 * 12), fc_b=det_floats(seed + 6, 4))


def run_chain(backend, kind, x, p, evk, sizes, dbc):
    """Runs the chain on `backend` (kind: 'ref' | 'oracle' | 'gpu'); returns the list of per-layer outputs (numpy)."""
    outs = []
    if kind == "ref":
        r = backend
        o = r.conv(x, 4, 4, 2, 1, 1, 2, 2, 3, p["conv_w"], p["conv_b"]); outs.append(o)
        o = r.pool(o, 3, 3, 3, 1, 1, 2, 2, avg=True); outs.append(o)
        o = r.bn(o, 3, 2, 2, p["mean"], p["invstd"]); outs.append(o)
        o = r.square_layer(o, 3, 2, 2); outs.append(o)
        o = r.fc3d(o, 3, 2, 2, 4, p["fc_w"], p["fc_b"]); outs.append(o)
    elif kind == "oracle":
        r = backend
        e = r.encode_many
        d, cc = r.encode(0.25)
        o = r.conv(x, 4, 4, 2, 1, 1, 2, 2, 3, e(p["conv_w"]), e(p["conv_b"])); outs.append(o)
        o = r.pool(o, 3, 3, 3, 1, 1, 2, 2, d, cc); outs.append(o)
        o = r.bn(o, 3, 2, 2, e(p["mean"]), e(p["invstd"])); outs.append(o)
        o = r.square_layer(o, evk, sizes, dbc); outs.append(o)
        o = r.fc(o, 12, 4, e(p["fc_w"]), e(p["fc_b"])); outs.append(o)
    else:
        g = backend
        e = g.plain_encode
        t = g.conv(g.upload(x), e(p["conv_w"]), e(p["conv_b"]), 1, 4, 4, 2, 1, 1, 2, 2, 3); outs.append(g.download(t))
        t = g.pool(t, 1, 3, 3, 3, 1, 1, 2, 2, scale=e([0.25])); outs.append(g.download(t))
        t = g.bn(t, 1, 3, 2, 2, e(p["mean"]), e(p["invstd"])); outs.append(g.download(t))
        t = g.square_layer(t, g.evk_upload(evk, sizes, dbc)); outs.append(g.download(t))
        t = g.fc(t, e(p["fc_w"]), e(p["fc_b"]), 1, 12, 4); outs.append(g.download(t))
    return outs
