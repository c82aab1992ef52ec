"""CPU tests of the product's host logic: parameter derivation (crcnn_b200/csrc/params.cpp) against the
oracle's independently written derivation, and the C-ABI library's exports.  No GPU calls."""
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELFTEST = os.path.join(ROOT, "build", "host_selftest")


def fnv(a):
    h = 1469598103934665603
    for x in a:
        h ^= int(x)
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.fixture(scope="module")
def selftest():
    if not os.path.exists(SELFTEST):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "crcnn_b200", "csrc"), "host_selftest"])
    return SELFTEST


@pytest.mark.parametrize("n,t", [(2048, 1 << 16), (4096, 1 << 18), (8192, 1 << 30), (16384, 1 << 30)])
def test_derived_constants_match_oracle(selftest, n, t):
    primes = port.DEFAULT_PRIMES_128[n]
    out = subprocess.check_output([selftest, str(n), str(t)] + [str(p) for p in primes], text=True)
    lines = out.strip().splitlines()
    assert lines[-1] == "OK", out[-400:]
    o = port.Oracle(n, primes, t)
    head = lines[0].split()
    K, S = int(head[1]), int(head[5])
    assert K == len(primes) and S == o.S
    for ln in lines:
        f = ln.split()
        if f[0] == "slot":
            s = int(f[1])
            base, idx = (0, s) if s < K else (1, s - K)
            assert int(f[3]) == o.modulus(base, idx)
            assert int(f[9]) == o.minimal_root(base, idx)
            for col, which in ((11, 0), (13, 1), (15, 2), (17, 3)):
                assert int(f[col]) == fnv(o.ntt_table(base, idx, which)), (s, which)
        if f[0] == "fold":
            # every SEAL default prime has the 2^k - delta shape the folded reduction of the limb-split GEMM needs; the selftest has
            # compared tcn_fold_reduce with unsigned __int128 on 200000 class-sum vectors per prime before printing OK
            j, q = int(f[1]), int(primes[int(f[1])])
            k = q.bit_length()
            assert int(f[3]) == 1 and int(f[7]) == (1 << k) - q and int(f[5]) == ((1 << k) - q) << (56 - k), ln
        if f[0] == "fold128":
            # reduce128_fold applies to every modulus of the context (coefficient and Bsk primes); checked against __int128 by the selftest
            assert int(f[3]) == 1, ln
        if f[0] == "enc":
            want, _ = o.encode(float(f[1]))
            got = np.zeros(n + 1, dtype=np.uint64)
            for pair in f[4:]:
                i, v = pair.split(":")
                got[int(i)] = int(v)
            assert np.array_equal(got, want), f[1]


def test_bad_parameters_are_rejected(selftest):
    # not an NTT prime for n = 4096; plain modulus not below the primes
    for args in (["4096", "1024", "1000003"], ["4096", str(1 << 60), "0x7fffffff380001"], ["3000", "1024", "0x7fffffff380001"]):
        p = subprocess.run([selftest] + args, capture_output=True, text=True)
        assert p.returncode != 0 and "ERROR" in p.stdout


def test_capi_exports_every_declared_symbol():
    from crcnn_b200 import lib
    header = open(os.path.join(ROOT, "include", "crcnn_b200.h")).read()
    declared = set(re.findall(r"\b(crcnn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    L = lib.load()  # dlopen only; no CUDA call is made
    for name in declared:
        assert hasattr(L, name), name
    assert declared == {s[0] for s in lib.SYMBOLS}


def test_missing_gpu_fails_loudly():
    """On a box without a CUDA device the product refuses to run instead of falling back."""
    import ctypes as C
    from crcnn_b200 import lib
    L = lib.load()
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    q = np.array(port.DEFAULT_PRIMES_128[4096], dtype=np.uint64)
    h = C.c_void_p()
    rc = L.crcnn_ctx_create(4096, 2, q.ctypes.data_as(lib._u64p), 1 << 18, 0, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in L.crcnn_last_error(None)
