"""GPU: the C++17 layer classes (crcnn_b200/cpp/crcnn_b200.hpp) driven the way a CrCNN program drives
the reference's classes -- istream constructors, Network::forward, Layer::forward by value,
save/loadPlaintextParameters -- byte-compared with the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import util
from oracle.port import Oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plain_record(words, n):
    """Plaintext::save format (SEAL/seal/plaintext.cpp:346-351) with coeff_count = n+1."""
    return struct.pack("<i", n + 1) + np.ascontiguousarray(words[:n + 1], dtype=np.uint64).tobytes()


def ct_record(ct, n, K):
    """Ciphertext::save format (SEAL/seal/ciphertext.cpp:103-113); the stand-in ignores the hash."""
    return bytes(32) + struct.pack("<iii", 2, n + 1, K) + np.ascontiguousarray(ct, dtype=np.uint64).tobytes()


def test_cpp_layer_classes_match_oracle(tmp_path):
    exe = str(tmp_path / "dropin_compat_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "dropin_compat_test.cpp"),
                           "-L" + os.path.join(ROOT, "crcnn_b200"), "-lcrcnn_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "crcnn_b200")])
    n, seed = 4096, 23
    primes, t = util.PRIMES[n], util.T_FOR_N[n]
    K = len(primes)
    o = Oracle(n, primes, t)
    x = util.det_cts(seed, n, primes, 2 * 4 * 4)
    p = util.chain_params(seed)
    evk, sizes, dbc = util.det_evk(seed, n, primes)
    e = o.encode_many
    cw, cb, mean, invstd, fw, fb = e(p["conv_w"]), e(p["conv_b"]), e(p["mean"]), e(p["invstd"]), e(p["fc_w"]), e(p["fc_b"])
    blob = struct.pack("<iiQ", n, K, t) + np.array(primes, dtype=np.uint64).tobytes()
    blob += struct.pack("<i", dbc) + np.array(sizes, dtype=np.int32).tobytes() + evk.tobytes()
    for ct in x:
        blob += ct_record(ct, n, K)
    per_filter = 2 * 2 * 2
    for f in range(3):  # ConvolutionalLayer::loadPlaintextParameters order
        for w in cw[f * per_filter:(f + 1) * per_filter]:
            blob += plain_record(w, n)
        blob += plain_record(cb[f], n)
    for c in range(3):  # BatchNormLayer: alternating mean / var
        blob += plain_record(mean[c], n) + plain_record(invstd[c], n)
    for r in range(4):  # FullyConnectedLayer: row weights then bias
        for w in fw[r * 12:(r + 1) * 12]:
            blob += plain_record(w, n)
        blob += plain_record(fb[r], n)
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    open(case, "wb").write(blob)
    res = subprocess.run([exe, case, out], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK"), res.stdout + res.stderr
    assert "saveload_same 1" in res.stdout and "bad_geometry_throws 1" in res.stdout and "avgpool_div_nonzero 1" in res.stdout
    want = util.run_chain(o, "oracle", x, p, evk, sizes, dbc)[-1].reshape(4, 2, K, n + 1)
    raw = open(out, "rb").read()
    rec = 32 + 12 + 2 * K * (n + 1) * 8
    assert len(raw) == 2 * 4 * rec
    for run in range(2):  # Network::forward, then layer-by-layer Layer::forward
        for i in range(4):
            r = raw[(run * 4 + i) * rec:(run * 4 + i + 1) * rec]
            assert struct.unpack("<iii", r[32:44]) == (2, n + 1, K)
            got = np.frombuffer(r[44:], dtype=np.uint64).reshape(2, K, n + 1)
            assert np.array_equal(got, want[i]), (run, i)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "dropin_seal_test")),
                    reason="oracle/_ref/dropin_seal_test not built (needs /root/reference at build time)")
@pytest.mark.parametrize("n,t", [(4096, 1 << 20), (8192, 1 << 30)])
def test_seal_mode_dropin_against_reference_classes(n, t):
    """One process, real SEAL types: the reference's layer classes vs crcnn_b200's on the same encrypted
    image and encoded weights; every layer memcmp-equal, decrypted scores / label / noise budget equal."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_seal_test")
    res = subprocess.run([exe, str(n), str(t)], capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "DROPIN OK" in res.stdout, res.stdout[-2000:] + res.stderr[-500:]
    assert res.stdout.count("bit-identical") == 9   # 7 layers + Network::forward skipped / with the re-encryption callback
    assert "forward without re-encryption policy refused: 1" in res.stdout and "re-encryption callback calls 1" in res.stdout
    assert "device re-encryption: decrypted plaintexts equal to the reference's Network::forward: 1" in res.stdout, res.stdout[-1500:]
