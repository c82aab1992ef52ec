"""GPU: CnnBuilder (crcnn_b200/cpp/cnn_builder.hpp) -- weights from an .h5 file, encoded on the device, network assembled from the
drop-in layer classes -- against the Python predict path (crcnn_b200/nets.py) on the same weights, and its encoder against the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

import util
from h5write import write_h5
from oracle.port import Oracle
from test_gpu_cpp_dropin import ct_record

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("topology,model", [("Approx", "ApproxPlainModel"), ("Tiny", "PlainModelTiny")])
def test_builder_network_matches_python_path(tmp_path, topology, model):
    from crcnn_b200 import nets
    from crcnn_b200.lib import Engine
    exe = str(tmp_path / "builder_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", "builder_test.cpp"),
                           "-L" + os.path.join(ROOT, "crcnn_b200"), "-lcrcnn_b200", "-Wl,-rpath," + os.path.join(ROOT, "crcnn_b200")])
    n = 2048
    primes, t = util.PRIMES[n], util.T_FOR_N[n]
    K = len(primes)
    w = nets.load_weights(model)
    h5 = str(tmp_path / "weights.h5")
    write_h5(h5, w)
    rng = np.random.default_rng(11)
    x = util.random_cts(rng, n, primes, 784)
    evk, sizes, dbc = util.random_evk(rng, n, primes)
    blob = struct.pack("<iiQ", n, K, t) + np.array(primes, dtype=np.uint64).tobytes()
    blob += struct.pack("<i", dbc) + np.array(sizes, dtype=np.int32).tobytes() + evk.tobytes()
    blob += struct.pack("<iii", 1, 28, 28) + b"".join(ct_record(ct, n, K) for ct in x)
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    open(case, "wb").write(blob)
    res = subprocess.run([exe, case, h5, topology, out], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("OK"), res.stdout + res.stderr
    assert "fc4_bias_count 10" in res.stdout and "saveload_same 1" in res.stdout and "missing_tensor_throws 1" in res.stdout
    assert "image_file_same 1" in res.stdout
    for flag in ("reencryption_policy_enforced 1", "segments_compose 1", "reencrypt_callback_path 1", "forward_batch 1"):
        assert flag in res.stdout, res.stdout
    raw = open(out, "rb").read()
    # the two encoded parameters: Plaintext::save records of n+1 words, equal to the oracle's FractionalEncoder restatement
    o = Oracle(n, primes, t)
    prec = 4 + (n + 1) * 8
    k0 = np.frombuffer(raw[4:prec], dtype=np.uint64)
    b0 = np.frombuffer(raw[prec + 4:2 * prec], dtype=np.uint64)
    want_k, want_b = o.encode_many(w["pool1_features.conv1.weight"].ravel()[:1]), o.encode_many(w["pool1_features.conv1.bias"][:1])
    assert struct.unpack("<i", raw[:4])[0] == n + 1
    assert np.array_equal(k0, np.asarray(want_k[0][:n + 1], dtype=np.uint64)) and np.array_equal(b0, np.asarray(want_b[0][:n + 1], dtype=np.uint64))
    # the scores: same bytes as the Python predict path on the same weights
    eng = Engine(n, primes, t)
    net = nets.Network(eng, model, weights=w, evk=eng.evk_upload(evk, sizes, dbc))
    want = eng.download(net.forward(eng.upload(x), batch=1))
    eng.close()
    rec = 32 + 12 + 2 * K * (n + 1) * 8
    body = raw[2 * prec:]
    assert len(body) == 10 * rec
    for i in range(10):
        got = np.frombuffer(body[i * rec + 44:(i + 1) * rec], dtype=np.uint64).reshape(2, K, n + 1)
        assert np.array_equal(got, want[i]), i
