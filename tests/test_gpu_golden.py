"""GPU: the CUDA path through the C ABI against the committed golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  Byte-exact, including n = 16384."""
import json
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gpu_matches_golden_layers():
    from crcnn_b200.lib import Engine
    g = np.load(os.path.join(GOLD, "layers_n2048.npz"))
    n, t = int(g["n"]), int(g["t"])
    eng = Engine(n, [int(q) for q in g["primes"]], t)
    p = eng.plain_encode(g["enc_vals"])
    for i, want in enumerate(g["enc_plain"]):
        assert np.array_equal(eng.plain_get(p, i), want)
    assert np.array_equal(eng.plain_get_ntt(eng.plain_encode(g["conv_w"][:1]), 0), g["w0_ntt"])
    x = g["x"]
    e = eng.plain_encode
    tconv = eng.conv(eng.upload(x), e(g["conv_w"]), e(g["conv_b"]), 1, 3, 3, 2, 1, 1, 2, 2, 2)
    assert np.array_equal(eng.download(tconv).ravel(), g["conv"].ravel())
    assert np.array_equal(eng.download(eng.pool(eng.upload(x), 1, 3, 3, 2, 1, 1, 2, 2)).ravel(), g["pool"].ravel())
    assert np.array_equal(eng.download(eng.pool(eng.upload(x), 1, 3, 3, 2, 1, 1, 2, 2, scale=e([0.25]))).ravel(), g["avgpool"].ravel())
    tbn = eng.bn(tconv, 1, 2, 2, 2, e(g["bn_mean"]), e(g["bn_invstd"]))
    assert np.array_equal(eng.download(tbn).ravel(), g["bn"].ravel())
    k = eng.evk_upload(g["evk"], g["evk_sizes"], int(g["dbc"]))
    t3 = eng.square(eng.slice(tconv, 0, 2))
    assert np.array_equal(eng.download(t3), g["sq3"])
    assert np.array_equal(eng.download(eng.relinearize(t3, k)), g["relin"])
    tsq = eng.square_layer(tbn, k)
    assert np.array_equal(eng.download(tsq).ravel(), g["sq_layer"].ravel())
    tfc = eng.fc(tsq, e(g["fc_w"]), e(g["fc_b"]), 1, 8, 3)
    assert np.array_equal(eng.download(tfc).ravel(), g["fc"].ravel())
    eng.close()


@pytest.mark.parametrize("n", [4096, 8192, 16384])
def test_gpu_matches_golden_chain_hashes(n):
    from crcnn_b200.lib import Engine
    h = json.load(open(os.path.join(GOLD, "chain_hashes.json")))[str(n)]
    seed, primes, t = h["seed"], h["primes"], h["t"]
    eng = Engine(n, primes, t)
    x = util.det_cts(seed, n, primes, 2 * 4 * 4)
    evk, sizes, dbc = util.det_evk(seed, n, primes)
    assert util.sha(evk) == h["evk_sha"]
    outs = util.run_chain(eng, "gpu", x, util.chain_params(seed), evk, sizes, dbc)
    assert [util.sha(a) for a in outs] == h["layers"]
    eng.close()
