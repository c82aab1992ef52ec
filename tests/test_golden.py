"""CPU: the oracle port against the committed golden vectors (tests/golden/, produced from the
unmodified reference by tests/golden/make_golden.py).  Byte-exact."""
import json
import os

import numpy as np
import pytest

import util
from oracle.port import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g2048():
    return np.load(os.path.join(GOLD, "layers_n2048.npz"))


def test_oracle_matches_golden_layers(g2048):
    g = g2048
    n, t = int(g["n"]), int(g["t"])
    o = Oracle(n, [int(q) for q in g["primes"]], t)
    for v, want in zip(g["enc_vals"], g["enc_plain"]):
        assert np.array_equal(o.encode(float(v))[0], want)
    assert np.array_equal(o.plain_to_ntt(o.encode(float(g["conv_w"][0]))[0]), g["w0_ntt"])
    x = g["x"]
    e = o.encode_many
    conv = o.conv(x, 3, 3, 2, 1, 1, 2, 2, 2, e(g["conv_w"]), e(g["conv_b"]))
    assert np.array_equal(conv, g["conv"])
    assert np.array_equal(o.pool(x, 3, 3, 2, 1, 1, 2, 2), g["pool"])
    d, cc = o.encode(0.25)
    assert np.array_equal(o.pool(x, 3, 3, 2, 1, 1, 2, 2, d, cc), g["avgpool"])
    bn = o.bn(conv, 2, 2, 2, e(g["bn_mean"]), e(g["bn_invstd"]))
    assert np.array_equal(bn, g["bn"])
    sq3 = o.square(conv.reshape(-1, 2, o.K, n + 1)[:2])
    assert np.array_equal(sq3, g["sq3"])
    evk, sizes, dbc = g["evk"], g["evk_sizes"], int(g["dbc"])
    assert np.array_equal(o.relinearize(sq3, evk, sizes, dbc), g["relin"])
    sql = o.square_layer(bn, evk, sizes, dbc)
    assert np.array_equal(sql.ravel(), g["sq_layer"].ravel())
    fc = o.fc(sql, 8, 3, e(g["fc_w"]), e(g["fc_b"]))
    assert np.array_equal(fc.ravel(), g["fc"].ravel())


@pytest.mark.parametrize("n", [4096, 8192])
def test_oracle_matches_golden_chain_hashes(n):
    h = json.load(open(os.path.join(GOLD, "chain_hashes.json")))[str(n)]
    seed, primes, t = h["seed"], h["primes"], h["t"]
    assert primes == util.PRIMES[n]
    o = Oracle(n, primes, t)
    x = util.det_cts(seed, n, primes, 2 * 4 * 4)
    evk, sizes, dbc = util.det_evk(seed, n, primes)
    assert util.sha(evk) == h["evk_sha"]
    outs = util.run_chain(o, "oracle", x, util.chain_params(seed), evk, sizes, dbc)
    assert [util.sha(a) for a in outs] == h["layers"]
