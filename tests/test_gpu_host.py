"""GPU: the C++17 host path behind libcrcnn_b200_host.so (crcnn_b200/cpp/host_api.cpp: Runtime + CnnBuilder + Network + BatchServer)
against the oracle and against itself: batched Network::forward_dev, the segment API on proper sub-ranges
(CrCNN/src/network.cpp:22-47 split around the re-encryption point), and the double-buffered serving loop."""
import numpy as np
import pytest

import util
from oracle.port import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny():
    from crcnn_b200 import host
    n = 2048
    primes, t = util.PRIMES[n], util.T_FOR_N[n]
    rng = np.random.default_rng(77)
    evk = util.random_evk(rng, n, primes)
    net = host.HostNetwork(n, primes, t, "ApproxPlainModel", evk=evk)
    yield n, primes, t, rng, evk, net
    net.close()


def test_batched_forward_equals_per_image_forward_and_the_python_path(tiny):
    from crcnn_b200 import nets
    from crcnn_b200.lib import Engine
    n, primes, t, rng, evk, net = tiny
    B = 3
    x = util.random_cts(rng, n, primes, B * 784)
    got, shape = net.forward(x, batch=B)
    assert shape == (1, 10, 1) and got.shape[0] == B * 10
    for b in range(B):
        one, _ = net.forward(x[b * 784:(b + 1) * 784], batch=1)
        assert np.array_equal(one, got[b * 10:(b + 1) * 10]), b
    eng = Engine(n, primes, t)
    pnet = nets.Network(eng, "ApproxPlainModel", evk=eng.evk_upload(*evk))
    want = eng.download(pnet.forward(eng.upload(x), batch=B))
    eng.close()
    assert np.array_equal(want, got)


def test_segments_around_the_reencryption_point_compose(tiny):
    """forward_dev(0, 6) then forward_dev(6, 9) == forward_dev(0, 9): the split the reference makes at layer 6 (network.cpp:23,30)."""
    n, primes, t, rng, evk, net = tiny
    x = util.random_cts(rng, n, primes, 2 * 784)
    full, _ = net.forward(x, batch=2)
    mid, mshape = net.forward(x, batch=2, first=0, last=6)
    assert mshape == (50, 4, 4)
    tail, tshape = net.forward(mid, batch=2, first=6, last=9, shape=mshape)
    assert tshape == (1, 10, 1) and np.array_equal(tail, full)
    # a middle segment on its own, checked against the oracle: conv2 -> square (layers 3, 4) of one image
    o = Oracle(n, primes, t)
    a, ashape = net.forward(x[:784], batch=1, first=0, last=3)
    b, bshape = net.forward(a, batch=1, first=3, last=5, shape=ashape)
    from crcnn_b200 import nets
    w = nets.load_weights("ApproxPlainModel")
    c = o.conv(a, 11, 11, 20, 2, 2, 3, 3, 50, o.encode_many(w["pool2_features.conv2.weight"].ravel()), o.encode_many(w["pool2_features.conv2.bias"]))
    want = o.square_layer(c, *evk)
    assert bshape == (50, 5, 5) and np.array_equal(np.asarray(want).reshape(b.shape), b)


def test_serving_loop_matches_forward(tiny):
    from crcnn_b200 import host
    n, primes, t, rng, evk, net = tiny
    B = 2
    words_in, words_out = B * 784 * net.ct_words(), B * 10 * net.ct_words()
    pin, own_in = host.pinned_array(words_in)
    pout, own_out = host.pinned_array(words_out)
    x = util.random_cts(rng, n, primes, B * 784)
    pin[:] = x.ravel()
    pout[:] = 0
    ms = net.serve(own_in.ptr, own_out.ptr, B, 4)
    want, _ = net.forward(x, batch=B)
    assert ms > 0 and np.array_equal(pout.reshape(want.shape), want)
    ms_res, per_layer = net.resident_steps(own_in.ptr, B, 1, 2)
    assert ms_res > 0 and len(per_layer) == 9 and all(v > 0 for v in per_layer)
