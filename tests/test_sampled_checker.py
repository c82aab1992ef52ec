"""CPU: the sampled-oracle checker (oracle/sampled.py) against full oracle layers -- its window / batch / row index arithmetic
is what the headline GPU parity tests rest on.  A stand-in 'engine' whose tensors are numpy arrays runs every layer with the
oracle on the whole tensor; the checker must then agree for every sample, and must flag a corrupted ciphertext."""
import numpy as np
import pytest

import util
from oracle.port import Oracle
from oracle import sampled


class _T:
    def __init__(self, a):
        self.a = a

    def free(self):
        pass


class _Eng:
    def __init__(self, orc):
        self.K, self.stride, self.n = orc.K, orc.stride, orc.n

    def slice(self, t, first, count):
        return _T(t.a[first:first + count])

    def download(self, t):
        return t.a.copy()


class _Net:
    def __init__(self, orc, layers, params, evk_host):
        self.orc, self.layers, self.params, self.evk_host = orc, layers, params, evk_host

    def num_layers(self):
        return len(self.layers)

    def forward_layer(self, i, x, B):
        o, layer = self.orc, self.layers[i]
        kind, name = layer[0], layer[1]
        xs_ = x.a.reshape(B, -1, 2, o.K, o.stride)
        outs = []
        for b in range(B):
            if kind == "conv":
                w, bias = self.params[name]
                y = o.conv(xs_[b], *layer[2:], o.encode_many(w), o.encode_many(bias))
            elif kind == "avgpool":
                d, cc = o.encode(1.0 / (layer[7] * layer[8]))
                y = o.pool(xs_[b], *layer[2:], d, cc)
            elif kind == "pool":
                y = o.pool(xs_[b], *layer[2:])
            elif kind == "bn":
                m, v = self.params[name]
                y = o.bn(xs_[b], layer[2], layer[3], layer[4], o.encode_many(m), o.encode_many(v))
            elif kind == "square":
                y = o.square_layer(xs_[b], *self.evk_host)
            else:
                w, bias = self.params[name]
                y = o.fc(xs_[b], layer[2], layer[3], o.encode_many(w), o.encode_many(bias))
            outs.append(np.asarray(y).reshape(-1, 2, o.K, o.stride))
        return _T(np.concatenate(outs))


LAYERS = [("conv", "c1", 6, 5, 2, 2, 1, 3, 2, 3), ("avgpool", "p1", 2, 4, 3, 1, 1, 2, 2), ("bn", "b1", 3, 1, 3),
          ("square", "s1", 3, 1, 3), ("pool", "p2", 1, 3, 3, 1, 1, 1, 2), ("fc", "f1", 6, 4)]


@pytest.fixture(scope="module")
def world():
    n = 2048
    primes, t = util.PRIMES[n], util.T_FOR_N[n]
    orc = Oracle(n, primes, t)
    rng = np.random.default_rng(3)
    f = lambda k: rng.uniform(-1, 1, size=k).astype(np.float32)
    params = {"c1": (f(3 * 2 * 3 * 2), f(3)), "b1": (f(3), np.abs(f(3)) + np.float32(0.5)), "f1": (f(24), f(4))}
    evk_host = util.random_evk(rng, n, primes)
    B = 2
    x = _T(util.random_cts(rng, n, primes, B * 2 * 6 * 5))
    return orc, _Net(orc, LAYERS, params, evk_host), x, B, evk_host


def test_checker_agrees_with_full_oracle_layers(world):
    orc, net, x, B, evk_host = world
    checked, _ = sampled.check_network(_Eng(orc), orc, net, x, B, evk_host=evk_host, samples=5, seed=0, threads=4)
    assert len(checked) == len(LAYERS) and all(v >= 2 for v in checked.values()), checked


@pytest.mark.parametrize("layer", [0, 1, 5])
def test_checker_flags_corruption(world, layer):
    orc, net, x, B, evk_host = world
    eng = _Eng(orc)
    _, xin = sampled.check_network(eng, orc, net, x, B, evk_host=evk_host, samples=1, seed=0, last=layer, keep_last=True) if layer else (None, x)
    y = net.forward_layer(layer, xin, B)
    y.a[-1, 1, 0, 5] ^= np.uint64(1)     # the last ciphertext of the batch is always among the samples
    with pytest.raises(AssertionError, match="differs from the oracle"):
        sampled.check_layer(eng, orc, net, layer, xin, y, B, np.random.default_rng(0), 3, evk_host)
