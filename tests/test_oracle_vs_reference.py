"""Pins the CPU oracle against the UNMODIFIED reference compiled into oracle/_ref (SEAL 2.3.1 +
CrCNN layers).  Runs wherever oracle/_ref/libcrcnn_ref.so exists (built in the dev container by
oracle/Makefile.ref; the prebuilt .so also travels to the GPU box).  Byte-for-byte comparisons."""
import numpy as np
import pytest

from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")

CASES = [(2048, 1 << 16), (4096, 1 << 18), (8192, 1 << 30)]


@pytest.fixture(scope="module", params=CASES, ids=lambda c: "n%d" % c[0])
def pair(request):
    n, t = request.param
    r = ref.Ref(n, t, seed=n)
    assert r.primes == port.DEFAULT_PRIMES_128[n]  # coeff_modulus_128(n)
    o = port.Oracle(n, r.primes, t)
    return r, o, np.random.default_rng(n)


def test_encoder(pair):
    r, o, rng = pair
    vals = [0.0867, -3.25, 0.0, 1.0, -1.0, 7.0, 0.25, 2.8215, -0.4242, 1 / 9.0, -17.5, 1e-9, 123456.789]
    vals += list(rng.uniform(-3, 3, size=40).astype(np.float32))
    for v in vals:
        a, ca = r.encode(float(v)); b, cb = o.encode(float(v))
        assert ca == cb and np.array_equal(a, b), v


def test_evaluator_ops(pair):
    r, o, rng = pair
    x = r.encrypt(rng.uniform(-0.4242, 2.8215, size=5).astype(np.float32))
    xn = r.ct_transform(x)
    assert np.array_equal(xn, o.ct_transform(x))
    assert np.array_equal(o.ct_transform(xn, inverse=True), x)
    assert np.array_equal(r.ct_transform(xn, inverse=True), x)
    w, _ = r.encode(0.0867)
    wn = r.plain_to_ntt(w)
    assert np.array_equal(wn, o.plain_to_ntt(w))
    assert np.array_equal(r.multiply_plain_ntt(xn, wn), o.multiply_plain_ntt(xn, wn))
    dense = np.zeros(r.n + 1, dtype=np.uint64)
    dense[:r.n] = rng.integers(0, r.t, size=r.n, dtype=np.uint64)
    for plain in (w, dense, np.array([r.t - 2], dtype=np.uint64), np.array([3], dtype=np.uint64)):
        for op in ("mul", "add", "sub"):
            assert np.array_equal(r.plain_op(x, plain, op), o.plain_op(x, plain, op)), (op, len(plain))
    assert np.array_equal(r.add_many(x), o.add_many(x))


def test_square_relinearize(pair):
    r, o, rng = pair
    vals = rng.uniform(-0.4242, 2.8215, size=3).astype(np.float32)
    x = r.encrypt(vals)
    s3 = r.square(x)
    assert np.array_equal(s3, o.square(x))
    ev, sizes, dbc = r.evk()
    rl = r.relinearize(s3)
    assert np.array_equal(rl, o.relinearize(s3, ev, sizes, dbc))
    dec, budget = r.decrypt(rl)
    assert np.allclose(dec, vals.astype(np.float64) ** 2, atol=1e-4) and (budget > 0).all()
    # uniformly random residues exercise every branch of the base conversions
    y = np.zeros_like(x)
    for j, q in enumerate(r.primes):
        y[:, :, j, :r.n] = rng.integers(0, q, size=(3, 2, r.n), dtype=np.uint64)
    s3 = r.square(y)
    assert np.array_equal(s3, o.square(y))
    assert np.array_equal(r.relinearize(s3), o.relinearize(s3, ev, sizes, dbc))


def test_layers(pair):
    r, o, rng = pair
    xi = r.encrypt(rng.uniform(-0.4242, 2.8215, size=2 * 3 * 3).astype(np.float32))
    wv = rng.uniform(-1, 1, size=2 * 2 * 2 * 2).astype(np.float32); bv = rng.uniform(-1, 1, size=2).astype(np.float32)
    a = r.conv(xi, 3, 3, 2, 1, 1, 2, 2, 2, wv, bv)
    assert np.array_equal(a, o.conv(xi, 3, 3, 2, 1, 1, 2, 2, 2, o.encode_many(wv), o.encode_many(bv)))
    wf = rng.uniform(-1, 1, size=3 * 18).astype(np.float32); bf = rng.uniform(-1, 1, size=3).astype(np.float32)
    assert np.array_equal(r.fc3d(xi, 2, 3, 3, 3, wf, bf), o.fc(xi, 18, 3, o.encode_many(wf), o.encode_many(bf)))
    assert np.array_equal(r.pool(xi, 3, 3, 2, 1, 1, 2, 2), o.pool(xi, 3, 3, 2, 1, 1, 2, 2))
    d, cc = o.encode(0.25)
    assert np.array_equal(r.pool(xi, 3, 3, 2, 1, 1, 2, 2, avg=True), o.pool(xi, 3, 3, 2, 1, 1, 2, 2, d, cc))
    # a window area that is not a power of two: the reference encodes the DOUBLE 1./(xf*yf) (avgPoolingLayer.cpp:10-13); the float32
    # value of 1/9 has other base-3 digits from about the 15th on and must NOT reproduce the reference's bytes
    d9, cc9 = o.encode(1.0 / 9.0)
    want9 = r.pool(xi, 3, 3, 2, 1, 1, 3, 3, avg=True)
    assert np.array_equal(want9, o.pool(xi, 3, 3, 2, 1, 1, 3, 3, d9, cc9))
    f9, fc9 = o.encode(float(np.float32(1.0 / 9.0)))
    assert not np.array_equal(d9, f9) and not np.array_equal(want9, o.pool(xi, 3, 3, 2, 1, 1, 3, 3, f9, fc9))
    assert np.array_equal(r.bn(xi, 2, 3, 3, [0.3, -0.2], [1.7, 0.9]),
                          o.bn(xi, 2, 3, 3, o.encode_many([0.3, -0.2]), o.encode_many([1.7, 0.9])))
    ev, sizes, dbc = r.evk()
    assert np.array_equal(r.square_layer(xi[:4], 1, 2, 2).ravel(), o.square_layer(xi[:4], ev, sizes, dbc).ravel())


@pytest.mark.parametrize("n,t", [(2048, 1 << 16), (4096, 1 << 20), (8192, 1 << 30)])
def test_reencryption_steps_against_seal(n, t):
    """the oracle's Decryptor::decrypt, FractionalEncoder::decode and Encryptor::encrypt restatements against
    SEAL's own objects -- decrypt bit for bit (fresh and used ciphertexts), decode -> float -> encode bit for bit, and an
    encryption with explicit sampled polynomials that SEAL's Decryptor opens to the same plaintext with a fresh noise budget."""
    r = ref.Ref(n, t, seed=5)
    o = port.Oracle(n, r.primes, t)
    sk, pk = r.keys()
    rng = np.random.default_rng(n)
    vals = rng.uniform(-3, 3, 6).astype(np.float32)
    cts = r.encrypt(vals)
    _, fresh_budget, plain = r.decrypt(cts, want_plain=True)
    assert np.array_equal(o.decrypt(cts, sk), plain)
    w = rng.uniform(-1, 1, 12).astype(np.float32); b = rng.uniform(-1, 1, 2).astype(np.float32)
    y = r.fc(cts, 6, 2, w, b, th=2).reshape(2, 2, r.K, r.stride)
    dec_vals, _, plain2 = r.decrypt(y, want_plain=True)
    assert np.array_equal(o.decrypt(y, sk), plain2)
    for i in range(2):
        assert o.decode(plain2[i]) == dec_vals[i]
        want_plain, want_val = r.reencode(y[i])
        enc, _ = o.encode(float(np.float32(o.decode(plain2[i]))))
        assert np.array_equal(enc, want_plain) and np.float32(o.decode(plain2[i])) == np.float32(want_val)
        u = rng.integers(-1, 2, n).astype(np.int8)
        e0 = np.clip(np.trunc(rng.normal(0, 3.19, n)), -19, 19).astype(np.int8)
        e1 = np.clip(np.trunc(rng.normal(0, 3.19, n)), -19, 19).astype(np.int8)
        ct = o.encrypt(want_plain, pk, u, e0, e1)
        _, budget, back = r.decrypt(ct[None], want_plain=True)
        assert np.array_equal(back[0], want_plain) and abs(int(budget[0]) - int(fresh_budget[0])) <= 1
