"""GPU parity tests proper: every C-ABI entry point against the CPU oracle, byte for byte.

The checker is oracle/liboracle.so (plain-C restatement, itself pinned to the compiled reference and
to the golden fixtures by the CPU tests).  Where oracle/_ref/libcrcnn_ref.so is present the real
reference (SEAL 2.3.1 + CrCNN layers) encrypts the inputs and decrypts the outputs as well.
Bar: bit-exact (integer residues) -- np.array_equal, no tolerance.
"""
import os

import numpy as np
import pytest

from util import PRIMES, T_FOR_N, random_cts, random_evk, have_ref
from oracle.port import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[2048, 4096, 8192])
def env(request):
    from crcnn_b200.lib import Engine
    n = request.param
    primes, t = PRIMES[n], T_FOR_N[n]
    eng = Engine(n, primes, t)
    orc = Oracle(n, primes, t)
    rng = np.random.default_rng(n)
    yield n, primes, t, eng, orc, rng
    eng.close()


def test_ntt_tables_match_oracle(env):
    n, primes, t, eng, orc, rng = env
    K = len(primes)
    assert eng.S == orc.S
    for slot in range(K + orc.S):
        base, idx = (0, slot) if slot < K else (1, slot - K)
        for which in range(4):
            assert np.array_equal(eng.ntt_table(slot, which), orc.ntt_table(base, idx, which)), (slot, which)


def test_upload_download_roundtrip(env):
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 5)
    tx = eng.upload(x)
    assert np.array_equal(eng.download(tx), x)


def test_transform_to_from_ntt(env):
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 7)
    tx = eng.upload(x)
    eng.to_ntt(tx)
    want = orc.ct_transform(x)
    got = eng.download(tx, ntt_form=True)
    assert np.array_equal(got, want)
    assert np.array_equal(eng.download(tx), x)  # download converts a copy back
    eng.from_ntt(tx)
    assert np.array_equal(eng.download(tx), x)


def test_plain_to_ntt_and_encode(env):
    n, primes, t, eng, orc, rng = env
    vals = np.array([0.0867, -3.25, 0.0, 1.0, -1.0, 7.0, 0.25, 2.8215, -0.4242, 1 / 9.0, -17.5, 1e-9], dtype=np.float32)
    p = eng.plain_encode(vals)
    for i, v in enumerate(vals):
        want, _ = orc.encode(float(v))
        got = eng.plain_get(p, i)
        assert np.array_equal(got, want), v
        assert np.array_equal(eng.plain_get_ntt(p, i), orc.plain_to_ntt(want)), v


@pytest.mark.parametrize("op", ["mul", "add", "sub"])
def test_plain_ops(env, op):
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 3)
    w, _ = orc.encode(0.0867)
    dense = np.zeros(n + 1, dtype=np.uint64)
    dense[:n] = rng.integers(0, t, size=n, dtype=np.uint64)  # a fully dense plaintext too
    const = np.array([t - 2], dtype=np.uint64)               # SEAL's constant-plaintext branch
    for plain in (w, dense, const):
        p = eng.plain_upload(plain)
        tx = eng.upload(x)
        eng.plain_op(tx, p, 0, op)
        assert np.array_equal(eng.download(tx), orc.plain_op(x, plain, op)), (op, len(plain))


def test_multiply_plain_ntt_semantics(env):
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 2)
    w, _ = orc.encode(-0.731)
    xn = orc.ct_transform(x)
    want = orc.multiply_plain_ntt(xn, orc.plain_to_ntt(w))
    tx = eng.upload(xn, ntt_form=True)
    eng.plain_op(tx, eng.plain_upload(w), 0, "mul")
    assert np.array_equal(eng.download(tx, ntt_form=True), want)


def test_add_many(env):
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 7)
    got = eng.download(eng.add_many(eng.upload(x)))
    assert np.array_equal(got[0], orc.add_many(x))
    # more than 128 ciphertexts take the two-level reduction (groups of ~sqrt(count) + a shorter tail group)
    for count in (129, 300):
        x = random_cts(rng, n, primes, count)
        got = eng.download(eng.add_many(eng.upload(x)))
        assert np.array_equal(got[0], orc.add_many(x)), count


def test_square_and_relinearize(env):
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 3)
    evk, sizes, dbc = random_evk(rng, n, primes)
    tx = eng.upload(x)
    t3 = eng.square(tx)
    want3 = orc.square(x)
    got3 = eng.download(t3)
    assert np.array_equal(got3, want3)
    k = eng.evk_upload(evk, sizes, dbc)
    got2 = eng.download(eng.relinearize(t3, k))
    assert np.array_equal(got2, orc.relinearize(want3, evk, sizes, dbc))
    got_layer = eng.download(eng.square_layer(eng.upload(x), k))
    assert np.array_equal(got_layer, got2)
    # NTT-form input (what a convolution hands over): the q limbs skip their forward transform inside square
    tn = eng.upload(x)
    eng.to_ntt(tn)
    assert np.array_equal(eng.download(eng.square(tn)), want3)
    assert np.array_equal(eng.download(tn), x), "square must not change the value of its input"


def test_relinearize_both_paths(env):
    """Word-size auxiliary-prime path (default) and the 64-bit path of the reference's own procedure give the oracle's
    bytes, also for extreme operands (keys and c2 at q-1 / 0: largest and smallest integer coefficient sums)."""
    n, primes, t, eng, orc, rng = env
    K = len(primes)
    evk, sizes, dbc = random_evk(rng, n, primes)
    x3 = random_cts(rng, n, primes, 5, size=3)
    top = np.zeros_like(x3[:1])
    for j, q in enumerate(primes):
        top[:, :, j, :n] = q - 1
    x3 = np.concatenate([x3, top, np.zeros_like(top)])
    evk_max = np.zeros_like(evk).reshape(-1, K, n + 1)
    for j, q in enumerate(primes):
        evk_max[:, j, :n] = q - 1
    for keys in (evk, evk_max.ravel()):
        want = orc.relinearize(x3, keys, sizes, dbc)
        k = eng.evk_upload(keys, sizes, dbc)
        try:
            for mode in (1, 0):
                eng.set_relin_mode(mode)
                assert np.array_equal(eng.download(eng.relinearize(eng.upload(x3, size=3), k)), want), mode
        finally:
            eng.set_relin_mode(1)


def test_relinearize_other_digit_width(env):
    """Keys with 20-bit digits do not fit the 16-bit digit planes of the auxiliary-prime path: the engine must fall back to the
    64-bit procedure on its own and still give the oracle's bytes."""
    n, primes, t, eng, orc, rng = env
    evk, sizes, dbc = random_evk(rng, n, primes, dbc=20)
    x3 = random_cts(rng, n, primes, 3, size=3)
    k = eng.evk_upload(evk, sizes, dbc)
    assert np.array_equal(eng.download(eng.relinearize(eng.upload(x3, size=3), k)), orc.relinearize(x3, evk, sizes, dbc))


def test_degree_1024_square_relinearize_transforms():
    """n = 1024 is the smallest degree the kernels are instantiated for (one 5-stage strided pass + the contiguous pass in the
    32-bit transforms, generic key-product kernel): not a CrCNN configuration, but SEAL accepts it."""
    from crcnn_b200.lib import Engine
    n, primes, t = 1024, [0x3fffffff000001], 1 << 10
    eng, orc = Engine(n, primes, t), Oracle(n, primes, t)
    rng = np.random.default_rng(1024)
    x = random_cts(rng, n, primes, 5)
    tx = eng.upload(x)
    eng.to_ntt(tx)
    assert np.array_equal(eng.download(tx, ntt_form=True), orc.ct_transform(x))
    want3 = orc.square(x)
    t3 = eng.square(tx)
    assert np.array_equal(eng.download(t3), want3)
    evk, sizes, dbc = random_evk(rng, n, primes)
    want2 = orc.relinearize(want3, evk, sizes, dbc)
    k = eng.evk_upload(evk, sizes, dbc)
    for mode in (1, 0):
        eng.set_relin_mode(mode)
        assert np.array_equal(eng.download(eng.relinearize(t3, k)), want2), mode
    eng.close()


def _layer_params(orc, rng, count):
    vals = rng.uniform(-1, 1, size=count).astype(np.float32)
    return vals, orc.encode_many(vals)


def test_conv_layer(env):
    n, primes, t, eng, orc, rng = env
    xd, yd, zd, xs, ys, xf, yf, nf = 5, 4, 2, 2, 1, 3, 2, 3
    x = random_cts(rng, n, primes, zd * xd * yd)
    wv, wp = _layer_params(orc, rng, nf * zd * xf * yf)
    bv, bp = _layer_params(orc, rng, nf)
    want = orc.conv(x, xd, yd, zd, xs, ys, xf, yf, nf, wp, bp)
    got = eng.download(eng.conv(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), 1, xd, yd, zd, xs, ys, xf, yf, nf))
    assert np.array_equal(got.reshape(want.shape), want)


def test_conv_layer_batched_and_sharded(env):
    n, primes, t, eng, orc, rng = env
    xd, yd, zd, xs, ys, xf, yf, nf, B = 4, 4, 1, 1, 1, 2, 2, 5, 3
    x = random_cts(rng, n, primes, B * zd * xd * yd)
    wv, wp = _layer_params(orc, rng, nf * zd * xf * yf)
    bv, bp = _layer_params(orc, rng, nf)
    per = zd * xd * yd
    want = np.stack([orc.conv(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, nf, wp, bp) for b in range(B)])
    w, b_ = eng.plain_encode(wv), eng.plain_encode(bv)
    got = eng.download(eng.conv(eng.upload(x), w, b_, B, xd, yd, zd, xs, ys, xf, yf, nf))
    assert np.array_equal(got.reshape(want.shape), want)
    # output-channel shard [1, 4) equals the same channels of the full result
    got_s = eng.download(eng.conv(eng.upload(x), w, b_, B, xd, yd, zd, xs, ys, xf, yf, nf, shard=(1, 3)))
    assert np.array_equal(got_s.reshape((B, 3) + want.shape[2:]), want[:, 1:4])


def test_fc_layer_resident_and_chunked(env):
    n, primes, t, eng, orc, rng = env
    in_dim, out_dim = 11, 13
    x = random_cts(rng, n, primes, in_dim)
    wv, wp = _layer_params(orc, rng, in_dim * out_dim)
    bv, bp = _layer_params(orc, rng, out_dim)
    want = orc.fc(x, in_dim, out_dim, wp, bp)
    got = eng.download(eng.fc(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), 1, in_dim, out_dim))
    assert np.array_equal(got.reshape(want.shape), want)
    # force the chunked path: room for ~3 output rows of NTT-form weights
    eng.set_weight_cache_bytes(3 * in_dim * len(primes) * n * 8)
    try:
        got2 = eng.download(eng.fc(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), 1, in_dim, out_dim))
    finally:
        eng.set_weight_cache_bytes(24 << 30)
    assert np.array_equal(got2.reshape(want.shape), want)


def test_fc_layer_batched(env):
    n, primes, t, eng, orc, rng = env
    in_dim, out_dim, B = 6, 4, 3
    x = random_cts(rng, n, primes, B * in_dim)
    wv, wp = _layer_params(orc, rng, in_dim * out_dim)
    bv, bp = _layer_params(orc, rng, out_dim)
    want = np.stack([orc.fc(x[b * in_dim:(b + 1) * in_dim], in_dim, out_dim, wp, bp) for b in range(B)])
    got = eng.download(eng.fc(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), B, in_dim, out_dim))
    assert np.array_equal(got.reshape(want.shape), want)


@pytest.mark.parametrize("avg", [False, True])
def test_pool_layers(env, avg):
    n, primes, t, eng, orc, rng = env
    # 3x3 windows: 1/9 is not a dyadic value -- the scale factor must be encoded from the DOUBLE (avgPoolingLayer.cpp:10-13)
    for (xd, yd, zd, xs, ys, xf, yf) in [(5, 5, 2, 1, 1, 2, 2), (6, 4, 3, 2, 2, 2, 2), (5, 4, 2, 1, 1, 3, 3), (6, 6, 1, 3, 3, 3, 3)]:
        x = random_cts(rng, n, primes, zd * xd * yd)
        if avg:
            d, cc = orc.encode(1.0 / (xf * yf))
            want = orc.pool(x, xd, yd, zd, xs, ys, xf, yf, d, cc)
            got = eng.download(eng.pool(eng.upload(x), 1, xd, yd, zd, xs, ys, xf, yf, scale=eng.plain_encode_f64([1.0 / (xf * yf)])))
            # coefficient-form input takes the coefficient-domain multiply; NTT-form input the NTT-domain kernel
            tn = eng.upload(x)
            eng.to_ntt(tn)
            got_n = eng.download(eng.pool(tn, 1, xd, yd, zd, xs, ys, xf, yf, scale=eng.plain_encode_f64([1.0 / (xf * yf)])))
            assert np.array_equal(got_n.reshape(want.shape), want)
        else:
            want = orc.pool(x, xd, yd, zd, xs, ys, xf, yf)
            got = eng.download(eng.pool(eng.upload(x), 1, xd, yd, zd, xs, ys, xf, yf))
        assert np.array_equal(got.reshape(want.shape), want)


def test_bn_layer(env):
    n, primes, t, eng, orc, rng = env
    zd, xd, yd = 3, 2, 2
    x = random_cts(rng, n, primes, zd * xd * yd)
    mv, mp = _layer_params(orc, rng, zd)
    vv = rng.uniform(0.5, 3, size=zd).astype(np.float32)
    vp = orc.encode_many(vv)
    want = orc.bn(x, zd, xd, yd, mp, vp)
    got = eng.download(eng.bn(eng.upload(x), 1, zd, xd, yd, eng.plain_encode(mv), eng.plain_encode(vv)))
    assert np.array_equal(got.reshape(want.shape), want)
    tn = eng.upload(x)       # NTT-form input: the NTT-domain kernel instead of the coefficient-domain multiply
    eng.to_ntt(tn)
    got_n = eng.download(eng.bn(tn, 1, zd, xd, yd, eng.plain_encode(mv), eng.plain_encode(vv)))
    assert np.array_equal(got_n.reshape(want.shape), want)
    # factors with an integer part (digits at x^0, x^1, ...) and negative ones
    vv2 = np.array([7.25, -0.3, 26.9], dtype=np.float32)
    want2 = orc.bn(x, zd, xd, yd, mp, orc.encode_many(vv2))
    got2 = eng.download(eng.bn(eng.upload(x), 1, zd, xd, yd, eng.plain_encode(mv), eng.plain_encode(vv2)))
    assert np.array_equal(got2.reshape(want2.shape), want2)


def test_pool_bn_fused(env):
    """crcnn_pool_bn_forward == AvgPoolingLayer::forward then BatchNormLayer::forward (avgPoolingLayer.cpp:16-45, batchNormLayer.cpp:29-40),
    for NTT-form activations (the one-pass kernel) and coefficient-form ones (the two layers in turn), batch 1 and 3, 2x2 and 3x3 windows."""
    n, primes, t, eng, orc, rng = env
    for (xd, yd, zd, xs, ys, xf, yf, batch) in [(4, 4, 3, 2, 2, 2, 2, 1), (6, 6, 2, 3, 3, 3, 3, 3), (5, 4, 5, 1, 1, 2, 2, 2)]:
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        mv, mp = _layer_params(orc, rng, zd)
        vv = rng.uniform(-3, 3, size=zd).astype(np.float32)
        vp = orc.encode_many(vv)
        d, cc = orc.encode(1.0 / (xf * yf))
        x = random_cts(rng, n, primes, batch * zd * xd * yd)
        per = zd * xd * yd
        want = np.concatenate([orc.bn(orc.pool(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, d, cc), zd, xo, yo, mp, vp) for b in range(batch)])
        packs = (eng.plain_encode_f64([1.0 / (xf * yf)]), eng.plain_encode(mv), eng.plain_encode(vv))
        tn = eng.upload(x)
        eng.to_ntt(tn)
        got_n = eng.download(eng.pool_bn(tn, batch, xd, yd, zd, xs, ys, xf, yf, *packs))
        assert np.array_equal(got_n.reshape(want.shape), want)
        again = eng.download(eng.pool_bn(tn, batch, xd, yd, zd, xs, ys, xf, yf, *packs))     # cached per-channel constants
        assert np.array_equal(again.reshape(want.shape), want)
        got_c = eng.download(eng.pool_bn(eng.upload(x), batch, xd, yd, zd, xs, ys, xf, yf, *packs))
        assert np.array_equal(got_c.reshape(want.shape), want)
        # the same scale pack with another batch-norm layer: the cache must not hand back the first layer's constants
        mv2, mp2 = _layer_params(orc, rng, zd)
        want2 = np.concatenate([orc.bn(orc.pool(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, d, cc), zd, xo, yo, mp2, vp) for b in range(batch)])
        got2 = eng.download(eng.pool_bn(tn, batch, xd, yd, zd, xs, ys, xf, yf, packs[0], eng.plain_encode(mv2), packs[2]))
        assert np.array_equal(got2.reshape(want2.shape), want2)
        # PoolingLayer (window sum, no factor -- the WoPad topology, cnnBuilder.cpp:136-155) followed by the batch-norm
        want3 = np.concatenate([orc.bn(orc.pool(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf), zd, xo, yo, mp, vp) for b in range(batch)])
        got3 = eng.download(eng.pool_bn(tn, batch, xd, yd, zd, xs, ys, xf, yf, None, packs[1], packs[2]))
        assert np.array_equal(got3.reshape(want3.shape), want3)
        got3c = eng.download(eng.pool_bn(eng.upload(x), batch, xd, yd, zd, xs, ys, xf, yf, None, packs[1], packs[2]))
        assert np.array_equal(got3c.reshape(want3.shape), want3)
        assert np.array_equal(eng.download(eng.pool_bn(tn, batch, xd, yd, zd, xs, ys, xf, yf, *packs)).reshape(want.shape), want)   # and back to the scaled constants


def test_conv_pool_bn_on_the_pooled_grid(env):
    """crcnn_conv_pool_bn_forward == ConvolutionalLayer, AvgPoolingLayer, BatchNormLayer::forward in a row (convolutionalLayer.cpp:136-196,
    avgPoolingLayer.cpp:16-45, batchNormLayer.cpp:29-40).  Stride-1 convolutions are evaluated on the pooled grid (window sums of the input,
    convolution at the pooling stride, bias once per window element); the rest take the layer-by-layer path.  Cases: the headline block's
    shape in small (5x5 filter, 2x2/2 pooling), 3x3/3 and overlapping 3x3/2 windows, several input channels, batch > 1, a strided convolution."""
    n, primes, t, eng, orc, rng = env
    cases = [  # xd yd zd  xs ys xf yf nf  pxs pys pxf pyf  batch  pooled grid?
        (8, 8, 1, 1, 1, 5, 5, 3, 2, 2, 2, 2, 2, True),
        (7, 8, 2, 1, 1, 2, 3, 2, 3, 3, 3, 3, 1, False),     # pooling stride 3 > filter 2: outputs the strided convolution would skip
        (8, 7, 2, 1, 1, 2, 2, 3, 2, 2, 3, 3, 2, True),      # overlapping windows
        (9, 9, 1, 2, 2, 5, 5, 2, 1, 1, 2, 2, 2, True),      # the headline block: 5x5 stride 2, then 2x2 stride 1 (cnnBuilder.cpp:115-117)
        (8, 9, 2, 2, 1, 3, 2, 2, 1, 2, 2, 2, 1, True),      # different strides per axis
        (9, 9, 1, 2, 2, 3, 3, 2, 2, 2, 2, 2, 1, False),     # combined stride 4 > filter 3
    ]
    for (xd, yd, zd, xs, ys, xf, yf, nf, pxs, pys, pxf, pyf, batch, pooled_grid) in cases:
        cxo, cyo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        pxo, pyo = (cxo - pxf) // pxs + 1, (cyo - pyf) // pys + 1
        wv, wp = _layer_params(orc, rng, nf * zd * xf * yf)
        bv, bp = _layer_params(orc, rng, nf)
        mv, mp = _layer_params(orc, rng, nf)
        vv = rng.uniform(-3, 3, size=nf).astype(np.float32)
        vp = orc.encode_many(vv)
        d, cc = orc.encode(1.0 / (pxf * pyf))
        per = zd * xd * yd
        x = random_cts(rng, n, primes, batch * per)
        want = np.concatenate([
            orc.bn(orc.pool(orc.conv(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, nf, wp, bp).reshape(nf * cxo * cyo, *x.shape[1:]),
                            cxo, cyo, nf, pxs, pys, pxf, pyf, d, cc), nf, pxo, pyo, mp, vp) for b in range(batch)])
        w, bias = eng.plain_encode(wv), eng.plain_encode(bv)
        packs = (eng.plain_encode_f64([1.0 / (pxf * pyf)]), eng.plain_encode(mv), eng.plain_encode(vv))
        geo = (batch, xd, yd, zd, xs, ys, xf, yf, nf, pxs, pys, pxf, pyf)
        def run():
            eng.prof_reset(); eng.prof_enable(True)
            y = eng.download(eng.conv_pool_bn(eng.upload(x), w, bias, *geo, *packs))
            work = eng.prof_work()
            eng.prof_enable(False)
            return y, work
        got, work = run()
        assert np.array_equal(got.reshape(want.shape), want), geo
        # which path ran: CRCNN_NO_POOLED_CONV forces the layer-by-layer path, whose weighted sum has more columns (other algorithmic work)
        # (ternary-tap GEMM off for the comparison, so that both paths count their work in the same kernel family)
        eng.set_tensor_core_mode(0)
        os.environ["CRCNN_NO_POOLED_CONV"] = "1"
        try:
            got_l, work_l = run()
            del os.environ["CRCNN_NO_POOLED_CONV"]
            got, work = run()
        finally:
            os.environ.pop("CRCNN_NO_POOLED_CONV", None)
            eng.set_tensor_core_mode(1)
        assert np.array_equal(got_l.reshape(want.shape), want) and np.array_equal(got.reshape(want.shape), want), geo
        macs = lambda wk: sum(v[1] for k, v in wk.items() if k.startswith("weighted_sum"))
        assert (macs(work) < macs(work_l)) == pooled_grid, (geo, work, work_l)
        tn = eng.upload(x)
        eng.to_ntt(tn)      # NTT-form input
        got_n = eng.download(eng.conv_pool_bn(tn, w, bias, *geo, *packs))
        assert np.array_equal(got_n.reshape(want.shape), want), geo
        # output-channel shard [1, nf) (ShardedNetwork): the same channels of the full result; only on the pooled grid
        if pooled_grid:
            got_sh = eng.download(eng.conv_pool_bn(eng.upload(x), w, bias, *geo, *packs, shard=(1, nf - 1)))
            want_sh = want.reshape(batch, nf, pxo * pyo, *x.shape[1:])[:, 1:]
            assert np.array_equal(got_sh.reshape(want_sh.shape), want_sh), geo
        else:
            with pytest.raises(Exception):
                eng.conv_pool_bn(eng.upload(x), w, bias, *geo, *packs, shard=(1, nf - 1))
        # sum pooling (no factor) in the middle: same path, C = invstd
        want_s = np.concatenate([
            orc.bn(orc.pool(orc.conv(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, nf, wp, bp).reshape(nf * cxo * cyo, *x.shape[1:]),
                            cxo, cyo, nf, pxs, pys, pxf, pyf), nf, pxo, pyo, mp, vp) for b in range(batch)])
        got_s = eng.download(eng.conv_pool_bn(eng.upload(x), w, bias, *geo, None, packs[1], packs[2]))
        assert np.array_equal(got_s.reshape(want_s.shape), want_s), geo
        assert np.array_equal(eng.download(eng.conv_pool_bn(eng.upload(x), w, bias, *geo, *packs)).reshape(want.shape), want), geo
        # the bias pack still adds ONE bias in a plain convolution afterwards (the repeated form lives in a derived pack)
        want_c = np.concatenate([orc.conv(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, nf, wp, bp).reshape(nf * cxo * cyo, *x.shape[1:]) for b in range(batch)])
        got_c = eng.download(eng.conv(eng.upload(x), w, bias, batch, xd, yd, zd, xs, ys, xf, yf, nf))
        assert np.array_equal(got_c.reshape(want_c.shape), want_c), geo


def test_two_fc_layers_as_one_composed_layer(env):
    """crcnn_fc_fc_forward == FullyConnectedLayer::forward twice (fullyConnectedLayer.cpp:96-166; fc3 -> fc4, cnnBuilder.cpp:121-122): the
    composed weights W2 W1 and bias W2 Delta b1 + Delta b2 are built on the device; odd fan-in (a last column without a partner), several
    build passes with a partial last one, batch > 1, and a shape where composing does not pay (falls back to the two layers)."""
    n, primes, t, eng, orc, rng = env
    cases = [  # in mid out batch  build pairs per pass  composed?
        (21, 12, 3, 2, 4, True),
        (16, 9, 2, 1, None, True),
        (6, 4, 40, 1, None, False),      # 40 x 6 terms composed vs 4 x 46 layer by layer: not smaller
    ]
    for (in_dim, mid, out, batch, pairs, composed) in cases:
        w1v, w1p = _layer_params(orc, rng, mid * in_dim)
        b1v, b1p = _layer_params(orc, rng, mid)
        w2v, w2p = _layer_params(orc, rng, out * mid)
        b2v, b2p = _layer_params(orc, rng, out)
        x = random_cts(rng, n, primes, batch * in_dim)
        want = np.concatenate([orc.fc(orc.fc(x[b * in_dim:(b + 1) * in_dim], in_dim, mid, w1p, b1p).reshape(mid, *x.shape[1:]), mid, out, w2p, b2p).reshape(out, *x.shape[1:])
                               for b in range(batch)])
        packs = [eng.plain_encode(v) for v in (w1v, b1v, w2v, b2v)]

        def run():
            eng.prof_reset(); eng.prof_enable(True)
            y = eng.download(eng.fc_fc(eng.upload(x), *packs, batch, in_dim, mid, out))
            work = eng.prof_work()
            eng.prof_enable(False)
            return y, work
        eng.set_tensor_core_mode(0)      # both paths count their work in the same kernel family
        if pairs:
            os.environ["CRCNN_FC_COMPOSE_PAIRS"] = str(pairs)
        try:
            got, _ = run()               # builds the composed layer
            got2, work = run()           # steady state: the composed layer alone
            os.environ["CRCNN_NO_FC_COMPOSE"] = "1"
            got_l, work_l = run()
        finally:
            os.environ.pop("CRCNN_NO_FC_COMPOSE", None)
            os.environ.pop("CRCNN_FC_COMPOSE_PAIRS", None)
            eng.set_tensor_core_mode(1)
        for y in (got, got2, got_l):
            assert np.array_equal(y.reshape(want.shape), want), (in_dim, mid, out, batch)
        macs = lambda wk: sum(v[1] for k, v in wk.items() if k.startswith("weighted_sum"))
        assert (macs(work) < macs(work_l)) == composed, (work, work_l)
        # NTT-form input, and the first layer on its own still gives its own output afterwards
        tn = eng.upload(x)
        eng.to_ntt(tn)
        assert np.array_equal(eng.download(eng.fc_fc(tn, *packs, batch, in_dim, mid, out)).reshape(want.shape), want)
        want1 = np.concatenate([orc.fc(x[b * in_dim:(b + 1) * in_dim], in_dim, mid, w1p, b1p).reshape(mid, *x.shape[1:]) for b in range(batch)])
        assert np.array_equal(eng.download(eng.fc(eng.upload(x), packs[0], packs[1], batch, in_dim, mid)).reshape(want1.shape), want1)


def test_pool_bn_fc_fc_as_window_sums_and_one_composed_layer(env):
    """crcnn_pool_bn_fc_fc_forward == AvgPoolingLayer, BatchNormLayer, FullyConnectedLayer x2 ::forward in a row (cnnBuilder.cpp:118-122),
    for coefficient-form (after a square layer) and NTT-form activations, batch > 1, 2x2 stride-1 and 3x3 stride-2 windows."""
    n, primes, t, eng, orc, rng = env
    for (xd, yd, zd, pxs, pys, pxf, pyf, mid, out, batch) in [(4, 4, 3, 1, 1, 2, 2, 8, 3, 2), (5, 7, 2, 2, 2, 3, 3, 5, 2, 1)]:
        pxo, pyo = (xd - pxf) // pxs + 1, (yd - pyf) // pys + 1
        in_dim = zd * pxo * pyo
        mv, mp = _layer_params(orc, rng, zd)
        vv = rng.uniform(-3, 3, size=zd).astype(np.float32)
        vp = orc.encode_many(vv)
        d, cc = orc.encode(1.0 / (pxf * pyf))
        w1v, w1p = _layer_params(orc, rng, mid * in_dim)
        b1v, b1p = _layer_params(orc, rng, mid)
        w2v, w2p = _layer_params(orc, rng, out * mid)
        b2v, b2p = _layer_params(orc, rng, out)
        per = zd * xd * yd
        x = random_cts(rng, n, primes, batch * per)
        ct = x.shape[1:]

        def ref_one(xi):
            y = orc.bn(orc.pool(xi, xd, yd, zd, pxs, pys, pxf, pyf, d, cc), zd, pxo, pyo, mp, vp).reshape(in_dim, *ct)
            return orc.fc(orc.fc(y, in_dim, mid, w1p, b1p).reshape(mid, *ct), mid, out, w2p, b2p).reshape(out, *ct)
        want = np.concatenate([ref_one(x[b * per:(b + 1) * per]) for b in range(batch)])
        packs = [eng.plain_encode_f64([1.0 / (pxf * pyf)]), eng.plain_encode(mv), eng.plain_encode(vv)] + [eng.plain_encode(v) for v in (w1v, b1v, w2v, b2v)]
        geo = (batch, xd, yd, zd, pxs, pys, pxf, pyf)

        def run(t_in):
            eng.prof_reset(); eng.prof_enable(True)
            y = eng.download(eng.pool_bn_fc_fc(t_in, *geo, *packs, mid, out))
            work = eng.prof_work()
            eng.prof_enable(False)
            return y, work
        eng.set_tensor_core_mode(0)
        try:
            got, _ = run(eng.upload(x))
            got2, work = run(eng.upload(x))
            tn = eng.upload(x)
            eng.to_ntt(tn)
            got_n, _ = run(tn)
            os.environ["CRCNN_NO_FC_COMPOSE"] = "1"
            got_l, work_l = run(eng.upload(x))
        finally:
            os.environ.pop("CRCNN_NO_FC_COMPOSE", None)
            eng.set_tensor_core_mode(1)
        for y in (got, got2, got_n, got_l):
            assert np.array_equal(y.reshape(want.shape), want), geo
        macs = lambda wk: sum(v[1] for k, v in wk.items() if k.startswith("weighted_sum"))
        assert macs(work) < macs(work_l), (work, work_l)
        def ref_sum(xi):
            y = orc.bn(orc.pool(xi, xd, yd, zd, pxs, pys, pxf, pyf), zd, pxo, pyo, mp, vp).reshape(in_dim, *ct)
            return orc.fc(orc.fc(y, in_dim, mid, w1p, b1p).reshape(mid, *ct), mid, out, w2p, b2p).reshape(out, *ct)
        want_s = np.concatenate([ref_sum(x[b * per:(b + 1) * per]) for b in range(batch)])
        got_s = eng.download(eng.pool_bn_fc_fc(eng.upload(x), *geo, None, *packs[1:], mid, out))
        assert np.array_equal(got_s.reshape(want_s.shape), want_s), geo
        # the composed layer with the affine map must not be mistaken for the plain composition of the same two layers (and back)
        y_in = np.concatenate([orc.bn(orc.pool(x[b * per:(b + 1) * per], xd, yd, zd, pxs, pys, pxf, pyf, d, cc), zd, pxo, pyo, mp, vp).reshape(in_dim, *ct) for b in range(batch)])
        assert np.array_equal(eng.download(eng.fc_fc(eng.upload(y_in), *packs[3:], batch, in_dim, mid, out)).reshape(want.shape), want)
        assert np.array_equal(eng.download(eng.pool_bn_fc_fc(eng.upload(x), *geo, *packs, mid, out)).reshape(want.shape), want)


def test_layer_chain_stays_exact(env):
    """conv -> avgpool -> bn -> square -> fc without leaving the device (lazy NTT domain inside)."""
    n, primes, t, eng, orc, rng = env
    x = random_cts(rng, n, primes, 16)  # 1x4x4
    evk, sizes, dbc = random_evk(rng, n, primes)
    wv, wp = _layer_params(orc, rng, 2 * 4)
    bv, bp = _layer_params(orc, rng, 2)
    mv, mp = _layer_params(orc, rng, 2)
    vv, vp = _layer_params(orc, rng, 2)
    fv, fp = _layer_params(orc, rng, 3 * 8)
    fb, fbp = _layer_params(orc, rng, 3)
    d, cc = orc.encode(0.25)
    o = orc.conv(x, 4, 4, 1, 1, 1, 2, 2, 2, wp, bp)              # 2x3x3
    o = orc.pool(o, 3, 3, 2, 1, 1, 2, 2, d, cc)                  # 2x2x2
    o = orc.bn(o, 2, 2, 2, mp, vp)
    o = orc.square_layer(o, evk, sizes, dbc)
    want = orc.fc(o, 8, 3, fp, fbp)
    g = eng.conv(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), 1, 4, 4, 1, 1, 1, 2, 2, 2)
    g = eng.pool(g, 1, 3, 3, 2, 1, 1, 2, 2, scale=eng.plain_encode([0.25]))
    g = eng.bn(g, 1, 2, 2, 2, eng.plain_encode(mv), eng.plain_encode(vv))
    g = eng.square_layer(g, eng.evk_upload(evk, sizes, dbc))
    g = eng.fc(g, eng.plain_encode(fv), eng.plain_encode(fb), 1, 8, 3)
    assert np.array_equal(eng.download(g).reshape(want.shape), want)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_end_to_end_with_real_seal_client():
    """Real SEAL keys/encryption on the host (client boundary), GPU forward, SEAL decryption."""
    from oracle.ref import Ref
    from crcnn_b200.lib import Engine
    n, t = 4096, 1 << 18
    r = Ref(n, t, seed=11)
    eng = Engine(n, r.primes, t)
    rng = np.random.default_rng(5)
    img = rng.uniform(-0.4242, 2.8215, size=16).astype(np.float32)
    x = r.encrypt(img)
    wv = rng.uniform(-1, 1, size=2 * 4).astype(np.float32)
    bv = rng.uniform(-1, 1, size=2).astype(np.float32)
    evk, sizes, dbc = r.evk()
    want = r.square_layer(r.conv(x, 4, 4, 1, 1, 1, 2, 2, 2, wv, bv), 2, 3, 3)
    g = eng.conv(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), 1, 4, 4, 1, 1, 1, 2, 2, 2)
    g = eng.square_layer(g, eng.evk_upload(evk, sizes, dbc))
    got = eng.download(g)
    assert np.array_equal(got.reshape(want.shape), want)
    vals, budgets = r.decrypt(got)
    ref_vals, ref_budgets = r.decrypt(want)
    assert np.array_equal(vals, ref_vals) and np.array_equal(budgets, ref_budgets)
    # and the decrypted numbers are the plain computation
    plain = np.zeros((2, 3, 3))
    im = img.reshape(4, 4)
    for k in range(2):
        for i in range(3):
            for j in range(3):
                plain[k, i, j] = (im[i:i + 2, j:j + 2].ravel() * wv[k * 4:(k + 1) * 4]).sum() + bv[k]
    assert np.allclose(vals.reshape(2, 3, 3), plain ** 2, atol=1e-3)
    eng.close()


def test_empty_shards_are_legal(env):
    """More ranks than outputs (fc4 has 10 rows, a box has up to 16 GPUs): a rank's range may be empty; the shard calls then return
    an empty tensor instead of an error, and the layer's kernel choice does not depend on the shard width."""
    n, primes, t, eng, orc, rng = env
    x = eng.upload(random_cts(rng, n, primes, 2 * 3 * 3))
    wv = rng.uniform(-1, 1, size=2 * 2 * 2 * 2).astype(np.float32)
    bv = rng.uniform(-1, 1, size=2).astype(np.float32)
    e = eng.conv(x, eng.plain_encode(wv), eng.plain_encode(bv), 1, 3, 3, 2, 1, 1, 2, 2, 2, shard=(2, 0))
    assert eng.count(e) == 0 and eng.download(e).shape[0] == 0
    fw = rng.uniform(-1, 1, size=18 * 3).astype(np.float32)
    fb = rng.uniform(-1, 1, size=3).astype(np.float32)
    e = eng.fc(x, eng.plain_encode(fw), eng.plain_encode(fb), 1, 18, 3, shard=(1, 0))
    assert eng.count(e) == 0
    # the non-empty shards of the same layer still give the layer's bytes
    want = orc.fc(eng.download(x), 18, 3, orc.encode_many(fw), orc.encode_many(fb)).reshape(3, 2, len(primes), n + 1)
    got = np.concatenate([eng.download(eng.fc(x, eng.plain_encode(fw), eng.plain_encode(fb), 1, 18, 3, shard=s))
                          for s in ((0, 1), (1, 0), (1, 2))])
    assert np.array_equal(got, want)
