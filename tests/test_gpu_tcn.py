"""NTT-domain limb-split tensor-core weighted sum (crcnn_b200/csrc/tcn_mac.cuh) against the CPU oracle, byte for byte.

This is the default kernel of every conv / fc layer that the ternary-tap kernel does not take: weights of ANY value
(also |w| >= 1/2), any fan-in up to 4096.  Every test checks through the kernel-class counters that the tcgen05
kernel is what ran, and that the CUDA-core kernel (limb-split mode off) gives the same bytes.
"""
import numpy as np
import pytest

from util import PRIMES, T_FOR_N, random_cts
from oracle.port import Oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[2048, 4096, 8192])
def env(request):
    from crcnn_b200.lib import Engine
    n = request.param
    primes, t = PRIMES[n], T_FOR_N[n]
    eng = Engine(n, primes, t)
    eng.set_tensor_core_mode(0)       # keep the ternary-tap kernel out of the way
    eng.set_limb_split_mode(1)
    orc = Oracle(n, primes, t)
    rng = np.random.default_rng(11 * n)
    yield n, primes, t, eng, orc, rng
    eng.close()


def _vals(orc, rng, count, lo=-3.0, hi=3.0):
    vals = rng.uniform(lo, hi, size=count).astype(np.float32)
    return vals, orc.encode_many(vals)


def _explain(got, want):
    bad = np.argwhere(got != want)
    if len(bad) == 0:
        return "equal"
    cts = sorted(set(int(b[0]) for b in bad))
    first = tuple(int(v) for v in bad[0])
    return "%d mismatching words in %d of %d ciphertexts (first cts %s); first at %s: got %d want %d" % (
        len(bad), len(cts), got.shape[0], cts[:8], first, int(got[first]), int(want[first]))


def _launches(eng):
    return eng.prof().get("weighted_sum_tcn_i8", (0, 0.0))[0]


def _cuda_core(eng, fn):
    eng.set_limb_split_mode(0)
    try:
        return fn()
    finally:
        eng.set_limb_split_mode(1)


@pytest.mark.parametrize("in_dim,out_dim", [(5, 3), (25, 20), (32, 7), (33, 4), (70, 9), (130, 66), (300, 2)])
def test_fc_layer_limb_split(env, in_dim, out_dim):
    """fan-in <= 32 takes the 32-byte-row variant, larger ones 128-byte K blocks (partial last block, 1-3 blocks);
    66 outputs need two row tiles."""
    n, primes, t, eng, orc, rng = env
    if (n == 8192 and in_dim * out_dim > 700) or (n == 4096 and in_dim * out_dim > 5000):
        pytest.skip("oracle time")
    x = random_cts(rng, n, primes, in_dim)
    wv, wp = _vals(orc, rng, in_dim * out_dim)
    bv, bp = _vals(orc, rng, out_dim)
    want = orc.fc(x, in_dim, out_dim, wp, bp)
    w, b_ = eng.plain_encode(wv), eng.plain_encode(bv)
    before = _launches(eng)
    got = eng.download(eng.fc(eng.upload(x), w, b_, 1, in_dim, out_dim)).reshape(want.shape)
    assert _launches(eng) > before, "the limb-split tensor-core kernel did not run"
    assert np.array_equal(got, want), _explain(got, want)
    got2 = _cuda_core(eng, lambda: eng.download(eng.fc(eng.upload(x), eng.plain_encode(wv), b_, 1, in_dim, out_dim)))
    assert np.array_equal(got2.reshape(want.shape), want)


def test_conv_layer_limb_split_batched_sharded_chunked(env):
    n, primes, t, eng, orc, rng = env
    xd, yd, zd, xs, ys, xf, yf, nf, B = 6, 5, 3, 2, 1, 3, 3, 5, 3   # fan-in 27, 2x3 positions, 18 columns per image pair
    per = zd * xd * yd
    x = random_cts(rng, n, primes, B * per)
    wv, wp = _vals(orc, rng, nf * zd * xf * yf)
    bv, bp = _vals(orc, rng, nf)
    want = np.stack([orc.conv(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, nf, wp, bp) for b in range(B)])
    w, b_ = eng.plain_encode(wv), eng.plain_encode(bv)
    before = _launches(eng)
    got = eng.download(eng.conv(eng.upload(x), w, b_, B, xd, yd, zd, xs, ys, xf, yf, nf)).reshape(want.shape)
    assert _launches(eng) == before + 1
    assert np.array_equal(got, want), _explain(got, want)
    # NTT-form input, tiny scratch budget: 32 slots per launch
    tx = eng.upload(x)
    eng.to_ntt(tx)
    eng.set_tensor_core_mode(0, 0, 1)
    try:
        got2 = eng.download(eng.conv(tx, w, b_, B, xd, yd, zd, xs, ys, xf, yf, nf)).reshape(want.shape)
    finally:
        eng.set_tensor_core_mode(0, 0, 12 << 30)
    assert _launches(eng) == before + 1 + len(primes) * n // 32
    assert np.array_equal(got2, want), _explain(got2, want)
    # output-channel shard
    want_s = want.reshape(B, nf, -1)[:, 2:5]
    got_s = eng.download(eng.conv(eng.upload(x), w, b_, B, xd, yd, zd, xs, ys, xf, yf, nf, shard=(2, 3))).reshape(want_s.shape)
    assert np.array_equal(got_s, want_s), _explain(got_s, want_s)


@pytest.mark.parametrize("zd,nf,variant", [(1, 4, 3), (8, 3, 3), (20, 5, 3), (20, 40, 3), (8, 3, 2), (20, 70, 2)])
def test_conv_layer_both_kernel_shapes(env, zd, nf, variant):
    """3x3 filters over zd channels: fan-in 9 (32-byte rows), 72 (one 128-byte K block), 180 (two blocks, the second
    partial -- conv2 of PlainModel.h5); 40 outputs need two 32-output tiles in the column-major kernel, 70 two 64-output
    tiles in the row-major one.  variant 3 forces the column-major kernel, 2 the row-major one."""
    n, primes, t, eng, orc, rng = env
    if (n >= 4096 and zd * nf > 100) or (n == 8192 and zd * nf > 30):
        pytest.skip("oracle time")
    xd, yd, xs, ys, xf, yf, B = 7, 5, 2, 1, 3, 3, 2     # 3x3 positions, 36 columns for the two images
    per = zd * xd * yd
    x = random_cts(rng, n, primes, B * per)
    wv, wp = _vals(orc, rng, nf * zd * xf * yf)
    bv, bp = _vals(orc, rng, nf)
    want = np.stack([orc.conv(x[b * per:(b + 1) * per], xd, yd, zd, xs, ys, xf, yf, nf, wp, bp) for b in range(B)])
    eng.set_limb_split_mode(variant)
    try:
        before = _launches(eng)
        got = eng.download(eng.conv(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), B, xd, yd, zd, xs, ys, xf, yf, nf))
    finally:
        eng.set_limb_split_mode(1)
    assert _launches(eng) == before + 1
    got = got.reshape(want.shape)
    assert np.array_equal(got, want), _explain(got, want)


def test_extreme_residues_limb_split(env):
    """All-(q-1) inputs against weights whose NTT values are arbitrary: the largest plane sums (every byte of every
    input 0xff in the top planes), fan-in 128."""
    n, primes, t, eng, orc, rng = env
    in_dim, out_dim = 128, 3
    K = len(primes)
    x = np.zeros((in_dim, 2, K, n + 1), dtype=np.uint64)
    for j, q in enumerate(primes):
        x[:, :, j, :n] = q - 1
    x[::5, :, :, :n] = 1
    wv, wp = _vals(orc, rng, in_dim * out_dim, -40.0, 40.0)
    bv, bp = _vals(orc, rng, out_dim)
    want = orc.fc(x, in_dim, out_dim, wp, bp)
    tx = eng.upload(x, ntt_form=True)     # the same words taken as NTT-form data: slot values q-1 everywhere
    want_ntt = None
    before = _launches(eng)
    got = eng.download(eng.fc(eng.upload(x), eng.plain_encode(wv), eng.plain_encode(bv), 1, in_dim, out_dim)).reshape(want.shape)
    assert _launches(eng) == before + 1
    assert np.array_equal(got, want), _explain(got, want)
    # NTT-form all-(q-1) slots: compare the two GPU kernels with each other in the NTT domain
    w, b_ = eng.plain_encode(wv), eng.plain_encode(bv)
    g1 = eng.download(eng.fc(tx, w, b_, 1, in_dim, out_dim), ntt_form=True)
    g2 = _cuda_core(eng, lambda: eng.download(eng.fc(tx, eng.plain_encode(wv), b_, 1, in_dim, out_dim), ntt_form=True))
    assert np.array_equal(g1, g2), _explain(g1, g2)


@pytest.mark.parametrize("variant", [2, 3])
def test_both_reductions_of_the_class_sums(env, variant):
    """The folded reduction through q = 2^k - delta (default) and the 128-bit recombination + Barrett give the oracle's bytes,
    in both kernel shapes; random and all-(q-1) NTT-form slots (the largest class sums), with and without a bias in play
    (polynomial 1 never takes one)."""
    n, primes, t, eng, orc, rng = env
    in_dim, out_dim = 40, 5
    K = len(primes)
    x = random_cts(rng, n, primes, in_dim)
    wv, wp = _vals(orc, rng, in_dim * out_dim, -30.0, 30.0)
    bv, bp = _vals(orc, rng, out_dim)
    want = orc.fc(x, in_dim, out_dim, wp, bp)
    xm = np.zeros((in_dim, 2, K, n + 1), dtype=np.uint64)
    for j, q in enumerate(primes):
        xm[:, :, j, :n] = q - 1
    w, b_ = eng.plain_encode(wv), eng.plain_encode(bv)
    outs = {}
    eng.set_limb_split_mode(variant)
    try:
        for mode in (1, 0):
            eng.set_limb_split_reduction(mode)
            before = _launches(eng)
            got = eng.download(eng.fc(eng.upload(x), w, b_, 1, in_dim, out_dim)).reshape(want.shape)
            assert _launches(eng) == before + 1
            assert np.array_equal(got, want), "reduction %d: %s" % (mode, _explain(got, want))
            outs[mode] = eng.download(eng.fc(eng.upload(xm, ntt_form=True), w, b_, 1, in_dim, out_dim), ntt_form=True)
    finally:
        eng.set_limb_split_reduction(1)
        eng.set_limb_split_mode(1)
    assert np.array_equal(outs[0], outs[1]), _explain(outs[1], outs[0])
    g2 = _cuda_core(eng, lambda: eng.download(eng.fc(eng.upload(xm, ntt_form=True), eng.plain_encode(wv), b_, 1, in_dim, out_dim), ntt_form=True))
    assert np.array_equal(outs[1], g2), _explain(outs[1], g2)
