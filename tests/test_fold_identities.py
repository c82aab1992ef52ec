"""CPU: the algebra behind the composed layers, stated with the oracle's own evaluator primitives and exact integer arithmetic.

crcnn_conv_pool_bn_forward and crcnn_fc_fc_forward / crcnn_pool_bn_fc_fc_forward do not evaluate the reference's layers one by one: they evaluate ONE
weighted sum whose constants were folded in the NTT domain.  The GPU tests compare those entry points with the oracle's layer-by-layer result; this file
checks, without a GPU, that the folded FORMULAS themselves give the reference's bytes -- every product and sum below is taken mod q_j on Python integers,
the only oracle calls are its transforms, multiply_plain_ntt and the layer-by-layer chains being matched (which are pinned to the compiled reference by
tests/test_oracle_vs_reference.py)."""
import numpy as np
import pytest

from util import PRIMES, T_FOR_N, random_cts
from oracle.port import Oracle

N = 2048


@pytest.fixture(scope="module")
def env():
    primes, t = PRIMES[N], T_FOR_N[N]
    return primes, t, Oracle(N, primes, t), np.random.default_rng(77)


def _obj(a):
    return np.asarray(a, dtype=np.uint64).astype(object)


def _mod_rows(a, primes):
    """reduce an object array [..., K, stride] limb by limb"""
    out = a.copy()
    for j, q in enumerate(primes):
        out[..., j, :] = out[..., j, :] % int(q)
    return out


def _ntt_plain(orc, plain):
    return _obj(orc.plain_to_ntt(plain))                                   # [K][stride], multiplicative lift (t-1 -> q-1 ...)


def _ntt_delta_plain(orc, plain, primes):
    """NTT form of what add_plain adds to polynomial 0 (Delta-scaled plaintext): add it to a zero ciphertext and transform that."""
    zero = np.zeros((1, 2, len(primes), N + 1), dtype=np.uint64)
    return _obj(orc.ct_transform(orc.plain_op(zero, plain, "add")))[0, 0]  # [K][stride]


def _weighted_sum_ntt(orc, primes, x_cts, weights_ntt, bias_ntt):
    """INTT( sum_r NTT(x_r) (.) W_r + bias on polynomial 0 ): x_cts [R][2][K][stride], weights_ntt [R][K][stride] (object), bias_ntt [K][stride] (object)"""
    xn = _obj(orc.ct_transform(x_cts))
    acc = np.zeros(xn.shape[1:], dtype=object)
    for r in range(xn.shape[0]):
        acc = acc + xn[r] * weights_ntt[r][None]
    acc[0] = acc[0] + bias_ntt
    acc = _mod_rows(acc, primes)
    acc[..., N] = 0
    return orc.ct_transform(acc.astype(np.uint64)[None], inverse=True)[0]


def test_conv_pool_bn_equals_one_weighted_sum_on_the_pooled_grid(env):
    primes, t, orc, rng = env
    xd = yd = 6; zd = 2; xs = ys = 2; xf = yf = 2; nf = 2          # conv 2x2 stride 2 -> 3x3
    pxs = pys = 1; pxf = pyf = 2                                    # avg-pool 2x2 stride 1 -> 2x2 (the headline block in small)
    cxo = (xd - xf) // xs + 1
    pxo = (cxo - pxf) // pxs + 1
    Rp, R = pxf * pyf, zd * xf * yf
    f = lambda k: rng.uniform(-1, 1, size=k).astype(np.float32)
    wv, bv, mv, vv = f(nf * R), f(nf), f(nf), rng.uniform(-3, 3, size=nf).astype(np.float32)
    wp, bp, mp, vp = (orc.encode_many(v) for v in (wv, bv, mv, vv))
    d, cc = orc.encode(1.0 / Rp)
    x = random_cts(rng, N, primes, zd * xd * yd)
    want = orc.bn(orc.pool(orc.conv(x, xd, yd, zd, xs, ys, xf, yf, nf, wp, bp).reshape(nf * cxo * cxo, *x.shape[1:]), cxo, cxo, nf, pxs, pys, pxf, pyf, d, cc),
                  nf, pxo, pxo, mp, vp)                                # [nf][pxo][pxo][2][K][stride]
    # constants: C_k = s (.) v_k ; D_k = Delta m_k (.) v_k ; W'[k,r] = W[k,r] (.) C_k ; B'_k = Rp * Delta B_k (.) C_k - D_k      (all mod q_j)
    s_ntt = _ntt_plain(orc, d[:cc] if cc else d)
    xi = x.reshape(zd, xd, yd, *x.shape[1:])
    for k in range(nf):
        v_ntt = _ntt_plain(orc, vp[k])
        C = _mod_rows(s_ntt * v_ntt, primes)
        D = _mod_rows(_ntt_delta_plain(orc, mp[k], primes) * v_ntt, primes)
        Wf = [_mod_rows(_ntt_plain(orc, wp[k * R + r]) * C, primes) for r in range(R)]
        Bf = _mod_rows(Rp * _ntt_delta_plain(orc, bp[k], primes) * C - D, primes)
        for i in range(pxo):
            for j in range(pxo):
                # S[z, u, v] = sum over the pooling window of X[z, u + a*cs, v + b*cs]; the convolution reads it at stride ps*cs
                cols = []
                for z in range(zd):
                    for kx in range(xf):
                        for ky in range(yf):
                            u, v = i * pxs * xs + kx, j * pys * ys + ky
                            win = np.stack([xi[z, u + a * xs, v + b * ys] for a in range(pxf) for b in range(pyf)])
                            cols.append(orc.add_many(win))
                got = _weighted_sum_ntt(orc, primes, np.stack(cols), Wf, Bf)
                assert np.array_equal(got, want[k, i, j]), (k, i, j)


def test_two_fc_layers_equal_one_composed_layer(env):
    primes, t, orc, rng = env
    in_dim, mid, out = 5, 4, 2
    f = lambda k: rng.uniform(-1, 1, size=k).astype(np.float32)
    w1p, b1p, w2p, b2p = (orc.encode_many(f(k)) for k in (mid * in_dim, mid, out * mid, out))
    x = random_cts(rng, N, primes, in_dim)
    want = orc.fc(orc.fc(x, in_dim, mid, w1p, b1p).reshape(mid, *x.shape[1:]), mid, out, w2p, b2p).reshape(out, *x.shape[1:])
    W1 = [[_ntt_plain(orc, w1p[o * in_dim + r]) for r in range(in_dim)] for o in range(mid)]
    for k in range(out):
        W2 = [_ntt_plain(orc, w2p[k * mid + o]) for o in range(mid)]
        # W[k,r] = sum_o W2[k,o] (.) W1[o,r] ;  B_k = sum_o W2[k,o] (.) Delta b1_o + Delta b2_k
        Wc = [_mod_rows(sum(W2[o] * W1[o][r] for o in range(mid)), primes) for r in range(in_dim)]
        Bc = _mod_rows(sum(W2[o] * _ntt_delta_plain(orc, b1p[o], primes) for o in range(mid)) + _ntt_delta_plain(orc, b2p[k], primes), primes)
        got = _weighted_sum_ntt(orc, primes, x, Wc, Bc)
        assert np.array_equal(got, want[k]), k


def test_pool_bn_fc_fc_equal_window_sums_and_one_composed_layer(env):
    primes, t, orc, rng = env
    xd = yd = 3; zd = 2; pxs = pys = 1; pxf = pyf = 2                # avg-pool 2x2 stride 1 on 3x3 -> 2x2, two channels
    pxo = (xd - pxf) // pxs + 1
    per_channel = pxo * pxo
    in_dim, mid, out = zd * per_channel, 3, 2
    f = lambda k: rng.uniform(-1, 1, size=k).astype(np.float32)
    mp, vp = orc.encode_many(f(zd)), orc.encode_many(rng.uniform(-3, 3, size=zd).astype(np.float32))
    w1p, b1p, w2p, b2p = (orc.encode_many(f(k)) for k in (mid * in_dim, mid, out * mid, out))
    d, cc = orc.encode(1.0 / (pxf * pyf))
    x = random_cts(rng, N, primes, zd * xd * yd)
    ct = x.shape[1:]
    y = orc.bn(orc.pool(x, xd, yd, zd, pxs, pys, pxf, pyf, d, cc), zd, pxo, pxo, mp, vp).reshape(in_dim, *ct)
    want = orc.fc(orc.fc(y, in_dim, mid, w1p, b1p).reshape(mid, *ct), mid, out, w2p, b2p).reshape(out, *ct)
    # window sums P[c, i, j] of the input, in the fully connected layer's input order
    xi = x.reshape(zd, xd, yd, *ct)
    P = np.stack([orc.add_many(np.stack([xi[c, i * pxs + a, j * pys + b] for a in range(pxf) for b in range(pyf)]))
                  for c in range(zd) for i in range(pxo) for j in range(pxo)])
    s_ntt = _ntt_plain(orc, d[:cc] if cc else d)
    C = [_mod_rows(s_ntt * _ntt_plain(orc, vp[c]), primes) for c in range(zd)]
    D = [_mod_rows(_ntt_delta_plain(orc, mp[c], primes) * _ntt_plain(orc, vp[c]), primes) for c in range(zd)]
    W1 = [[_ntt_plain(orc, w1p[o * in_dim + r]) for r in range(in_dim)] for o in range(mid)]
    for k in range(out):
        W2 = [_ntt_plain(orc, w2p[k * mid + o]) for o in range(mid)]
        Wc = [_mod_rows(sum(W2[o] * W1[o][r] for o in range(mid)), primes) for r in range(in_dim)]
        Bc = _mod_rows(sum(W2[o] * _ntt_delta_plain(orc, b1p[o], primes) for o in range(mid)) + _ntt_delta_plain(orc, b2p[k], primes), primes)
        # the affine map in front: x_r = P_r (.) C_c(r) - D_c(r)   =>   W[k,r] (.)= C_c(r),  B_k -= sum_r W[k,r] (.) D_c(r)   (unscaled W in the bias term)
        Bf = _mod_rows(Bc - sum(Wc[r] * D[r // per_channel] for r in range(in_dim)), primes)
        Wf = [_mod_rows(Wc[r] * C[r // per_channel], primes) for r in range(in_dim)]
        got = _weighted_sum_ntt(orc, primes, P, Wf, Bf)
        assert np.array_equal(got, want[k]), k
