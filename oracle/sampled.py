"""TEST INFRASTRUCTURE ONLY.  Sampled-oracle check of a FULL-SIZE encoded network on the GPU.

The reference's layers are sums over a window of input ciphertexts, one output ciphertext at a time
(CrCNN/src/convolutionalLayer.cpp:56-93, fullyConnectedLayer.cpp:113-168, poolingLayer.cpp:22-44,
batchNormLayer.cpp:29-40, squareLayer.cpp:22-71): one output depends on its own window only.  So a layer
run at its real shape (PlainModel.h5 at n = 8192, batch 8: 3,136 GEMM columns, fan-in 1250 x 500 outputs,
14,400 squares ...) is checked by drawing output ciphertexts at random, downloading exactly the input
ciphertexts of their windows from the device, and letting the CPU oracle compute those outputs with a
one-filter / few-row layer of the same geometry.  Every layer's check starts from the GPU's own previous
activation, so a network whose every layer passes is correct end to end by induction.

Only tests/, __graft_entry__.smoke() and bench.py's checker leg (outside the timed region) use this.
"""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np


def _download_cts(eng, t, idxs):
    """Ciphertexts `idxs` of device tensor t as [len][2][K][n+1] (coefficient form), contiguous runs in one copy each."""
    idxs = [int(i) for i in idxs]
    out = np.empty((len(idxs), 2, eng.K, eng.stride), dtype=np.uint64)
    order = np.argsort(idxs, kind="stable")
    p = 0
    while p < len(order):
        q = p
        while q + 1 < len(order) and idxs[order[q + 1]] == idxs[order[q]] + 1:
            q += 1
        s = eng.slice(t, idxs[order[p]], q - p + 1)
        blk = eng.download(s)
        s.free()
        for r in range(p, q + 1):
            out[order[r]] = blk[r - p]
        p = q + 1
    return out


def _pick(rng, total, want, always=()):
    """`want` distinct indices in [0,total), always including `always` (first / last / tile-boundary cases)."""
    s = [int(a) % total for a in always]
    pool = [int(v) for v in rng.permutation(total)[:want + len(s)]]
    for v in pool:
        if len(s) >= max(want, len(set(s))):
            break
        if v not in s:
            s.append(v)
    return sorted(set(s))


def layer_jobs(net, i, B, rng, samples):
    """Sampled outputs of layer i at batch B: list of (out_index, in_indices, kind-specific payload)."""
    layer = net.layers[i]
    kind = layer[0]
    jobs = []
    if kind == "conv":
        _, name, xd, yd, zd, xs, ys, xf, yf, nf = layer
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        nin, nout = zd * xd * yd, nf * xo * yo
        corner = [0, nout - 1, (B - 1) * nout + nout - 1, (B - 1) * nout + (nf - 1) * xo * yo, (min(nf, 33) - 1) * xo * yo + 1]
        for o in _pick(rng, B * nout, samples, corner):
            b, r = divmod(o, nout)
            k, r = divmod(r, xo * yo)
            oi, oj = divmod(r, yo)
            ins = [b * nin + (z * xd + oi * xs + kx) * yd + oj * ys + ky for z in range(zd) for kx in range(xf) for ky in range(yf)]
            jobs.append((o, ins, k))
    elif kind in ("pool", "avgpool"):
        _, name, xd, yd, zd, xs, ys, xf, yf = layer
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        nin, nout = zd * xd * yd, zd * xo * yo
        for o in _pick(rng, B * nout, samples, [0, B * nout - 1]):
            b, r = divmod(o, nout)
            z, r = divmod(r, xo * yo)
            oi, oj = divmod(r, yo)
            ins = [b * nin + (z * xd + oi * xs + kx) * yd + oj * ys + ky for kx in range(xf) for ky in range(yf)]
            jobs.append((o, ins, None))
    elif kind == "bn":
        _, name, zd, xd, yd = layer
        per = zd * xd * yd
        for o in _pick(rng, B * per, samples, [0, B * per - 1]):
            jobs.append((o, [o], (o % per) // (xd * yd)))
    elif kind == "square":
        _, name, zd, xd, yd = layer
        per = zd * xd * yd
        for o in _pick(rng, B * per, samples, [0, B * per - 1, B * per // 2, B * per // 2 - 1]):
            jobs.append((o, [o], None))
    elif kind == "fc":
        _, name, in_dim, out_dim = layer
        rows_always = [0, out_dim - 1, min(out_dim - 1, 127), min(out_dim - 1, 128), min(out_dim - 1, 63), min(out_dim - 1, 64)]
        per_image = max(1, samples // 2)
        for b in sorted({0, B - 1}):   # one job per output row (they run in parallel); the image's inputs are downloaded once
            for r in _pick(rng, out_dim, per_image, rows_always[:max(2, per_image)]):
                jobs.append(([b * out_dim + r], [b * in_dim + j for j in range(in_dim)], [r]))
    else:
        raise ValueError(kind)
    return jobs


def check_layer(eng, orc, net, i, x, y, B, rng, samples, evk_host=None, pool=None):
    """Compares `samples` output ciphertexts of layer i (device tensors x -> y) with the oracle.  Returns the number checked;
    raises AssertionError naming the layer and the ciphertext on the first difference."""
    layer = net.layers[i]
    kind, name = layer[0], layer[1]
    P = net.params.get(name)
    jobs = layer_jobs(net, i, B, rng, samples)

    prepared, cache = [], {}
    for o, ins, extra in jobs:   # device access stays on this thread (one context = one stream); only the oracle runs in the pool
        key = (ins[0], len(ins)) if kind == "fc" else None
        xin = cache.get(key) if key else None
        if xin is None:
            xin = _download_cts(eng, x, ins)
            if key:
                cache[key] = xin
        outs = o if isinstance(o, list) else [o]
        got = _download_cts(eng, y, outs)
        prepared.append((outs, xin, extra, got))

    def oracle(item):
        outs, xin, extra, got = item
        if kind == "conv":
            _, _, xd, yd, zd, xs, ys, xf, yf, nf = layer
            w, b = P
            k = extra
            want = orc.conv(xin, xf, yf, zd, 1, 1, xf, yf, 1, orc.encode_many(w.reshape(nf, -1)[k]), orc.encode_many(b[k:k + 1]))
        elif kind in ("pool", "avgpool"):
            _, _, xd, yd, zd, xs, ys, xf, yf = layer
            if kind == "avgpool":
                d, cc = orc.encode(1.0 / (xf * yf))   # a double: avgPoolingLayer.cpp:10-13
                want = orc.pool(xin, xf, yf, 1, 1, 1, xf, yf, d, cc)
            else:
                want = orc.pool(xin, xf, yf, 1, 1, 1, xf, yf)
        elif kind == "bn":
            m, v = P
            z = extra
            want = orc.bn(xin, 1, 1, 1, orc.encode_many(m[z:z + 1]), orc.encode_many(v[z:z + 1]))
        elif kind == "square":
            words, sizes, dbc = evk_host
            want = orc.square_layer(xin, words, sizes, dbc)
        else:
            _, _, in_dim, out_dim = layer
            w, b = P
            rows = extra
            want = orc.fc(xin, in_dim, len(rows), orc.encode_many(w.reshape(out_dim, in_dim)[rows]), orc.encode_many(b[rows]))
        want = np.asarray(want).reshape(got.shape)
        for r, o in enumerate(outs):
            if not np.array_equal(want[r], got[r]):
                bad = np.argwhere(want[r] != got[r])
                raise AssertionError("layer %d (%s %s): output ciphertext %d differs from the oracle at %d words, first at %s"
                                     % (i, kind, name, o, len(bad), bad[0].tolist()))
        return len(outs)

    if pool is None:
        return sum(oracle(it) for it in prepared)
    return sum(pool.map(oracle, prepared))


def check_network(eng, orc, net, x, B, evk_host=None, samples=6, seed=0, first=0, last=None, threads=None, keep_last=False, log=None):
    """Runs layers [first,last) of `net` on device tensor x (batch B) one by one through the engine and checks `samples` outputs of
    every layer against the oracle.  Returns ({layer name: ciphertexts checked}, output tensor or None)."""
    rng = np.random.default_rng(seed)
    last = net.num_layers() if last is None else last
    threads = threads or min(16, os.cpu_count() or 1)
    checked = {}
    with ThreadPoolExecutor(max_workers=threads) as pool:
        for i in range(first, last):
            y = net.forward_layer(i, x, B)
            checked["%d:%s" % (i, net.layers[i][1])] = check_layer(eng, orc, net, i, x, y, B, rng, samples, evk_host, pool)
            if log:
                log("layer %d %s: %d sampled ciphertexts bit-identical to the oracle" % (i, net.layers[i][1], checked["%d:%s" % (i, net.layers[i][1])]))
            if i > first:
                x.free()
            x = y
    if keep_last:
        return checked, x
    if last > first:
        x.free()
    return checked, None
