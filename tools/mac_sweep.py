"""Times the weighted-sum kernel for one tile shape (CRCNN_MAC_TILE) on conv2- and fc-shaped layers
of the benchmark network (n = 8192).  Usage: python tools/mac_sweep.py 22 [batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tile = sys.argv[1] if len(sys.argv) > 1 else "0"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
os.environ["CRCNN_MAC_TILE"] = tile

import numpy as np  # noqa: E402
from crcnn_b200.lib import Engine  # noqa: E402
import bench  # noqa: E402

eng = Engine(bench.N_POLY, bench.PRIMES, bench.T_PLAIN)
rng = np.random.default_rng(0)
res = {}
for name, kind, shape in [("conv2", "conv", (13, 13, 20, 2, 2, 3, 3, 50)), ("fc_1250x64", "fc", (1250, 64))]:
    if kind == "conv":
        xd, yd, zd, xs, ys, xf, yf, nf = shape
        nin, terms = zd * xd * yd, nf * 36 * zd * xf * yf
        w = eng.plain_encode(rng.uniform(-1, 1, nf * zd * xf * yf).astype(np.float32))
        b = eng.plain_encode(rng.uniform(-1, 1, nf).astype(np.float32))
    else:
        i, o = shape
        nin, terms = i, i * o
        w = eng.plain_encode(rng.uniform(-1, 1, i * o).astype(np.float32))
        b = eng.plain_encode(rng.uniform(-1, 1, o).astype(np.float32))
    x = eng.upload(bench.synth_residues(rng, (B * nin, 2), bench.PRIMES, bench.N_POLY))
    eng.to_ntt(x)
    run = (lambda: eng.conv(x, w, b, B, *shape)) if kind == "conv" else (lambda: eng.fc(x, w, b, B, *shape))
    run().free()
    eng.sync()
    eng.prof_reset(); eng.prof_enable(True)
    for _ in range(3):
        run().free()
    eng.sync()
    ms = eng.prof()["weighted_sum_mac"][1] / 3
    eng.prof_enable(False)
    macs = terms * B * 2 * eng.K * eng.n
    res[name] = (round(ms, 2), round(macs / ms / 1e9, 1))
print("tile", tile, "batch", B, {k: "%s ms, %s TMAC/s" % (v[0], v[1] / 1000) for k, v in res.items()})
