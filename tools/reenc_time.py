#!/usr/bin/env python
"""Device re-encryption at the shape of the bench network's re-encryption point: the activations before layer 6 of
PlainModel.h5 at n = 8192 are batch x 50 x 5 x 5 ciphertexts.  Prints ms per batch and per image (CUDA-event time of the class)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from crcnn_b200.lib import Engine

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n, primes, t = bench.N_POLY, bench.PRIMES, bench.T_PLAIN
eng = Engine(n, primes, t)
rng = np.random.default_rng(1)
K = len(primes)
sk = bench.synth_residues(rng, (), primes, n)            # any residues do for timing: the work is data independent
pk = bench.synth_residues(rng, (2,), primes, n)
keys = eng.keys_upload(sk, pk)
x = eng.upload(bench.synth_residues(rng, (batch * 1250, 2), primes, n))
eng.reencrypt(keys, x, seed=1).free()
eng.sync()
eng.prof_enable(True); eng.prof_reset()
for i in range(3):
    eng.reencrypt(keys, x, seed=2 + i).free()
eng.sync()
ms = eng.prof()["reencrypt"][1] / 3
print("device re-encryption of %d ciphertexts (batch %d x 50x5x5, n=%d): %.2f ms per batch, %.3f ms per image; "
      "the reference's host re-encryption: 3.2 s per image at n=4096 (Doc/Tesi.lyx:13020-13700)" % (batch * 1250, batch, n, ms, ms / batch))
eng.close()
