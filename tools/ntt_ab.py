#!/usr/bin/env python
"""A/B timing of the transform-bound classes at n = 8192 (K = 4): NTT round trip, square, relinearize.
Prints one compact JSON line (CUDA-event ms per class); run under gpurun after every kernel change.

  python tools/ntt_ab.py [--cts 2048] [--tag name]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crcnn_b200.lib import Engine  # noqa: E402
from oracle.port import DEFAULT_PRIMES_128  # noqa: E402  (prime table only)
import bench  # noqa: E402
from tools.kernel_sweep import measure  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cts", type=int, default=2048)
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    n = args.n
    primes = DEFAULT_PRIMES_128[n]
    eng = Engine(n, primes, 1 << 30 if n >= 8192 else 1 << 18)
    rng = np.random.default_rng(1)
    x = eng.upload(bench.synth_residues(rng, (args.cts, 2), primes, n))
    evk_words, sizes, dbc = bench.synth_evk(rng, primes, n)
    evk = eng.evk_upload(evk_words, sizes, dbc)
    out = {"tag": args.tag, "n": n, "cts": args.cts}

    def roundtrip():
        eng.to_ntt(x); eng.from_ntt(x)
    r = measure(eng, roundtrip, reps=5)
    out["fwd_ms"] = round(r["ntt_forward"]["ms"], 4)
    out["inv_ms"] = round(r["ntt_inverse"]["ms"], 4)
    out["fwd_gbfly_s"] = round(r["ntt_forward"]["gops"], 1)
    out["inv_gbfly_s"] = round(r["ntt_inverse"]["gops"], 1)
    r = measure(eng, lambda: eng.square(x).free(), reps=3)
    out["square"] = {k: round(v["ms"], 4) for k, v in r.items()}
    out["square_total_ms"] = round(sum(v["ms"] for v in r.values()), 4)
    t3 = eng.square(x)
    r = measure(eng, lambda: eng.relinearize(t3, evk).free(), reps=3)
    out["relin"] = {k: round(v["ms"], 4) for k, v in r.items()}
    out["relin_total_ms"] = round(sum(v["ms"] for v in r.values()), 4)
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
