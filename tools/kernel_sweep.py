#!/usr/bin/env python
"""Kernel sweep (BASELINE.json config 5): NTT / INTT, multiply_plain_ntt, add_plain, add_many, square, relinearize and the
weighted sum (tensor-core and CUDA-core kernels) at n = 4096 / 8192 / 16384 (K = 2 / 4 / 8 limbs), each against its
roofline.  Times are CUDA-event times of the kernel classes (crcnn_prof_get), work is the engine's own count of
algorithmic bytes and operations (crcnn_prof_get_work).

  python tools/kernel_sweep.py [--cts 512] > profiles/r01_kernel_sweep.json
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crcnn_b200.lib import Engine  # noqa: E402
from crcnn_b200.nets import DEFAULT_PRIMES_128  # noqa: E402
import bench  # noqa: E402

T = {4096: 1 << 18, 8192: 1 << 30, 16384: 1 << 30}


def measure(eng, fn, reps=3):
    fn()
    eng.sync()
    eng.prof_reset(); eng.prof_enable(True)
    for _ in range(reps):
        fn()
    eng.sync()
    prof, work = eng.prof(), eng.prof_work()
    eng.prof_enable(False)
    out = {}
    for k, (launches, ms) in prof.items():
        if launches and ms > 0:
            b, o = work[k]
            out[k] = {"ms": ms / reps, "launches": launches / reps, "alg_gb": b / reps / 1e9, "gbs": b / ms / 1e6, "gops": o / ms / 1e6}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cts", type=int, default=8192, help="ciphertexts at n = 8192 (scaled by 8192/n for the other degrees): HBM-bound kernels need GBs of data to show their rate")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    res = {"hbm_peak_gbs": hbm, "cts": args.cts, "sweep": {}}
    for n in (4096, 8192, 16384):
        primes = DEFAULT_PRIMES_128[n]
        K = len(primes)
        eng = Engine(n, primes, T[n])
        rng = np.random.default_rng(n)
        C = max(8, args.cts * 8192 // n)
        x = eng.upload(bench.synth_residues(rng, (C, 2), primes, n))
        evk_words, sizes, dbc = bench.synth_evk(rng, primes, n)
        evk = eng.evk_upload(evk_words, sizes, dbc)
        pl = eng.plain_encode(np.array([0.37, -1.25], dtype=np.float32))
        probe_ms = eng.probe_imad(148 * 8, 256, 4096)
        probe = 148 * 8 * 256 * 4096 * 8 / (probe_ms / 1000.0)
        wide, _ = eng.probe_pipe(1, 148 * 8, 256, 4096)
        r = {"K": K, "ciphertexts": C, "int_pipe_probe_gmac_s": probe / 1e9, "imad_wide_probe_ginstr_s": wide / 1e9}

        def roundtrip():
            eng.to_ntt(x); eng.from_ntt(x)
        r["ntt_roundtrip"] = measure(eng, roundtrip)
        eng.to_ntt(x)
        r["multiply_plain_ntt"] = measure(eng, lambda: eng.plain_op(x, pl, 0, "mul"))
        r["add_plain"] = measure(eng, lambda: eng.plain_op(x, pl, 1, "add"))
        r["add_many"] = measure(eng, lambda: eng.add_many(x).free())
        eng.from_ntt(x)
        sq = eng.upload(bench.synth_residues(rng, (max(8, C // 8), 2), primes, n))
        r["square"] = measure(eng, lambda: eng.square(sq).free())
        t3 = eng.square(sq)
        r["relinearize"] = measure(eng, lambda: eng.relinearize(t3, evk).free())
        # weighted sum, conv2-like: 3x3 window over 20 channels, 50 outputs, 6x6 positions
        shape = (13, 13, 20, 2, 2, 3, 3, 50)
        Bc = max(1, 4 * 8192 // n)
        xin = eng.upload(bench.synth_residues(rng, (Bc * 20 * 13 * 13, 2), primes, n))
        eng.to_ntt(xin)
        wv, bv = rng.uniform(-1, 1, 50 * 180).astype(np.float32), rng.uniform(-1, 1, 50).astype(np.float32)
        for mode, name in ((1, "weighted_sum_tensor_core"), (0, "weighted_sum_cuda_core")):
            eng.set_tensor_core_mode(0)
            eng.set_limb_split_mode(mode)
            w, b = eng.plain_encode(wv), eng.plain_encode(bv)
            r[name] = measure(eng, lambda: eng.conv(xin, w, b, Bc, *shape).free())
            r[name]["batch"] = Bc
        for k, v in r.items():
            if isinstance(v, dict):
                for cls, e in v.items():
                    if isinstance(e, dict) and "gbs" in e:
                        e["hbm_frac"] = e["gbs"] / hbm
        res["sweep"]["n=%d" % n] = r
        eng.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
