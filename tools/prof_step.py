#!/usr/bin/env python
"""One forward pass of the bench workload for ncu (no torch import, so the profiler starts in seconds).

  ncu --profile-from-start off ... python tools/prof_step.py --batch 8

A warm-up forward builds the resident weight forms; the profiled region (cuProfilerStart/Stop) is exactly one
forward of the encoded PlainModel.h5 network on synthetic ciphertexts, as in bench.py's device-resident arm.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from crcnn_b200 import nets  # noqa: E402
from crcnn_b200.lib import Engine  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--model", default=bench.MODEL)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--last", type=int, default=None)
    ap.add_argument("--path", default="host", choices=("host", "layers"),
                    help="host: the C++17 host path bench.py times (crcnn_b200::Network::forward_dev with its layer fusion); "
                         "layers: one C-ABI call per reference layer (crcnn_b200/nets.py), any sub-range with --first/--last")
    args = ap.parse_args()
    n, primes, t = bench.N_POLY, bench.PRIMES, bench.T_PLAIN
    if args.path == "host" and args.first == 0 and args.last is None:
        from crcnn_b200 import host
        rng = np.random.default_rng(5)
        evk = bench.synth_evk(rng, primes, n)
        net = host.HostNetwork(n, primes, t, args.model, device=0, evk=evk)
        per_image = net.zd * net.xd * net.yd
        K = len(primes)
        pin, own = host.pinned_array(args.batch * per_image * 2 * K * (n + 1))
        bench.synth_residues(rng, (args.batch * per_image, 2), primes, n, out=pin.reshape(args.batch * per_image, 2, K, n + 1))
        net.resident_begin(own.ptr, args.batch)
        net.resident_run(2)
        cuda = ctypes.CDLL("libcuda.so.1")
        cuda.cuProfilerStart()
        net.resident_run(1)
        cuda.cuProfilerStop()
        net.resident_end()
        print("profiled one forward of the C++ host path, batch %d" % args.batch)
        net.close()
        return
    eng = Engine(n, primes, t, device=0)
    rng = np.random.default_rng(5)
    evk_words, sizes, dbc = bench.synth_evk(rng, primes, n)
    net = nets.Network(eng, args.model, evk=eng.evk_upload(evk_words, sizes, dbc))
    zd, xd, yd = net.input_shape
    # input of layer `first` (synthetic residues have the right distribution at any depth)
    first = args.first
    nin = nets.layer_io_counts(net.layers[first])[0]
    one = bench.synth_residues(rng, (nin, 2), primes, n)
    x0 = eng.upload(np.concatenate([one] * args.batch))
    cuda = ctypes.CDLL("libcuda.so.1")

    def fwd():
        x = eng.slice(x0, 0, args.batch * nin)
        y = net.forward(x, batch=args.batch, first=first, last=args.last)
        x.free()
        y.free()
        eng.sync()

    fwd()
    cuda.cuProfilerStart()
    fwd()
    cuda.cuProfilerStop()
    print("profiled one forward, batch %d, layers [%d, %s)" % (args.batch, first, args.last))
    eng.close()


if __name__ == "__main__":
    main()
