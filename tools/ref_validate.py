#!/usr/bin/env python
"""Baseline hygiene for the CPU reference arm, run on the GPU box's host cores:
  (1) ONE UNCROPPED layer of the bench network timed for real (conv2 of PlainModel.h5: 13x13x20 -> 50 filters 3x3 / 2, 324,000
      ciphertext x plaintext terms) next to the figure bench.py extrapolates for it from its cropped sample -- validates the
      linear extrapolation;
  (2) the whole-network estimate with the reference's LITERAL thread constants (CrCNN/src/cnnBuilder.cpp:109: 40 / 50 threads;
      conv1 with 20 filters and fc4 with 10 rows then run on ONE thread: convolutionalLayer.cpp:28-31,177-187) next to the
      one-thread-per-core variant bench.py reports.
Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from crcnn_b200 import nets  # noqa: E402
from oracle import ref as oref  # noqa: E402


def main():
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    res = {"cores": cores}
    fair = bench.run_reference_sample(30.0, rng)
    res["fair_threads"] = {"images_per_s": fair[0], "per_layer_ms": {k: 1000 * v for k, v in fair[3].items()}}
    lit = bench.run_reference_sample(30.0, rng, literal_threads=True)
    res["literal_threads"] = {"images_per_s": lit[0], "per_layer_ms": {k: 1000 * v for k, v in lit[3].items()}, "sample": lit[2]}
    # the uncropped layer
    r = oref.Ref(bench.N_POLY, bench.T_PLAIN, seed=1)
    w = nets.load_weights(bench.MODEL)
    name = "pool2_features.conv2"
    x = bench.synth_residues(rng, (20 * 13 * 13, 2), r.primes, bench.N_POLY)
    t0 = time.perf_counter()
    enc, first, steady = r.conv_timed(x, 13, 13, 20, 2, 2, 3, 3, 50, w[name + ".weight"].ravel(), w[name + ".bias"], th=min(cores, 50), reps=1)
    res["uncropped_conv2"] = {"terms": 50 * 36 * 180, "threads": min(cores, 50), "encode_s": enc, "first_forward_s": first, "steady_forward_s": steady,
                              "wall_s": time.perf_counter() - t0,
                              "extrapolated_from_cropped_sample_s": fair[3][name],
                              "extrapolation_error": fair[3][name] / steady - 1.0}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
