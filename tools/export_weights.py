"""Exports the reference's trained MNIST weights (data, not code) to weights/*.npz.

Run once in the dev container (needs /root/reference): reads PlainModel/*.pth with torch.load and
cross-checks every tensor against the raw float32-LE payload of the matching .h5 file (the .h5
files h5py wrote store each dataset as one contiguous block), then writes
weights/<name>.npz.  bench.py and the tests load the .npz; nothing reads /root/reference at run time.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/PlainModel"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights")


def main():
    os.makedirs(OUT, exist_ok=True)
    for name in ["PlainModel", "ApproxPlainModel", "PlainModelWoPad", "PlainModelTiny"]:
        sd = torch.load(os.path.join(REF, name + ".pth"), map_location="cpu")
        raw = open(os.path.join(REF, name + ".h5"), "rb").read()
        arrays = {}
        for k, v in sd.items():
            if v.ndim == 0:
                continue  # num_batches_tracked
            a = v.detach().numpy().astype(np.float32)
            if a.size >= 16 and raw.find(a.tobytes()) < 0:
                print("WARNING: %s/%s not found verbatim in the .h5 payload" % (name, k), file=sys.stderr)
            arrays[k] = a
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        print(name, {k: a.shape for k, a in arrays.items()})


if __name__ == "__main__":
    main()
