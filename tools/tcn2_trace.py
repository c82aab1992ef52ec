#!/usr/bin/env python
"""Timeline of ONE CTA of tcn2_mac_kernel (debug build: make EXTRA=-DCRCNN_TCN2_TRACE OUT=../../ab/libT.so OBJDIR=../../build/objT).

  CRCNN_B200_LIB=$PWD/ab/libT.so python tools/tcn2_trace.py --layer 0      # conv1 of the bench network (3 = conv2)

Prints, averaged over the recorded chunks of CTA 0 (SM clock cycles): where the MMA issuer waits (accumulator, input planes),
how long the first epilogue warp waits for the accumulator and how long its share of a chunk takes, and the chunk period."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crcnn_b200 import nets, lib  # noqa: E402
from crcnn_b200.lib import Engine  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", type=int, default=0)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    n, primes, t = bench.N_POLY, bench.PRIMES, bench.T_PLAIN
    eng = Engine(n, primes, t, device=0)
    rng = np.random.default_rng(5)
    evk_words, sizes, dbc = bench.synth_evk(rng, primes, n)
    net = nets.Network(eng, bench.MODEL, evk=eng.evk_upload(evk_words, sizes, dbc))
    nin = nets.layer_io_counts(net.layers[args.layer])[0]
    one = bench.synth_residues(rng, (nin, 2), primes, n)
    x0 = eng.upload(np.concatenate([one] * args.batch))
    eng.to_ntt(x0)

    def fwd():
        x = eng.slice(x0, 0, args.batch * nin)
        y = net.forward(x, batch=args.batch, first=args.layer, last=args.layer + 1)
        x.free(); y.free(); eng.sync()

    fwd()
    L = lib.load()
    L.crcnn_debug_tcn2_trace_reset.restype = C.c_int
    L.crcnn_debug_tcn2_trace_read.restype = C.c_int
    L.crcnn_debug_tcn2_trace_read.argtypes = [C.c_int, C.c_void_p, C.c_int]
    assert L.crcnn_debug_tcn2_trace_reset() == 0
    fwd()
    ev = []
    for role in range(3):
        buf = np.zeros(8192, dtype=np.uint64)
        cnt = L.crcnn_debug_tcn2_trace_read(role, buf.ctypes.data, 8192)
        tags = (buf[:cnt] >> np.uint64(56)).astype(np.int64)
        ts = (buf[:cnt] & np.uint64((1 << 56) - 1)).astype(np.int64)
        ev.append((tags, ts))
    # MMA issuer: 100 wait-acc, 101 acc free, plane arrivals, 102 issued
    tags, ts = ev[1]
    i100 = np.nonzero(tags == 100)[0]
    rows = []
    for a, b in zip(i100[:-1], i100[1:]):
        seg_t, seg_tag = ts[a:b], tags[a:b]
        if 101 not in seg_tag or 102 not in seg_tag:
            continue
        t100, t101, t102 = seg_t[0], seg_t[seg_tag == 101][0], seg_t[seg_tag == 102][0]
        planes = seg_t[(seg_tag < 100)]
        rows.append((t101 - t100, planes[-1] - t101, t102 - planes[-1], ts[b] - t100, planes - t101))
    r = np.array([x[:4] for x in rows][5:], dtype=np.float64)
    print("chunks analysed:", len(r))
    print("MMA issuer per chunk [clk]: wait accumulator %.0f | wait input planes %.0f | issue after last plane %.0f | chunk period %.0f"
          % tuple(r.mean(axis=0)))
    pl = np.array([x[4] for x in rows[5:] if len(x[4]) == len(rows[5][4])], dtype=np.float64)
    print("plane arrival after accumulator free [clk], per plane in stream order:", np.round(pl.mean(axis=0)).astype(int).tolist())
    tags, ts = ev[2]
    i200 = np.nonzero(tags == 200)[0]
    e = []
    for a, b in zip(i200[:-1], i200[1:]):
        seg_t, seg_tag = ts[a:b], tags[a:b]
        if 201 in seg_tag and 202 in seg_tag:
            e.append((seg_t[seg_tag == 201][0] - seg_t[0], seg_t[seg_tag == 202][0] - seg_t[seg_tag == 201][0]))
    e = np.array(e[5:], dtype=np.float64)
    print("epilogue warp 2 per chunk [clk]: wait accumulator %.0f | recombine + store %.0f" % tuple(e.mean(axis=0)))
    # raw timeline of three consecutive chunks: when each plane's load was issued (producer) and when the MMA issuer saw it
    tags1, ts1 = ev[1]
    tags0, ts0 = ev[0]
    i101 = np.nonzero(tags1 == 101)[0]
    if len(i101) > 104 and os.environ.get("TCN2_TRACE_RAW"):
        t_ref = ts1[i101[100]]
        print("raw (cycles relative to 'accumulator free' of chunk 100):")
        for c in range(100, 103):
            a = i101[c]
            b = i101[c + 1]
            seg = [(int(tags1[k]), int(ts1[k] - t_ref)) for k in range(a, b)]
            print("  MMA   chunk %d: %s" % (c, seg))
        lo, hi = ts1[i101[99]], ts1[i101[103]]
        sel = (ts0 >= lo) & (ts0 <= hi)
        print("  producer issues (plane, t):", [(int(g), int(t - t_ref)) for g, t in zip(tags0[sel], ts0[sel])])
    tags, ts = ev[0]
    d = np.diff(ts)
    print("producer: %d stage-free events, median gap %.0f clk, mean %.0f clk" % (len(ts), np.median(d), d.mean()))
    eng.close()


if __name__ == "__main__":
    main()
