#!/usr/bin/env python
"""Roofline denominators measured on the box (crcnn_probe_pipe): 64x64->128 MAC chain, IMAD.WIDE issue rate, UMMA kind::i8 rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from crcnn_b200.lib import Engine
eng = Engine(bench.N_POLY, bench.PRIMES, bench.T_PLAIN, device=0)
for which, name, args in ((0, "64x64->128 MAC chain (8 per thread)      MAC/s", (148 * 8, 256, 4096)),
                          (1, "IMAD.WIDE.U32 independent chains       instr/s", (148 * 8, 256, 4096)),
                          (2, "tcgen05.mma kind::i8 128x256x32 smem   MAC/s", (148, 128, 4096)),
                          (3, "  same, N = 224                        MAC/s", (148, 128, 4096)),
                          (4, "  N = 224, 4 stages, commit every 4th  MAC/s", (148, 128, 4096))):
    eng.probe_pipe(which, *args[:2], 256)
    best = max(eng.probe_pipe(which, *args) for _ in range(5))
    extra = ""
    if which == 1:
        extra = "  = %.3f warp-instr/clk/SM at 1965 MHz" % (best[0] / 32 / 148 / 1.965e9)
    if which >= 2:
        extra = "  = %.1f int8 TOPS" % (2 * best[0] / 1e12)
    print("%-58s %.4g (%.2f ms)%s" % (name, best[0], best[1], extra))
eng.close()
