#!/usr/bin/env python
"""Benchmark of the encrypted-forward hot path (BASELINE.json metric: encrypted MNIST images/sec,
per-layer latency in ms, roofline fraction of the dominant kernel).

One step = one forward pass of the encoded PlainModel.h5 network (n = 8192, K = 4, t = 2^30,
32x32 zero-bordered input) over a batch of B independent
synthetic encrypted images on one GPU.  With --gpus N every rank runs its own batch (image
replicas, no data-path collective: weak scaling).

  python bench.py --gpus 1 --steps 3 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...     # the reference's own CPU path on the host cores

Prints ONE JSON line (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
through the C ABI with host buffers (pinned H2D of every input ciphertext and D2H of the 10 output
ciphertexts per image inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line: whatever NCCL has to say (its version banner under NCCL_DEBUG=VERSION/INFO) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from crcnn_b200 import nets  # noqa: E402

N_POLY = 8192
PRIMES = [0x7fffffff380001, 0x7ffffffef00001, 0x3fffffff000001, 0x3ffffffef40001]  # coeff_modulus_128(8192)
T_PLAIN = 1 << 30
MODEL = "PlainModel"
W = 8  # bytes per residue


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per step per GPU")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0, help="CPU seconds for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default=None, choices=sorted(nets.TOPOLOGIES),
                    help="network to run instead of the headline PlainModel (e.g. PlainModelTiny with --n 4096: BASELINE config 1)")
    ap.add_argument("--n", type=int, default=None, choices=[4096, 8192, 16384], help="polynomial degree (default 8192)")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "shard"],
                    help="replicas: every GPU runs its own image batch (default, weak scaling); shard: ONE image, the Approx "
                         "network's conv / fc layers split by output neuron across the GPUs with NCCL all-gathers of the "
                         "activation ciphertexts (BASELINE config 3, strong scaling)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------
def synth_residues(rng, shape_front, primes, n, out=None):
    """Uniform canonical residues in SEAL ciphertext layout [..., K, n+1] (pad word 0): the
    distribution FV ciphertexts have; the evaluator's work is data independent."""
    K = len(primes)
    if out is None:
        out = np.zeros(tuple(shape_front) + (K, n + 1), dtype=np.uint64)
    for j, q in enumerate(primes):
        out[..., j, :n] = rng.integers(0, q, size=tuple(shape_front) + (n,), dtype=np.uint64)
    out[..., n] = 0
    return out


def synth_evk(rng, primes, n, dbc=16):
    sizes = [2 * ((int(q).bit_length() + dbc - 1) // dbc) for q in primes]
    parts = [synth_residues(rng, (s,), primes, n).ravel() for s in sizes]
    return np.concatenate(parts), sizes, dbc


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU baseline: the reference's own layer classes (oracle/_ref) on a bounded sample
# ------------------------------------------------------------------------------------------
def _crop_layer(layer, cores, term_budget, ct_budget):
    """Same layer type and fan-in, reduced width/extent so the sample fits the time budget.
    Returns (cropped layer, scale) with scale = full work / sample work (work = weighted-sum terms for
    conv/fc, ciphertexts for the others; both are exactly linear in the cropped dimensions)."""
    kind = layer[0]
    if kind == "conv":
        _, name, xd, yd, zd, xs, ys, xf, yf, nf = layer
        nf2 = min(nf, cores)
        yo = (yd - yf) // ys + 1
        per_row = yo * zd * xf * yf                    # terms per filter per output row
        rows = max(1, min((xd - xf) // xs + 1, int(term_budget // per_row)))
        xd2 = (rows - 1) * xs + xf
        crop = ("conv", name, xd2, yd, zd, xs, ys, xf, yf, nf2)
        return crop, nets.layer_terms(layer) / nets.layer_terms(crop)
    if kind == "fc":
        _, name, i, o = layer
        o2 = min(o, cores)
        return ("fc", name, i, o2), o / o2
    if kind in ("pool", "avgpool"):
        _, name, xd, yd, zd, xs, ys, xf, yf = layer
        yo = (yd - yf) // ys + 1
        rows = max(1, min((xd - xf) // xs + 1, int(ct_budget // yo)))
        xd2 = (rows - 1) * xs + xf
        crop = (kind, name, xd2, yd, 1, xs, ys, xf, yf)
        return crop, nets.layer_io_counts(layer)[1] / nets.layer_io_counts(crop)[1]
    if kind == "bn":
        _, name, zd, xd, yd = layer
        xd2 = max(1, min(xd, int(ct_budget // yd)))
        return ("bn", name, 1, xd2, yd), (zd * xd * yd) / (xd2 * yd)
    if kind == "square":
        _, name, zd, xd, yd = layer
        zd2 = min(zd, cores)
        xd2 = max(1, min(xd, int(ct_budget // yd)))
        return ("square", name, zd2, xd2, yd), (zd * xd * yd) / (zd2 * xd2 * yd)
    raise ValueError(kind)


def run_reference_sample(budget_s, rng):
    """Times the UNMODIFIED reference layer classes (SEAL 2.3.1 Evaluator on the host cores) on a
    cropped PlainModel network and extrapolates linearly to the full layers.
    Returns (images_per_s, cores, description, per_layer_seconds_full)."""
    from oracle import ref as oref
    if not oref.available():
        return None
    cores = os.cpu_count() or 1
    r = oref.Ref(N_POLY, T_PLAIN, seed=1)
    weights = nets.load_weights(MODEL)
    layers = nets.TOPOLOGIES[MODEL]["layers"]
    # calibrate: one weighted-sum term and one generic multiply_plain, single thread
    x1 = synth_residues(rng, (2, 2), r.primes, N_POLY)
    t0 = time.perf_counter()
    r.fc(x1, 2, 2, np.full(4, 0.37, np.float32), np.full(2, 0.1, np.float32), th=1)
    c_term = max((time.perf_counter() - t0) / 4, 1e-4)
    t0 = time.perf_counter()
    r.bn(x1, 1, 2, 1, [0.3], [1.7])
    c_ct = max((time.perf_counter() - t0) / 2, 1e-4)
    per_layer = budget_s / len(layers)
    total, desc, full_times = 0.0, [], {}
    for layer in layers:
        crop, scale = _crop_layer(layer, cores, per_layer / c_term, per_layer / (4 * c_ct))
        kind, name = crop[0], crop[1]
        nin = nets.layer_io_counts(crop)[0]
        x = synth_residues(rng, (nin, 2), r.primes, N_POLY)
        t0 = time.perf_counter()
        if kind == "conv":
            _, _, xd, yd, zd, xs, ys, xf, yf, nf = crop
            w = weights[name + ".weight"][:nf].ravel(); b = weights[name + ".bias"][:nf]
            r.conv(x, xd, yd, zd, xs, ys, xf, yf, nf, w, b, th=min(cores, nf))
        elif kind == "fc":
            _, _, i, o = crop
            r.fc(x, i, o, weights[name + ".weight"][:o].ravel(), weights[name + ".bias"][:o], th=min(cores, o))
        elif kind in ("pool", "avgpool"):
            _, _, xd, yd, zd, xs, ys, xf, yf = crop
            r.pool(x, xd, yd, zd, xs, ys, xf, yf, avg=(kind == "avgpool"))
        elif kind == "bn":
            _, _, zd, xd, yd = crop
            var = weights[name + ".running_var"][:zd]
            r.bn(x, zd, xd, yd, weights[name + ".running_mean"][:zd], 1 / np.sqrt(var + 1e-5))
        elif kind == "square":
            _, _, zd, xd, yd = crop
            r.square_layer(x, zd, xd, yd, th=min(cores, zd))
        dt = time.perf_counter() - t0
        full_times[name] = dt * scale
        total += dt * scale
        desc.append("%s x%.3g" % (name.split(".")[-1], scale))
    sample = ("reference layer classes (SEAL 2.3.1, -O3) on a cropped %s net, n=%d; per-layer time x " % (MODEL, N_POLY) +
              "(full work / sample work): " + ", ".join(desc))
    return 1.0 / total, cores, sample, full_times


def main_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import ref as oref
    steps, warm = max(1, args.steps), max(0, args.warmup)
    if not oref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcrcnn_ref.so was not built (needs /root/reference at build time)"}))
        return
    rng = np.random.default_rng(0)
    per_step = max(4.0, 150.0 / (steps + warm))
    vals = []
    for i in range(steps + warm):
        ips, cores, sample, full = run_reference_sample(per_step, rng)
        if i >= warm:
            vals.append(ips)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "encrypted MNIST images/sec", "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "PlainModel.h5 encoded net (conv-avgpool-bn-conv-square-avgpool-bn-fc-fc), n=8192, K=4, t=2^30, 32x32 zero-bordered input",
                   "note": "host CPU only; per-image time extrapolated from a bounded sample"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "per_layer_ms": {k: 1000 * s for k, s in full.items()},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def main_shard(args, rank, world, local_rank):
    """BASELINE config 3: ApproxPlainModel.h5, n = 8192, one image, output-neuron sharding over the GPUs of one box.
    A step is one sharded forward (latency path); value = images/s = 1 / step time."""
    import torch
    import torch.distributed as dist
    from crcnn_b200.lib import Engine

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
    K, n = len(PRIMES), N_POLY
    eng = Engine(n, PRIMES, T_PLAIN, device=local_rank)
    rng = np.random.default_rng(1000)          # the same image and keys on every rank
    evk_words, sizes, dbc = synth_evk(rng, PRIMES, n)
    net = nets.ShardedNetwork(eng, "ApproxPlainModel", dist, evk=eng.evk_upload(evk_words, sizes, dbc))
    zd, xd, yd = net.input_shape
    x0 = eng.upload(synth_residues(rng, (zd * xd * yd, 2), PRIMES, n))
    ct_bytes = 2 * K * n * W
    # all-gather volume per image (bytes received per rank): activations before conv2, fc3, fc4 and the 10 scores
    gathered = (20 * 11 * 11 + 800 + 500 + 10) * ct_bytes

    def step():
        x = eng.slice(x0, 0, zd * xd * yd)
        y = net.forward(x)
        y.free()

    for _ in range(max(3, args.warmup)):
        step()
    eng.sync()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank); sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        step()
    ev[1].record()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([ev[0].elapsed_time(ev[1])], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "encrypted MNIST images/sec", "value": 1000.0 / ms, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic: uniform residues in SEAL ciphertext layout; weights = ApproxPlainModel.h5",
            "config": {"workload": "ApproxPlainModel.h5 encoded net, n=8192, K=4, t=2^30, ONE image, conv/fc layers sharded by output neuron",
                       "parallelism": "output-neuron shards x%d, NCCL all-gather of activation ciphertexts before conv2 / fc3 / fc4 and of the scores" % world,
                       "all_gather_bytes_per_image": gathered},
            "clocks": clocks}))
    eng.close()
    dist.destroy_process_group()


def main_b200(args, rank, world, local_rank):
    import torch
    from crcnn_b200.lib import Engine

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B, K, n = args.batch, len(PRIMES), N_POLY
    eng = Engine(n, PRIMES, T_PLAIN, device=local_rank)
    stream = torch.cuda.Stream()
    copy_stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(1000 + rank)
    evk_words, sizes, dbc = synth_evk(rng, PRIMES, n)
    net = nets.Network(eng, MODEL, evk=eng.evk_upload(evk_words, sizes, dbc))
    zd, xd, yd = net.input_shape
    per_image = zd * xd * yd
    ct_words = 2 * K * (n + 1)
    in_words = B * per_image * ct_words
    # pinned host buffers (torch for page-locked memory only)
    host_in = torch.empty(in_words, dtype=torch.int64, pin_memory=True)
    view = host_in.numpy().view(np.uint64).reshape(B * per_image, 2, K, n + 1)
    synth_residues(rng, (B * per_image, 2), PRIMES, n, out=view)
    n_scores = nets.layer_io_counts(net.layers[-1])[1]
    host_out = torch.empty(B * n_scores * ct_words, dtype=torch.int64, pin_memory=True)
    h2d = in_words * 8
    d2h = B * n_scores * ct_words * 8

    layer_names = [l[1] for l in net.layers]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(x0, events=None):
        x = eng.slice(x0, 0, B * per_image)  # fresh coefficient-form copy: the input NTT is part of every step
        cb = None
        if events is not None:
            def cb(i, layer):
                e = torch.cuda.Event(enable_timing=True); e.record(stream); events.append(e)
        y = net.forward(x, batch=B, on_layer=cb)
        x.free()
        return y

    with torch.cuda.stream(stream):
        x0 = eng.upload_ptr(host_in.data_ptr(), B * per_image)
        eng.sync()
        # ---- warm-up (also builds the resident NTT-form weights of conv1/conv2/fc4 once)
        for _ in range(max(3, args.warmup)):
            step_resident(x0).free()
        eng.sync()
        barrier()
        # ---- timed: device resident
        eng.prof_reset(); eng.prof_enable(True)
        sampler = ClockSampler(local_rank); sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        layer_events = []
        ev[0].record(stream)
        for s in range(args.steps):
            e0 = torch.cuda.Event(enable_timing=True); e0.record(stream)
            evs = [e0]
            step_resident(x0, evs).free()
            layer_events.append(evs)
        ev[1].record(stream)
        barrier()
        clocks = sampler.stop()
        ms_total = ev[0].elapsed_time(ev[1])
        prof = eng.prof()
        work = eng.prof_work()
        eng.prof_enable(False)
        # ---- timed: end to end through the C ABI with host buffers
        mem_free = [torch.cuda.mem_get_info()[0] / 1e9]
        bufs = [eng.alloc(B * per_image), eng.alloc(B * per_image)]
        # one untimed end-to-end step: the staging buffer and the two input tensors exist before the clock starts
        for b_ in bufs:
            eng.upload_into(b_, host_in.data_ptr(), copy_stream.cuda_stream)
        eng.wait_stream(copy_stream.cuda_stream)
        y = net.forward(bufs[0], batch=B)
        eng.download_ptr(y, host_out.data_ptr())
        y.free()
        barrier()
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record(stream)
        # double-buffered: the H2D copy of step s+1 runs on a copy stream while step s computes; the two input
        # tensors and the staging buffer are allocated once (crcnn_tensor_upload_into), nothing per step
        done = [None, None]                      # event: forward that consumed the buffer has finished
        up_ev = []                               # (start, end) of every upload on the copy stream
        fwd_ev = []                              # (start, forward done, download done) of every step on the compute stream
        def upload(i):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(copy_stream)
            eng.upload_into(bufs[i], host_in.data_ptr(), copy_stream.cuda_stream)
            b_.record(copy_stream)
            up_ev.append((a_, b_))
        upload(0)
        for s in range(args.steps):
            cur = bufs[s & 1]
            eng.wait_stream(copy_stream.cuda_stream)
            if s + 1 < args.steps:
                nxt = (s + 1) & 1
                if done[nxt] is not None:
                    copy_stream.wait_event(done[nxt])
                upload(nxt)
            f0, f1, f2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            f0.record(stream)
            y = net.forward(cur, batch=B)
            f1.record(stream)
            done[s & 1] = f1
            eng.download_ptr(y, host_out.data_ptr())
            y.free()
            f2.record(stream)
            fwd_ev.append((f0, f1, f2))
        ev2[1].record(stream)
        barrier()
        ms_e2e = ev2[0].elapsed_time(ev2[1])
        mem_free.append(torch.cuda.mem_get_info()[0] / 1e9)
        # the host link by itself: one plain pinned H2D copy of a step's input (explains e2e when the link is the limit)
        dev_buf = torch.empty(in_words, dtype=torch.int64, device="cuda")
        ev3 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev3[0].record(stream)
        dev_buf.copy_(host_in, non_blocking=True)
        ev3[1].record(stream)
        stream.synchronize()
        h2d_gbs = h2d / ev3[0].elapsed_time(ev3[1]) / 1e6
        del dev_buf

    # max over ranks
    t = torch.tensor([ms_total, ms_e2e], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    images = B * args.steps * world
    value = images / (ms_total / 1000.0)
    e2e_value = images / (ms_e2e / 1000.0)

    # per-layer latency (mean over steps, this rank)
    per_layer = {}
    for i, name in enumerate(layer_names):
        per_layer[name] = float(np.mean([evs[i].elapsed_time(evs[i + 1]) for evs in layer_events]))

    # ---- per-class rooflines from the engine's own work counters (crcnn_prof_get_work: algorithmic bytes and
    # operations counted at launch time) and the CUDA-event time of every launch of the class
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bf16_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    layers = net.layers
    # integer pipe: 64x64->128-bit multiply-accumulates per second of a register-only loop, measured in this run
    probe_ms = eng.probe_imad(148 * 8, 256, 4096)
    probe_rate = 148 * 8 * 256 * 4096 * 8 / (probe_ms / 1000.0)
    # hard denominators, measured in this run on this device (crcnn_probe_pipe): the IMAD.WIDE issue rate (independent chains, nothing
    # else in the loop) and the kind::i8 UMMA rate with smem-resident operands and no epilogue
    eng.probe_pipe(1, 148 * 8, 256, 512); eng.probe_pipe(2, 148, 128, 256)            # warm
    wide_rate, _ = eng.probe_pipe(1, 148 * 8, 256, 4096)                               # IMAD.WIDE thread-instructions / s
    umma_rate, _ = eng.probe_pipe(2, 148, 128, 4096)                                   # int8 MACs / s
    counters = {}  # ncu counters of each class's main kernel from the committed capture of this round (tools/make_traffic.py)
    try:
        counters = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("counters", {})
    except Exception:
        pass
    traffic = {}   # measured DRAM bytes per launch of each class's main kernel, from the committed ncu capture (tools/make_traffic.py)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = {k: v["dram_bytes_per_launch"] for k, v in tj["classes"].items()}
    except Exception:
        pass
    MAC_CLASSES = ("weighted_sum_mac", "behz_lift", "behz_floor_sk")
    BFLY_CLASSES = ("ntt_forward", "ntt_inverse", "plain_expand_ntt", "relinearize")
    BFLY_MACS = 2.5  # one Harvey butterfly = 10 IMAD-pipe instructions (mulhi64 + two mullo64) = 2.5 split-accumulator MACs of 4
    # IMAD-pipe instruction slots per counted operation (an IMAD / IMAD.WIDE occupies the pipe 2 clk per warp, an IMAD.HI 4):
    # 64x64->128 MAC = 4 IMAD.WIDE; 64-bit Harvey butterfly = mulhi64 (4) + two mullo64 (3 each) = 10; 32-bit butterfly = IMAD.HI (2) + 2 IMAD = 4
    IMAD_SLOTS = {"mac": 4.0, "bfly64": 10.0, "bfly32": 4.0}
    kernel_ms, classes = {}, {}
    for name, (launches_c, ms_c) in prof.items():
        if not launches_c:
            continue
        byts, ops = work.get(name, (0.0, 0.0))
        sec = ms_c / 1000.0
        ent = {"launches_per_step": launches_c / args.steps, "ms_per_step": ms_c / args.steps,
               "share_of_step": ms_c / ms_total, "alg_gb_per_step": byts / args.steps / 1e9,
               "hbm_gbs": byts / sec / 1e9 if sec else None, "hbm_frac": byts / sec / 1e9 / hbm_peak if sec else None}
        if name == "weighted_sum_tc_i8":
            ent["int8_tops"] = 2 * ops / sec / 1e12
            ent["tensor_frac"] = ops / sec / umma_rate                       # vs the UMMA kind::i8 rate measured in this run
            ent["tensor_frac_vs_2x_bf16"] = ent["int8_tops"] / (2 * bf16_peak)   # round 1's derived denominator, kept for comparison
        elif name in MAC_CLASSES:
            ent["gmac_s"] = ops / sec / 1e9
            ent["int_pipe_frac"] = IMAD_SLOTS["mac"] * ops / sec / wide_rate
            ent["int_pipe_frac_vs_mac_chain"] = ops / sec / probe_rate
        elif name in BFLY_CLASSES:
            ent["gbutterfly_s"] = ops / sec / 1e9
            ent["int_pipe_frac"] = IMAD_SLOTS["bfly64"] * ops / sec / wide_rate
            ent["int_pipe_frac_vs_mac_chain"] = BFLY_MACS * ops / sec / probe_rate
        elif name == "relinearize_u32":
            # relinearize through 30-bit auxiliary primes (relin32.cuh): ops = 32-bit Harvey butterflies (1 IMAD.HI + 2 IMAD)
            ent["gbutterfly32_s"] = ops / sec / 1e9
            ent["int_pipe_frac"] = IMAD_SLOTS["bfly32"] * ops / sec / wide_rate
            ent["int_pipe_frac_vs_mac_chain"] = 0.75 * ops / sec / probe_rate
        if name in counters:
            ent["ncu"] = counters[name]      # pipe_tensor / pipe_fma / issue_active ... of the class's main kernel (committed capture)
        classes[name] = ent
        kernel_ms[name] = {"launches_per_step": ent["launches_per_step"], "ms_per_step": ent["ms_per_step"]}
    dom = max(classes.items(), key=lambda kv: kv[1]["ms_per_step"])
    dname, d = dom
    avg_launch_ms = d["ms_per_step"] / d["launches_per_step"]
    if dname == "weighted_sum_tc_i8":
        roofline = {"kernel": dname, "bound": "tensor", "achieved": d["int8_tops"], "peak": 2 * umma_rate / 1e12, "unit": "TFLOP/s",
                    "frac": d["tensor_frac"],
                    "peak_source": "tcgen05.mma kind::i8 M128xN256xK32 issued back to back from shared-memory operands on all SMs, measured in "
                                   "this run (crcnn_probe_pipe which=2); achieved counts 2 ops per int8 multiply-accumulate of the unpadded GEMM",
                    "frac_vs_2x_measured_bf16": d["tensor_frac_vs_2x_bf16"]}
    else:
        roofline = {"kernel": dname, "bound": "hbm", "achieved": d["hbm_gbs"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": d["hbm_frac"], "peak_source": peak_src}
        if "int_pipe_frac" in d:
            roofline["int_pipe"] = {"frac": d["int_pipe_frac"], "imad_wide_ginstr_s": wide_rate / 1e9,
                                    "note": "the binding resource of this kernel: the IMAD pipe; peak = independent IMAD.WIDE.U32 chains "
                                            "measured in this run; counted IMAD-pipe slots per operation: %r" % (IMAD_SLOTS,)}
    roofline.update({"traffic": traffic.get(dname), "avg_launch_ms": avg_launch_ms, "launches_per_step": d["launches_per_step"],
                     "share_of_step": d["share_of_step"], "alg_bytes_per_launch": d["alg_gb_per_step"] * 1e9 / d["launches_per_step"],
                     "int_pipe_probe_gmac_s": probe_rate / 1e9, "imad_wide_probe_ginstr_s": wide_rate / 1e9,
                     "umma_i8_probe_tops": 2 * umma_rate / 1e12, "classes": classes})
    launches = int(sum(v[0] for v in prof.values()))

    line = {
        "metric": "encrypted MNIST images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic: uniform residues in SEAL ciphertext layout; weights = PlainModel.h5",
        "config": {"workload": ("PlainModel.h5 encoded net (conv-avgpool-bn-conv-square-avgpool-bn-fc-fc), n=8192, K=4, t=2^30, 32x32 zero-bordered input"
                                if (MODEL, N_POLY) == ("PlainModel", 8192) else "%s encoded net, n=%d, K=%d, t=2^%d" % (MODEL, N_POLY, K, T_PLAIN.bit_length() - 1)),
                   "images_per_step_per_gpu": B, "parallelism": "image replicas x%d (no collective)" % world,
                   "l2": "inputs (%.1f GB per step) and weights exceed the 126 MB L2; no flush needed" % (h2d / 1e9),
                   "weights": "conv1/conv2/fc4: byte planes of the NTT-form plaintexts resident (limb-split tcgen05 kind::i8 weighted sum in the NTT domain); fc3: ternary tap matrix resident (tcgen05 kind::i8 weighted sum in the coefficient domain)"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "h2d_link_gbs": h2d_gbs,
                "upload_ms": [round(a_.elapsed_time(b_), 1) for a_, b_ in up_ev],
                "forward_ms": [round(a_.elapsed_time(b_), 1) for a_, b_, _ in fwd_ev],
                "download_ms": [round(b_.elapsed_time(c_), 1) for _, b_, c_ in fwd_ev],
                "device_mem_free_gb": {"after_resident_loop": round(mem_free[0], 1), "after_e2e_loop": round(mem_free[1], 1)},
                "note": "H2D + re-stride of step s+1 overlap the forward of step s on a copy stream; the first upload is not overlapped"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "per_layer_ms": per_layer, "kernel_ms": kernel_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            res = run_reference_sample(args.cpu_budget_s, np.random.default_rng(0))
            if res is not None:
                ips, cores, sample, full = res
                line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "reference", "sample": sample,
                                        "per_layer_ms": {k: 1000 * s for k, s in full.items()}}
            else:
                line["cpu_baseline"] = port_baseline(args.cpu_budget_s)
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
    if rank == 0:
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def port_baseline(budget_s):
    """Fallback when oracle/_ref is absent: the plain-C oracle port, one thread, a few terms."""
    from oracle.port import Oracle
    o = Oracle(N_POLY, PRIMES, T_PLAIN)
    rng = np.random.default_rng(0)
    x = synth_residues(rng, (4, 2), PRIMES, N_POLY)
    w = o.encode_many(np.full(8, 0.37, np.float32)); b = o.encode_many(np.full(2, 0.1, np.float32))
    t0 = time.perf_counter()
    o.fc(x, 4, 2, w, b)
    c_term = (time.perf_counter() - t0) / 8
    terms = sum(nets.layer_terms(l) for l in nets.TOPOLOGIES[MODEL]["layers"])
    return {"value": 1.0 / (terms * c_term), "unit": "images/s", "cores": 1, "kind": "port",
            "sample": "oracle port, 8 weighted-sum terms at n=8192, scaled by the net's %d terms (pool/bn/square excluded)" % terms}


def main():
    global MODEL, N_POLY, PRIMES, T_PLAIN
    args = parse()
    if args.model:
        MODEL = args.model
    if args.n:
        from oracle.port import DEFAULT_PRIMES_128   # the table of SEAL's default primes only
        N_POLY, PRIMES = args.n, [int(q) for q in DEFAULT_PRIMES_128[args.n]]
        T_PLAIN = (1 << 18) if args.n == 4096 else (1 << 30)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        main_reference(args, rank, world)
    elif args.mode == "shard":
        main_shard(args, rank, world, local_rank)
    else:
        main_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
