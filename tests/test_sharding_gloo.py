"""CPU, world_size 2, gloo: the host-side logic of the N>1 paths.

(1) Output-neuron sharding: each rank computes its slice of a conv and an fc layer (the CPU oracle stands
    in for the GPU), the slices are all-gathered as ciphertext words in the order ShardedNetwork uses, and
    the reassembled tensor must equal the unsharded layer byte for byte.
(2) Image replicas (what bench.py --gpus N does): ranks process disjoint images, no collective on the data
    path; the only communication is the max-reduction of the timing.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import util
    from crcnn_b200 import nets
    from oracle.port import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 2048
        primes, t = util.PRIMES[n], util.T_FOR_N[n]
        o = Oracle(n, primes, t)
        K = len(primes)
        words = 2 * K * (n + 1)

        def gather(mine, counts):
            """max-sized slots + compaction, exactly as ShardedNetwork.all_gather does on the device"""
            slot = max(counts) * words
            padded = torch.zeros(world * slot, dtype=torch.int64)
            local = torch.zeros(slot, dtype=torch.int64)
            local[:counts[rank] * words] = torch.from_numpy(mine.view(np.int64).ravel())
            dist.all_gather_into_tensor(padded, local)
            parts = [padded[r * slot:r * slot + c * words] for r, c in enumerate(counts)]
            return torch.cat(parts).numpy().view(np.uint64)

        # --- conv with 5 filters split 3 + 2, fc with 7 rows split 4 + 3
        x = util.det_cts(5, n, primes, 2 * 3 * 3)
        cw, cb = util.det_floats(1, 5 * 2 * 2 * 2), util.det_floats(2, 5)
        full = o.conv(x, 3, 3, 2, 1, 1, 2, 2, 5, o.encode_many(cw), o.encode_many(cb))  # [5][2][2]
        k0, kc = nets.shard_range(5, world, rank)
        per_filter = 2 * 2 * 2
        mine = o.conv(x, 3, 3, 2, 1, 1, 2, 2, kc, o.encode_many(cw[k0 * per_filter:(k0 + kc) * per_filter]), o.encode_many(cb[k0:k0 + kc]))
        counts = nets.gather_counts(5, world, 4)
        got = gather(mine, counts).reshape(full.shape)
        assert np.array_equal(got, full), "conv shards do not reassemble"
        fw, fb = util.det_floats(3, 7 * 20), util.det_floats(4, 7)
        xin = got.reshape(20, 2, K, n + 1)
        full_fc = o.fc(xin, 20, 7, o.encode_many(fw), o.encode_many(fb))
        r0, rc = nets.shard_range(7, world, rank)
        mine = o.fc(xin, 20, rc, o.encode_many(fw[r0 * 20:(r0 + rc) * 20]), o.encode_many(fb[r0:r0 + rc]))
        counts = nets.gather_counts(7, world, 1)
        got = gather(mine, counts).reshape(full_fc.shape)
        assert np.array_equal(got, full_fc), "fc shards do not reassemble"
        # --- replicas: disjoint images, timing reduced with MAX like bench.py
        img = util.det_cts(100 + rank, n, primes, 4)
        res = o.pool(img, 2, 2, 1, 1, 1, 2, 2)
        tms = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        assert float(tms[0]) == 10.0 + world - 1
        open(os.path.join(out_dir, "ok%d" % rank), "w").write(util.sha(res))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    from crcnn_b200 import nets
    for total in (1, 7, 10, 20, 50, 500):
        for world in (1, 2, 3, 4, 8):
            spans = [nets.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (a, ca), (b, _) in zip(spans, spans[1:]):
                assert a + ca == b
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_gloo_shards_and_replicas(tmp_path):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = open(tmp_path / "ok0").read(), open(tmp_path / "ok1").read()
    assert a != b  # the replicas really processed different images
