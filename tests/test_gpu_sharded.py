"""Output-neuron sharding across 2 GPUs (BASELINE config 3): every conv / fc layer split by output
channel / row, an NCCL all-gather of the activation ciphertexts before each layer that consumes all channels.  The
sharded forward of the Approx network must give, on every rank, exactly the bytes the unsharded forward gives."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir, n):
    import torch
    import torch.distributed as dist
    import util
    from crcnn_b200 import nets
    from crcnn_b200.lib import Engine

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        primes, t = util.PRIMES[n], util.T_FOR_N[n]
        eng = Engine(n, primes, t, device=rank)
        rng = np.random.default_rng(3)      # same inputs and keys on every rank
        K = len(primes)

        def residues(count, size=2):
            a = np.zeros((count, size, K, n + 1), dtype=np.uint64)
            for j, q in enumerate(primes):
                a[:, :, j, :n] = rng.integers(0, q, size=(count, size, n), dtype=np.uint64)
            return a

        sizes = [2 * ((int(q).bit_length() + 15) // 16) for q in primes]
        evk_words = np.concatenate([residues(1, s).ravel() for s in sizes])
        x = residues(28 * 28)
        net = nets.ShardedNetwork(eng, "ApproxPlainModel", dist, evk=eng.evk_upload(evk_words, sizes, 16))
        y = net.forward(eng.upload(x))
        got = eng.download(y)
        assert got.shape[0] == 10
        np.save(os.path.join(out_dir, "sharded%d.npy" % rank), got)
        if rank == 0:
            ref = nets.Network(eng, "ApproxPlainModel", evk=eng.evk_upload(evk_words, sizes, 16))
            want = eng.download(ref.forward(eng.upload(x), batch=1))
            np.save(os.path.join(out_dir, "single.npy"), want)
        eng.close()
    finally:
        dist.destroy_process_group()


def test_neuron_sharded_forward_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path), 4096), nprocs=2, join=True)
    a, b = np.load(tmp_path / "sharded0.npy"), np.load(tmp_path / "sharded1.npy")
    want = np.load(tmp_path / "single.npy")
    assert np.array_equal(a, b), "ranks disagree after the final all-gather"
    assert np.array_equal(a, want), "sharded forward differs from the single-GPU forward"
