"""Output-neuron sharding across GPUs (BASELINE config 3) through the C++17 host path: crcnn_b200::ShardedNetwork
(crcnn_b200/cpp/crcnn_b200.hpp) over crcnn_comm_all_gather (NCCL point-to-point group on the context's stream).  Every conv / fc
layer is split by output channel / row exactly like the reference splits them over threads (CrCNN/src/convolutionalLayer.cpp:177-187,
fullyConnectedLayer.cpp:148-158); the sharded forward must give, on every rank, the bytes of the unsharded forward -- for one image
and for a batch, for even and for uneven splits (20 / 50 / 500 / 10 outputs over 2 and 3 ranks)."""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _worker(rank, world, out_dir, n, batch):
    import util
    from crcnn_b200 import host
    primes, t = util.PRIMES[n], util.T_FOR_N[n]
    idfile = os.path.join(out_dir, "nccl_id")
    if rank == 0:
        open(idfile + ".tmp", "wb").write(host.nccl_unique_id())
        os.rename(idfile + ".tmp", idfile)
    while not os.path.exists(idfile):
        time.sleep(0.05)
    nccl_id = open(idfile, "rb").read()
    rng = np.random.default_rng(3)      # same inputs and keys on every rank
    evk = util.random_evk(rng, n, primes)
    x = util.random_cts(rng, n, primes, batch * 784)
    net = host.HostNetwork(n, primes, t, "ApproxPlainModel", device=rank, evk=evk, world=world, rank=rank, nccl_id=nccl_id)
    got, shape = net.forward(x, batch=batch)
    assert shape == (1, 10, 1)
    np.save(os.path.join(out_dir, "sharded%d.npy" % rank), got)
    # a second forward through the same communicator (steady state) and a segment that ends inside the sharded region
    again, _ = net.forward(x, batch=batch)
    assert np.array_equal(again, got)
    mid, mshape = net.forward(x, batch=batch, first=0, last=5)
    np.save(os.path.join(out_dir, "mid%d.npy" % rank), mid)
    net.close()
    if rank == 0:
        ref = host.HostNetwork(n, primes, t, "ApproxPlainModel", device=0, evk=evk)
        want, _ = ref.forward(x, batch=batch)
        wmid, wshape = ref.forward(x, batch=batch, first=0, last=5)
        assert wshape == mshape == (50, 5, 5)
        np.save(os.path.join(out_dir, "single.npy"), want)
        np.save(os.path.join(out_dir, "single_mid.npy"), wmid)
        ref.close()


@pytest.mark.parametrize("world,batch", [(2, 1), (2, 3), (3, 2)])
def test_neuron_sharded_forward_matches_single_gpu(tmp_path, world, batch):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run under gpurun --gpus %d)" % (world, max(2, world)))
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, str(tmp_path), 4096, batch), nprocs=world, join=True)
    want, wmid = np.load(tmp_path / "single.npy"), np.load(tmp_path / "single_mid.npy")
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / ("sharded%d.npy" % r)), want), "rank %d: sharded forward differs from the single-GPU forward" % r
        assert np.array_equal(np.load(tmp_path / ("mid%d.npy" % r)), wmid), "rank %d: sharded segment [0,5) differs" % r
