"""GPU: the HEADLINE configuration itself, byte-checked against the CPU oracle by sampling (oracle/sampled.py).

bench.py's number comes from the PlainModel.h5 network at n = 8192, K = 4, t = 2^30, batch 8.  These tests run exactly that code
path (nets.Network over the C ABI, default kernel selection) at exactly those shapes and compare randomly drawn output ciphertexts
of every layer -- plus the first / last / tile-boundary ones -- with the oracle evaluated on the same window of input ciphertexts,
which is what the reference computes for that output (CrCNN/src/convolutionalLayer.cpp:56-93, fullyConnectedLayer.cpp:113-168,
squareLayer.cpp:22-71 ...).  Covered at real size: the column-major limb-split GEMM with 25 column chunks and a partial last one
(conv1: 3,136 columns), fan-in 180 x 50 outputs (conv2, two 32-output tiles), the ternary-tap GEMM at fan-in 1250 x 500 outputs
x 8 images (fc3), square + relinearize over 14,400 ciphertexts with scratch chunking, both domains of pool / batch-norm, fc4.
Other networks: Tiny at n = 4096 (conv2: fan-in 800, 64 outputs; fc 1024 -> 512) and Approx at n = 8192 (fc 800 -> 500).
"""
import numpy as np
import pytest

import util
from oracle.port import Oracle
from oracle import sampled

pytestmark = pytest.mark.gpu


def _setup(model, n, B, seed):
    from crcnn_b200 import nets
    from crcnn_b200.lib import Engine
    primes, t = util.PRIMES[n], util.T_FOR_N[n]
    eng = Engine(n, primes, t)
    orc = Oracle(n, primes, t)
    rng = np.random.default_rng(seed)
    evk_host = util.random_evk(rng, n, primes)
    net = nets.Network(eng, model, evk=eng.evk_upload(*evk_host))
    zd, xd, yd = net.input_shape
    x = eng.upload(util.random_cts(rng, n, primes, B * zd * xd * yd))
    return eng, orc, net, x, evk_host


def _kernel_classes(eng):
    return {k for k, v in eng.prof().items() if v[0]}


def test_plainmodel_n8192_batch8_every_layer_sampled():
    """The bench configuration, all nine layers, 6+ sampled ciphertexts per layer."""
    B = 8
    eng, orc, net, x, evk_host = _setup("PlainModel", 8192, B, 42)
    eng.prof_reset()
    checked, _ = sampled.check_network(eng, orc, net, x, B, evk_host=evk_host, samples=6, seed=1, log=print)
    assert len(checked) == 9 and all(v >= 4 for v in checked.values()), checked
    # the kernels the bench times are the ones that ran: both tensor-core weighted sums, the 32-bit relinearisation
    ran = _kernel_classes(eng)
    assert {"weighted_sum_tcn_i8", "weighted_sum_tc_i8", "relinearize_u32", "behz_floor_sk"} <= ran, ran
    assert "weighted_sum_mac" not in ran, ran
    eng.close()


def test_plainmodel_n8192_batch1_every_layer_sampled():
    """Batch 1 (the reference's own calling convention): different tile counts, partial tiles everywhere."""
    eng, orc, net, x, evk_host = _setup("PlainModel", 8192, 1, 43)
    checked, _ = sampled.check_network(eng, orc, net, x, 1, evk_host=evk_host, samples=4, seed=2, log=print)
    assert len(checked) == 9
    eng.close()


def test_tiny_n4096_batch4_every_layer_sampled():
    """BASELINE config 1: conv2 has fan-in 800 and 64 outputs (many K blocks), fc 1024 -> 512."""
    B = 4
    eng, orc, net, x, evk_host = _setup("PlainModelTiny", 4096, B, 44)
    checked, _ = sampled.check_network(eng, orc, net, x, B, evk_host=evk_host, samples=5, seed=3, log=print)
    assert len(checked) == 6
    eng.close()


def test_approx_n8192_batch2_tail_sampled():
    """BASELINE config 3's network: the layers whose shapes differ from PlainModel (5x5 square, fc 800 -> 500), from a random
    activation of the right shape."""
    from crcnn_b200 import nets
    B = 2
    eng, orc, net, x, evk_host = _setup("ApproxPlainModel", 8192, B, 45)
    x.free()
    rng = np.random.default_rng(7)
    nin = nets.layer_io_counts(net.layers[3])[0]
    x = eng.upload(util.random_cts(rng, 8192, util.PRIMES[8192], B * nin))
    checked, _ = sampled.check_network(eng, orc, net, x, B, evk_host=evk_host, samples=4, seed=4, first=3, log=print)
    assert len(checked) == 6
    eng.close()


def test_tiny_n2048_reference_parameters_every_layer_sampled():
    """The parameters the reference itself ran the Tiny network with (n = 2048, one 54-bit prime, t = 2^16: Doc/Tesi.lyx:14789-14913)."""
    eng, orc, net, x, evk_host = _setup("PlainModelTiny", 2048, 2, 47)
    checked, _ = sampled.check_network(eng, orc, net, x, 2, evk_host=evk_host, samples=4, seed=5, log=print)
    assert len(checked) == 6
    eng.close()


def test_approx_n16384_batch1_tail_sampled():
    """The largest degree of the kernel sweep (n = 16384, K = 8, 9 Bsk primes, four 30-bit relinearisation primes): conv2 -> square ->
    pool -> bn -> fc -> fc of the Approx network at full layer shapes."""
    from crcnn_b200 import nets
    eng, orc, net, x, evk_host = _setup("ApproxPlainModel", 16384, 1, 48)
    x.free()
    rng = np.random.default_rng(8)
    nin = nets.layer_io_counts(net.layers[3])[0]
    x = eng.upload(util.random_cts(rng, 16384, util.PRIMES[16384], nin))
    checked, _ = sampled.check_network(eng, orc, net, x, 1, evk_host=evk_host, samples=3, seed=6, first=3, log=print)
    assert len(checked) == 6
    eng.close()


def test_checker_detects_a_wrong_ciphertext():
    """The sampled check is not vacuous: one flipped word in one sampled output fails it."""
    eng, orc, net, x, evk_host = _setup("PlainModel", 8192, 1, 46)
    y = net.forward_layer(0, x, 1)
    rng = np.random.default_rng(5)
    assert sampled.check_layer(eng, orc, net, 0, x, y, 1, rng, 3) >= 3
    bad = eng.download(y)
    bad[0, 1, 2, 77] ^= np.uint64(1)          # ciphertext 0 is always sampled
    yb = eng.upload(bad)
    with pytest.raises(AssertionError, match="differs from the oracle"):
        sampled.check_layer(eng, orc, net, 0, x, yb, 1, np.random.default_rng(5), 3)
    eng.close()
