"""BASELINE config 4: a large plaintext modulus T split into pairwise coprime t_1, t_2 -- one independent context, key set,
weight encoding and network instance per modulus (one per GPU when there are several), no collective; the host recombines
the DECRYPTED plaintext coefficients by CRT.  Not in the reference: each instance is an ordinary
context, so per-instance parity is pinned by the oracle like everything else; this test checks the harness end to end:
CRT(instance results) equals the plaintext the reference computes under the single modulus T = t_1 t_2."""
import numpy as np
import pytest

from util import have_ref

pytestmark = pytest.mark.gpu


def _crt(a1, t1, a2, t2):
    """x mod t1 t2 from (x mod t1, x mod t2), element-wise on python ints."""
    inv = pow(t1, -1, t2)
    return [(int(x1) + t1 * (((int(x2) - int(x1)) * inv) % t2)) % (t1 * t2) for x1, x2 in zip(a1, a2)]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_crt_instances_recombine_to_the_single_modulus_result():
    import torch
    from oracle.ref import Ref
    from crcnn_b200.lib import Engine
    n, t1, t2 = 8192, 32771, 32779            # primes near 2^15; T ~ 2^30
    T = t1 * t2
    rng = np.random.default_rng(9)
    img = rng.uniform(-0.4242, 2.8215, size=16).astype(np.float32)
    wv = rng.uniform(-1, 1, size=2 * 4).astype(np.float32)
    bv = rng.uniform(-1, 1, size=2).astype(np.float32)
    ngpu = max(1, torch.cuda.device_count())
    plains = []
    for i, t in enumerate((t1, t2)):
        r = Ref(n, t, seed=21 + i)             # the client of instance i: its own keys under modulus t_i
        eng = Engine(n, r.primes, t, device=i % ngpu)
        evk, sizes, dbc = r.evk()
        g = eng.conv(eng.upload(r.encrypt(img)), eng.plain_encode(wv), eng.plain_encode(bv), 1, 4, 4, 1, 1, 1, 2, 2, 2)
        g = eng.square_layer(g, eng.evk_upload(evk, sizes, dbc))
        vals, budgets, plain = r.decrypt(eng.download(g), want_plain=True)
        assert budgets.min() > 0
        plains.append(plain)
        eng.close()
    r = Ref(n, T, seed=33)                     # the reference itself under the single modulus
    want_ct = r.square_layer(r.conv(r.encrypt(img), 4, 4, 1, 1, 1, 2, 2, 2, wv, bv), 2, 3, 3)
    vals_T, budgets_T, plain_T = r.decrypt(want_ct, want_plain=True)
    assert budgets_T.min() > 0
    for k in range(plain_T.shape[0]):
        got = _crt(plains[0][k], t1, plains[1][k], t2)
        assert got == [int(v) for v in plain_T[k]], "ciphertext %d: CRT of the instances differs from the single-modulus plaintext" % k
