"""CPU: the algorithm behind crcnn_b200/csrc/relin32.cu, restated in numpy and checked against the oracle's relinearize
(itself pinned to the compiled reference).  What is checked is the mathematics the CUDA path relies on -- the integer
bound, the auxiliary primes, Garner's mixed-radix reconstruction with the sign taken from the digits of (P-1)/2, the
reduction mod q_j -- independently of any GPU: out_p[j] = c_p[j] + (W mod q_j) with W the integer negacyclic product sum
sum_d digit_d (*) key_(d,p) recovered from its residues modulo three primes below 2^30."""
import numpy as np

from util import PRIMES, T_FOR_N, random_cts, random_evk
from oracle.port import Oracle

AUX = [1073643521, 1073479681, 1073184769]     # k * 2^15 + 1 just below 2^30 (relin32.cu: kAuxPrimes)


def negacyclic_mod(a, b, p):
    """(a * b mod x^n + 1) mod p for int64 vectors with a < 2^16, b < 2^30: every partial sum stays below 2^63."""
    n = len(a)
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))           # < n * 2^46
    lo, hi = full[:n], np.concatenate([full[n:], [0]])
    return (lo - hi) % p


def test_auxiliary_prime_reconstruction_equals_relinearize():
    n = 2048
    primes, t = PRIMES[n], T_FOR_N[n]
    assert len(primes) == 1
    q = int(primes[0])
    orc = Oracle(n, primes, t)
    rng = np.random.default_rng(5)
    evk, sizes, dbc = random_evk(rng, n, primes)
    D = sizes[0] // 2
    x3 = random_cts(rng, n, primes, 2, size=3)
    # extreme operands as well: c2 = q-1 everywhere (largest digits)
    x3[1, 2, 0, :n] = q - 1
    want = orc.relinearize(x3, evk, sizes, dbc)
    # integer bound the three primes must cover: 2 |W| < P
    P = AUX[0] * AUX[1] * AUX[2]
    assert 2 * D * n * ((1 << dbc) - 1) * (q - 1) < P
    # keys in coefficient form (the CUDA path converts them once at upload)
    key_coef = orc.ct_transform(evk.reshape(1, sizes[0], 1, n + 1), size=sizes[0], inverse=True)[0, :, 0, :n]
    half = [(p - 1) // 2 for p in AUX]
    inv = {(s, k): pow(AUX[k], -1, AUX[s]) for s in range(3) for k in range(s)}
    for ct in range(x3.shape[0]):
        d = x3[ct, 2, 0, :n]                       # K = 1: (q/q_0)^-1 = 1
        digits = [((d >> np.uint64(dbc * k)) & np.uint64((1 << dbc) - 1)).astype(np.int64) for k in range(D)]
        for p_out in range(2):
            res = []
            for p in AUX:
                acc = np.zeros(n, dtype=np.int64)
                for k in range(D):
                    acc = (acc + negacyclic_mod(digits[k], (key_coef[2 * k + p_out] % np.uint64(p)).astype(np.int64), p)) % p
                res.append(acc)
            got = np.zeros(n, dtype=np.uint64)
            for e in range(n):
                a = []
                for s in range(3):                  # Garner: mixed-radix digits of W mod P
                    v = int(res[s][e])
                    for k in range(s):
                        v = (v - a[k]) * inv[(s, k)] % AUX[s]
                    a.append(v)
                w = a[0] + a[1] * AUX[0] + a[2] * AUX[0] * AUX[1]
                neg = (a[2], a[1], a[0]) > (half[2], half[1], half[0])     # lexicographic comparison with the digits of (P-1)/2
                assert neg == (w > (P - 1) // 2)
                if neg:
                    w -= P
                got[e] = (int(x3[ct, p_out, 0, e]) + w) % q
            assert np.array_equal(got, want[ct, p_out, 0, :n]), (ct, p_out)
            assert want[ct, p_out, 0, n] == 0
