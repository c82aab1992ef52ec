"""Pins the CPU oracle (oracle/fv_oracle.c) to SEAL 2.3.1's own known-answer vectors.

Every expected value below is copied from the reference's unit tests (SEALTest/util/*.cpp, cited
per test); none was produced by our code.
"""
import ctypes as C

import numpy as np

from oracle import port

L = port.load()


def test_barrett_reduce_128_kat():
    # SEALTest/util/uintarithsmallmod.cpp:143-187
    cases = [
        (2, 0, 0, 0), (2, 1, 0, 1), (2, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 1),
        (3, 0, 0, 0), (3, 1, 0, 1), (3, 123, 456, 0), (3, 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF, 0),
        (13131313131313, 0, 0, 0), (13131313131313, 1, 0, 1), (13131313131313, 123, 456, 8722750765283),
        (13131313131313, 24242424242424, 79797979797979, 1010101010101),
    ]
    for q, lo, hi, want in cases:
        assert L.orc_barrett_reduce_128(lo, hi, q) == want, (q, lo, hi)


def test_multiply_uint_uint_mod_kat():
    # SEALTest/util/uintarithsmallmod.cpp:187-202
    for q, a, b, want in [(2, 0, 0, 0), (2, 0, 1, 0), (2, 1, 0, 0), (2, 1, 1, 1),
                          (10, 0, 0, 0), (10, 0, 1, 0), (10, 1, 0, 0), (10, 1, 1, 1), (10, 7, 7, 9)]:
        assert L.orc_mulmod(a, b, q) == want


def test_try_minimal_primitive_root_kat():
    # SEALTest/util/uintarithsmallmod.cpp:356-375
    r = C.c_uint64()
    for q, degree, want in [(11, 2, 10), (29, 2, 28), (29, 4, 12), (1234565441, 2, 1234565440), (1234565441, 8, 249725733)]:
        assert L.orc_try_minimal_primitive_root(degree, q, C.byref(r)) == 1
        assert r.value == want, (q, degree)


def test_ntt_root_powers_kat():
    # SEALTest/util/smallntt.cpp:52-72: root_powers for q = 0xffffffffffc0001, n = 2 and n = 4
    q = 0xffffffffffc0001
    r = C.c_uint64()
    assert L.orc_try_minimal_primitive_root(4, q, C.byref(r)) == 1
    assert r.value == 288794978602139552
    # n = 4: root_powers[bitrev(i)] = psi^i with psi the minimal primitive 8th root
    assert L.orc_try_minimal_primitive_root(8, q, C.byref(r)) == 1
    psi = r.value
    pw = [pow(psi, i, q) for i in range(4)]
    table = [pw[0], pw[2], pw[1], pw[3]]  # bit-reversed order
    assert table == [1, 288794978602139552, 178930308976060547, 748001537669050592]


def test_negacyclic_ntt_kat():
    # SEALTest/util/smallntt.cpp:83-100
    q = 0xffffffffffc0001
    for inp, want in [([0, 0], [0, 0]), ([1, 0], [1, 1]), ([1, 1], [288794978602139553, 864126526004445282])]:
        a = np.array(inp, dtype=np.uint64)
        assert L.orc_ntt_single(a.ctypes.data_as(port._u64p), 1, q, 0) == 1
        assert [int(x) for x in a] == want


def test_inverse_ntt_roundtrip():
    # SEALTest/util/smallntt.cpp:103-134 (random data, n = 8)
    q = 0xffffffffffc0001
    rng = np.random.default_rng(0)
    for logn in (3, 10, 12):
        a = rng.integers(0, q, size=1 << logn, dtype=np.uint64)
        b = a.copy()
        L.orc_ntt_single(b.ctypes.data_as(port._u64p), logn, q, 0)
        assert not np.array_equal(a, b)
        L.orc_ntt_single(b.ctypes.data_as(port._u64p), logn, q, 1)
        assert np.array_equal(a, b)


def test_dyadic_and_scalar_kat():
    # SEALTest/util/polyarithsmallmod.cpp:262-290 and :94-108
    a = np.array([1, 1, 1], dtype=np.uint64); b = np.array([2, 3, 4], dtype=np.uint64); o = np.zeros(3, dtype=np.uint64)
    L.orc_dyadic_product(a.ctypes.data_as(port._u64p), b.ctypes.data_as(port._u64p), 3, 13, o.ctypes.data_as(port._u64p))
    assert list(o) == [2, 3, 4]
    a = np.array([0, 0, 0], dtype=np.uint64)
    L.orc_dyadic_product(a.ctypes.data_as(port._u64p), b.ctypes.data_as(port._u64p), 3, 13, o.ctypes.data_as(port._u64p))
    assert list(o) == [0, 0, 0]
    p = np.array([1, 3, 4], dtype=np.uint64)
    L.orc_multiply_poly_scalar(p.ctypes.data_as(port._u64p), 3, 3, 5, p.ctypes.data_as(port._u64p))
    assert list(p) == [3, 4, 2]


def test_ntt_of_constant_plaintext_is_constant():
    # SEALTest/evaluator.cpp:957-990 (TransformPlainToNTT): NTT of constant c is all-c, pad word 0
    n = 2048
    o = port.Oracle(n, port.DEFAULT_PRIMES_128[n], 1 << 6)
    for c in (0, 1, 2):
        out = o.plain_to_ntt(np.array([c], dtype=np.uint64))
        assert (out[:, :n] == c).all() and (out[:, n] == 0).all()
