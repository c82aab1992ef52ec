"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so, the plain-C CPU restatement
(fv_oracle.c) of the reference algorithm.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this; the product (crcnn_b200/) never does.

Same buffer conventions as oracle/ref.py (SEAL layout, stride n+1).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int)

# coeff_modulus_128(n) of SEAL 2.3.1 (SEAL/seal/util/globals.cpp:50-74 via defaultparams.h:22-26),
# pinned against the compiled reference in tests/test_oracle_vs_reference.py.
DEFAULT_PRIMES_128 = {
    2048: [0x3fffffff000001],
    4096: [0x7fffffff380001, 0x3fffffff000001],
    8192: [0x7fffffff380001, 0x7ffffffef00001, 0x3fffffff000001, 0x3ffffffef40001],
    16384: [0x7fffffff380001, 0x7ffffffef00001, 0x7ffffffeac0001, 0x7ffffffe700001,
            0x7ffffffe600001, 0x7ffffffe4c0001, 0x3fffffff000001, 0x3ffffffef40001],
}


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def _p(a, t=_u64p):
    return a.ctypes.data_as(t)


def load():
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, _u64p, C.c_uint64]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_ntt_table.restype = _u64p
    L.orc_ntt_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.orc_modulus.restype = C.c_uint64
    L.orc_modulus.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_minimal_root.restype = C.c_uint64
    L.orc_minimal_root.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.orc_bsk_count.argtypes = [C.c_void_p]
    L.orc_barrett_reduce_128.restype = C.c_uint64
    L.orc_barrett_reduce_128.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.orc_mulmod.restype = C.c_uint64
    L.orc_mulmod.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    L.orc_try_minimal_primitive_root.argtypes = [C.c_uint64, C.c_uint64, _u64p]
    L.orc_ntt_single.argtypes = [_u64p, C.c_int, C.c_uint64, C.c_int]
    L.orc_dyadic_product.argtypes = [_u64p, _u64p, C.c_int, C.c_uint64, _u64p]
    L.orc_multiply_poly_scalar.argtypes = [_u64p, C.c_int, C.c_uint64, C.c_uint64, _u64p]
    L.orc_encode_fractional.argtypes = [C.c_void_p, C.c_double, _u64p]
    L.orc_ct_transform.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, C.c_int]
    L.orc_plain_to_ntt.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_multiply_plain_ntt.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p]
    L.orc_plain_op.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p, C.c_int, C.c_int]
    L.orc_add_many.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p]
    L.orc_square.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p]
    L.orc_relinearize.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p, _i32p, C.c_int, _u64p]
    L.orc_conv_forward.argtypes = [C.c_void_p, _u64p] + [C.c_int] * 8 + [_u64p, _u64p, _u64p]
    L.orc_fc_forward.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, _u64p, _u64p, _u64p]
    L.orc_pool_forward.argtypes = [C.c_void_p, _u64p] + [C.c_int] * 7 + [_u64p, C.c_int, _u64p]
    L.orc_bn_forward.argtypes = [C.c_void_p, _u64p, C.c_int, C.c_int, C.c_int, _u64p, _u64p, _u64p]
    L.orc_square_forward.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p, _i32p, C.c_int, _u64p]
    L.orc_decrypt.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p, _u64p]
    L.orc_decode_fractional.restype = C.c_double
    L.orc_decode_fractional.argtypes = [C.c_void_p, _u64p]
    _i8p = C.POINTER(C.c_int8)
    L.orc_encrypt.argtypes = [C.c_void_p, _u64p, C.c_int, _u64p, _i8p, _i8p, _i8p, _u64p]
    return L


class Oracle:
    def __init__(self, n, primes, t):
        self.lib = load()
        self.n, self.K, self.t = n, len(primes), int(t)
        self.primes = [int(p) for p in primes]
        self.stride = n + 1
        pa = np.array(self.primes, dtype=np.uint64)
        self.h = self.lib.orc_create(n, self.K, _p(pa), self.t)
        if not self.h:
            raise RuntimeError("orc_create failed (bad parameters)")
        self.S = self.lib.orc_bsk_count(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def ct_words(self, size=2):
        return size * self.K * self.stride

    def ntt_table(self, base, idx, which):
        ptr = self.lib.orc_ntt_table(self.h, base, idx, which)
        return np.ctypeslib.as_array(ptr, shape=(self.n,)).copy()

    def modulus(self, base, idx):
        return int(self.lib.orc_modulus(self.h, base, idx))

    def minimal_root(self, base, idx):
        return int(self.lib.orc_minimal_root(self.h, base, idx))

    def encode(self, v):
        out = np.zeros(self.stride, dtype=np.uint64)
        cc = self.lib.orc_encode_fractional(self.h, float(v), _p(out))
        return out, cc

    def encode_many(self, vals):
        vals = np.asarray(vals, dtype=np.float32).ravel()
        out = np.zeros((len(vals), self.stride), dtype=np.uint64)
        for i, v in enumerate(vals):
            self.lib.orc_encode_fractional(self.h, float(v), _p(out[i]))
        return out

    def ct_transform(self, cts, size=2, inverse=False):
        a = np.array(cts, dtype=np.uint64, copy=True, order="C")
        self.lib.orc_ct_transform(self.h, _p(a), a.size // self.ct_words(size), size, int(inverse))
        return a

    def plain_to_ntt(self, plain, coeff_count=None):
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        cc = len(plain) if coeff_count is None else coeff_count
        out = np.zeros((self.K, self.stride), dtype=np.uint64)
        self.lib.orc_plain_to_ntt(self.h, _p(plain), cc, _p(out))
        return out

    def multiply_plain_ntt(self, cts, plain_ntt, size=2):
        a = np.array(cts, dtype=np.uint64, copy=True, order="C")
        p = np.ascontiguousarray(plain_ntt, dtype=np.uint64)
        self.lib.orc_multiply_plain_ntt(self.h, _p(a), a.size // self.ct_words(size), size, _p(p))
        return a

    def plain_op(self, cts, plain, op, coeff_count=None, size=2):
        a = np.array(cts, dtype=np.uint64, copy=True, order="C")
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        cc = len(p) if coeff_count is None else coeff_count
        self.lib.orc_plain_op(self.h, _p(a), a.size // self.ct_words(size), size, _p(p), cc,
                              {"mul": 0, "add": 1, "sub": 2}[op])
        return a

    def add_many(self, cts, size=2):
        a = np.ascontiguousarray(cts, dtype=np.uint64)
        out = np.zeros((size, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_add_many(self.h, _p(a), a.size // self.ct_words(size), size, _p(out))
        return out

    def square(self, cts):
        a = np.ascontiguousarray(cts, dtype=np.uint64)
        count = a.size // self.ct_words(2)
        out = np.zeros((count, 3, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_square(self.h, _p(a), count, _p(out))
        return out

    def relinearize(self, cts3, evk, sizes, dbc=16):
        a = np.ascontiguousarray(cts3, dtype=np.uint64)
        e = np.ascontiguousarray(evk, dtype=np.uint64)
        s = np.ascontiguousarray(sizes, dtype=np.int32)
        count = a.size // self.ct_words(3)
        out = np.zeros((count, 2, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_relinearize(self.h, _p(a), count, _p(e), _p(s, _i32p), dbc, _p(out))
        return out

    # ---- client-side steps of the re-encryption ----
    def decrypt(self, cts, sk_ntt):
        a = np.ascontiguousarray(cts, dtype=np.uint64)
        sk = np.ascontiguousarray(sk_ntt, dtype=np.uint64)
        count = a.size // self.ct_words(2)
        out = np.zeros((count, self.stride), dtype=np.uint64)
        self.lib.orc_decrypt(self.h, _p(a), count, _p(sk), _p(out))
        return out

    def decode(self, plain):
        p = np.zeros(self.stride, dtype=np.uint64)
        p[:len(plain)] = plain
        return float(self.lib.orc_decode_fractional(self.h, _p(p)))

    def encrypt(self, plain, pk, u, e0, e1, coeff_count=None):
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        k = np.ascontiguousarray(pk, dtype=np.uint64)
        i8 = C.POINTER(C.c_int8)
        uu, a0, a1 = (np.ascontiguousarray(v, dtype=np.int8) for v in (u, e0, e1))
        out = np.zeros((2, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_encrypt(self.h, _p(p), len(p) if coeff_count is None else coeff_count, _p(k), uu.ctypes.data_as(i8),
                             a0.ctypes.data_as(i8), a1.ctypes.data_as(i8), _p(out))
        return out

    # ---- layers (plaintext parameters as [...][n+1] coefficient-form words) ----
    def conv(self, x, xd, yd, zd, xs, ys, xf, yf, nf, w_plain, b_plain):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w_plain, dtype=np.uint64)
        b = np.ascontiguousarray(b_plain, dtype=np.uint64)
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        out = np.zeros((nf, xo, yo, 2, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_conv_forward(self.h, _p(x), xd, yd, zd, xs, ys, xf, yf, nf, _p(w), _p(b), _p(out))
        return out

    def fc(self, x, in_dim, out_dim, w_plain, b_plain):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w_plain, dtype=np.uint64)
        b = np.ascontiguousarray(b_plain, dtype=np.uint64)
        out = np.zeros((1, out_dim, 1, 2, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_fc_forward(self.h, _p(x), in_dim, out_dim, _p(w), _p(b), _p(out))
        return out

    def pool(self, x, xd, yd, zd, xs, ys, xf, yf, div_plain=None, div_cc=None):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        out = np.zeros((zd, xo, yo, 2, self.K, self.stride), dtype=np.uint64)
        if div_plain is None:
            self.lib.orc_pool_forward(self.h, _p(x), xd, yd, zd, xs, ys, xf, yf, None, 0, _p(out))
        else:
            d = np.ascontiguousarray(div_plain, dtype=np.uint64)
            self.lib.orc_pool_forward(self.h, _p(x), xd, yd, zd, xs, ys, xf, yf, _p(d),
                                      len(d) if div_cc is None else div_cc, _p(out))
        return out

    def bn(self, x, zd, xd, yd, mean_plain, invstd_plain):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        m = np.ascontiguousarray(mean_plain, dtype=np.uint64)
        v = np.ascontiguousarray(invstd_plain, dtype=np.uint64)
        out = np.zeros((zd, xd, yd, 2, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_bn_forward(self.h, _p(x), zd, xd, yd, _p(m), _p(v), _p(out))
        return out

    def square_layer(self, x, evk, sizes, dbc=16):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        e = np.ascontiguousarray(evk, dtype=np.uint64)
        s = np.ascontiguousarray(sizes, dtype=np.int32)
        count = x.size // self.ct_words(2)
        out = np.zeros((count, 2, self.K, self.stride), dtype=np.uint64)
        self.lib.orc_square_forward(self.h, _p(x), count, _p(e), _p(s, _i32p), dbc, _p(out))
        return out
