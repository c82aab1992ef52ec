"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libcrcnn_ref.so, the
UNMODIFIED reference (SEAL 2.3.1 + CrCNN layers) built by oracle/Makefile.ref.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (crcnn_b200/) never does.

Buffers are numpy uint64 arrays in SEAL layout: ciphertexts [count][size][K][n+1]
(SEAL/seal/ciphertext.h:448-452), plaintexts [coeff_count].
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcrcnn_ref.so")

_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)


def available():
    return os.path.exists(LIB_PATH)


def _p(a, t):
    return a.ctypes.data_as(t)


class Ref:
    """One process-global reference context (the reference keeps its SEAL objects in globals,
    CrCNN/src/globals.cpp:11-19)."""

    def __init__(self, n, t, seed=1, primes=None):
        if not available():
            raise RuntimeError("oracle/_ref/libcrcnn_ref.so missing: run `make -C oracle -f Makefile.ref`")
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.ref_last_error.restype = C.c_char_p
        L.ref_t.restype = C.c_uint64
        L.ref_evk.restype = C.c_long
        L.ref_init.argtypes = [C.c_int, C.c_uint64, C.c_uint64, _u64p, C.c_int]
        L.ref_evk.argtypes = [_u64p, _i32p]
        L.ref_encode.argtypes = [C.c_double, _u64p, _i32p]
        if primes is not None:
            pa = np.ascontiguousarray(primes, dtype=np.uint64)
            self._chk(L.ref_init(n, t, seed, _p(pa, _u64p), len(pa)))
        else:
            self._chk(L.ref_init(n, t, seed, None, 0))
        self.n = L.ref_n()
        self.K = L.ref_K()
        self.t = int(L.ref_t())
        pr = np.zeros(self.K, dtype=np.uint64)
        L.ref_primes(_p(pr, _u64p))
        self.primes = [int(x) for x in pr]
        self.stride = self.n + 1

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("reference error: " + self.lib.ref_last_error().decode())

    def ct_words(self, size=2):
        return size * self.K * self.stride

    # ---- keys / client side ----
    def evk(self):
        sizes = np.zeros(self.K, dtype=np.int32)
        words = self.lib.ref_evk(None, _p(sizes, _i32p))
        out = np.zeros(words, dtype=np.uint64)
        self.lib.ref_evk(_p(out, _u64p), _p(sizes, _i32p))
        return out, [int(s) for s in sizes], int(self.lib.ref_evk_dbc())

    def set_evk(self, words, sizes, dbc=16):
        """Replace the reference's evaluation keys by caller-supplied key material."""
        w = np.ascontiguousarray(words, dtype=np.uint64)
        s = np.ascontiguousarray(sizes, dtype=np.int32)
        self._chk(self.lib.ref_set_evk(_p(w, _u64p), _p(s, _i32p), dbc))

    def encode(self, v):
        out = np.zeros(self.stride, dtype=np.uint64)
        cc = C.c_int(0)
        self._chk(self.lib.ref_encode(float(v), _p(out, _u64p), C.byref(cc)))
        return out, cc.value

    def encode_many(self, vals):
        vals = np.asarray(vals, dtype=np.float32).ravel()
        out = np.zeros((len(vals), self.stride), dtype=np.uint64)
        for i, v in enumerate(vals):
            out[i], _ = self.encode(float(v))
        return out

    def encrypt(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float32).ravel()
        out = np.zeros((len(vals), 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_encrypt_values(_p(vals, _f32p), len(vals), _p(out, _u64p)))
        return out

    def decrypt(self, cts, size=2, want_plain=False):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        count = cts.size // self.ct_words(size)
        vals = np.zeros(count, dtype=np.float64)
        budgets = np.zeros(count, dtype=np.int32)
        plain = np.zeros((count, self.stride), dtype=np.uint64) if want_plain else None
        self._chk(self.lib.ref_decrypt_values(_p(cts, _u64p), count, size, _p(vals, _f64p), _p(budgets, _i32p),
                                              _p(plain, _u64p) if want_plain else None))
        return (vals, budgets, plain) if want_plain else (vals, budgets)

    def keys(self):
        """(secret key in NTT form [K][n+1], public key [2][K][n+1] in NTT form) of the reference's key generator."""
        sk = np.zeros((self.K, self.stride), dtype=np.uint64)
        pk = np.zeros((2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_keys(_p(sk, _u64p), _p(pk, _u64p)))
        return sk, pk

    def reencode(self, ct):
        """decrypt -> decode to float -> encode: the plaintext Network::forward re-encrypts (network.cpp:30-33). -> (plain n+1 words, value)"""
        c = np.ascontiguousarray(ct, dtype=np.uint64)
        out = np.zeros(self.stride, dtype=np.uint64)
        v = C.c_double()
        self._chk(self.lib.ref_reencode(_p(c, _u64p), _p(out, _u64p), C.byref(v)))
        return out, v.value

    # ---- evaluator-level ----
    def ct_transform(self, cts, size=2, inverse=False):
        a = np.array(cts, dtype=np.uint64, copy=True, order="C")
        self._chk(self.lib.ref_ct_transform(_p(a, _u64p), a.size // self.ct_words(size), size, int(inverse)))
        return a

    def plain_to_ntt(self, plain, coeff_count=None):
        plain = np.ascontiguousarray(plain, dtype=np.uint64)
        cc = len(plain) if coeff_count is None else coeff_count
        out = np.zeros((self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_plain_to_ntt(_p(plain, _u64p), cc, _p(out, _u64p)))
        return out

    def multiply_plain_ntt(self, cts, plain_ntt, size=2):
        a = np.array(cts, dtype=np.uint64, copy=True, order="C")
        p = np.ascontiguousarray(plain_ntt, dtype=np.uint64)
        self._chk(self.lib.ref_multiply_plain_ntt(_p(a, _u64p), a.size // self.ct_words(size), size, _p(p, _u64p)))
        return a

    def plain_op(self, cts, plain, op, coeff_count=None, size=2):
        """op: 'mul' | 'add' | 'sub' -> Evaluator::multiply_plain / add_plain / sub_plain."""
        a = np.array(cts, dtype=np.uint64, copy=True, order="C")
        p = np.ascontiguousarray(plain, dtype=np.uint64)
        cc = len(p) if coeff_count is None else coeff_count
        code = {"mul": 0, "add": 1, "sub": 2}[op]
        self._chk(self.lib.ref_plain_op(_p(a, _u64p), a.size // self.ct_words(size), size, _p(p, _u64p), cc, code))
        return a

    def add_many(self, cts, size=2):
        a = np.ascontiguousarray(cts, dtype=np.uint64)
        out = np.zeros((size, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_add_many(_p(a, _u64p), a.size // self.ct_words(size), size, _p(out, _u64p)))
        return out

    def square(self, cts):
        a = np.ascontiguousarray(cts, dtype=np.uint64)
        count = a.size // self.ct_words(2)
        out = np.zeros((count, 3, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_square(_p(a, _u64p), count, _p(out, _u64p)))
        return out

    def relinearize(self, cts3):
        a = np.ascontiguousarray(cts3, dtype=np.uint64)
        count = a.size // self.ct_words(3)
        out = np.zeros((count, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_relinearize(_p(a, _u64p), count, _p(out, _u64p)))
        return out

    # ---- layers ----
    def conv(self, x, xd, yd, zd, xs, ys, xf, yf, nf, w, b, th=8):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        out = np.zeros((nf, xo, yo, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_conv_forward(_p(x, _u64p), xd, yd, zd, xs, ys, xf, yf, nf, th,
                                            _p(w, _f32p), _p(b, _f32p), _p(out, _u64p)))
        return out

    def fc(self, x, in_dim, out_dim, w, b, th=8):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        out = np.zeros((1, out_dim, 1, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_fc_forward(_p(x, _u64p), in_dim, out_dim, th, _p(w, _f32p), _p(b, _f32p),
                                          _p(out, _u64p)))
        return out

    def conv_timed(self, x, xd, yd, zd, xs, ys, xf, yf, nf, w, b, th=8, reps=1):
        """(encode_s, first_forward_s, steady_forward_s): ONE layer object, weights transformed in its first forward
        (convolutionalLayer.cpp:149-168) and reused by the `reps` later ones -- what the reference pays per image."""
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        times = np.zeros(3, dtype=np.float64)
        self.lib.ref_conv_forward_timed.argtypes = [_u64p] + [C.c_int] * 9 + [_f32p, _f32p, C.c_int, _f64p, _u64p]
        self._chk(self.lib.ref_conv_forward_timed(_p(x, _u64p), xd, yd, zd, xs, ys, xf, yf, nf, th, _p(w, _f32p), _p(b, _f32p),
                                                  reps, _p(times, _f64p), None))
        return tuple(float(t) for t in times)

    def fc_timed(self, x, in_dim, out_dim, w, b, th=8, reps=1):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        times = np.zeros(3, dtype=np.float64)
        self.lib.ref_fc_forward_timed.argtypes = [_u64p, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _f64p, _u64p]
        self._chk(self.lib.ref_fc_forward_timed(_p(x, _u64p), in_dim, out_dim, th, _p(w, _f32p), _p(b, _f32p), reps,
                                                _p(times, _f64p), None))
        return tuple(float(t) for t in times)

    def fc3d(self, x, zd, xd, yd, out_dim, w, b, th=8):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        w = np.ascontiguousarray(w, dtype=np.float32).ravel()
        b = np.ascontiguousarray(b, dtype=np.float32).ravel()
        out = np.zeros((1, out_dim, 1, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_fc_forward_3d(_p(x, _u64p), zd, xd, yd, out_dim, th, _p(w, _f32p), _p(b, _f32p),
                                             _p(out, _u64p)))
        return out

    def pool(self, x, xd, yd, zd, xs, ys, xf, yf, avg=False):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        xo, yo = (xd - xf) // xs + 1, (yd - yf) // ys + 1
        out = np.zeros((zd, xo, yo, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_pool_forward(_p(x, _u64p), xd, yd, zd, xs, ys, xf, yf, int(avg), _p(out, _u64p)))
        return out

    def bn(self, x, zd, xd, yd, mean, invstd):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        mean = np.ascontiguousarray(mean, dtype=np.float32)
        invstd = np.ascontiguousarray(invstd, dtype=np.float32)
        out = np.zeros((zd, xd, yd, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_bn_forward(_p(x, _u64p), zd, xd, yd, _p(mean, _f32p), _p(invstd, _f32p),
                                          _p(out, _u64p)))
        return out

    def square_layer(self, x, zd, xd, yd, th=8):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        out = np.zeros((zd, xd, yd, 2, self.K, self.stride), dtype=np.uint64)
        self._chk(self.lib.ref_square_forward(_p(x, _u64p), zd, xd, yd, th, _p(out, _u64p)))
        return out
