// 64-bit modular arithmetic primitives shared by every kernel (and, compiled for the host, by
// tests/host_arith_test.cpp).  All functions are exact; results that the reference stores are
// canonical residues in [0, q), so "bit-identical to SEAL" reduces to mathematical equality
// (SURVEY.md section 0 item 5).  Moduli are <= 61 bits (coefficient primes <= 60 bits,
// SEAL/seal/util/defines.h:20; Bsk primes 61 bits, SEAL/seal/util/globals.cpp:321-340).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CRCNN_HD __host__ __device__ __forceinline__
#else
#define CRCNN_HD inline
#endif

namespace crcnn {

struct Mod {
    uint64_t q;   // modulus
    uint64_t r0;  // floor(2^128 / q), low word   (same constant as SmallModulus::const_ratio,
    uint64_t r1;  //                   high word   SEAL/seal/smallmodulus.cpp:62-73)
};

CRCNN_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

struct U128 {
    uint64_t lo, hi;
};

CRCNN_HD U128 mul128(uint64_t a, uint64_t b) {
    U128 r;
    r.lo = a * b;
    r.hi = mulhi64(a, b);
    return r;
}

// acc += a*b (mod 2^128).  Callers bound the number of accumulated terms so this never wraps.
CRCNN_HD void mac128(U128 &acc, uint64_t a, uint64_t b) {
    uint64_t lo = a * b;
    uint64_t hi = mulhi64(a, b);
    acc.lo += lo;
    acc.hi += hi + (acc.lo < lo);
}

CRCNN_HD void add128_64(U128 &acc, uint64_t v) {
    acc.lo += v;
    acc.hi += (acc.lo < v);
}

// Barrett reduction of a 128-bit value to [0, q): q_est = floor(z * floor(2^128/q) / 2^128)
// is floor(z/q) or one less, so a single conditional subtraction suffices (same bound as
// barrett_reduce_128, SEAL/seal/util/uintarithsmallmod.h:137-176).
CRCNN_HD uint64_t barrett128(U128 z, const Mod &m) {
    // word 2 of the 256-bit product z * r  (only bits [128,192) are needed)
    uint64_t carry = mulhi64(z.lo, m.r0);
    uint64_t t_lo = z.lo * m.r1;
    uint64_t t_hi = mulhi64(z.lo, m.r1);
    uint64_t s1 = t_lo + carry;
    uint64_t c1 = t_hi + (s1 < carry);
    uint64_t u_lo = z.hi * m.r0;
    uint64_t u_hi = mulhi64(z.hi, m.r0);
    uint64_t s2 = s1 + u_lo;
    uint64_t c2 = u_hi + (s2 < u_lo);
    uint64_t qest = z.hi * m.r1 + c1 + c2;
    uint64_t r = z.lo - qest * m.q;
    return r >= m.q ? r - m.q : r;
}

CRCNN_HD uint64_t mulmod(uint64_t a, uint64_t b, const Mod &m) { return barrett128(mul128(a, b), m); }

CRCNN_HD uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) {
    uint64_t s = a + b;
    return s >= q ? s - q : s;
}
CRCNN_HD uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : a + q - b; }
CRCNN_HD uint64_t negmod(uint64_t a, uint64_t q) { return a ? q - a : 0; }

// Shoup / Harvey multiplication by a constant w with precomputed w' = floor(w * 2^64 / q):
// returns w*y mod q as a representative in [0, 2q) for any 64-bit y.
CRCNN_HD uint64_t mulshoup_lazy(uint64_t y, uint64_t w, uint64_t wp, uint64_t q) {
    uint64_t h = mulhi64(wp, y);
    return y * w - h * q;
}

}  // namespace crcnn
