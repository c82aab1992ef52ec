// Relinearisation 3 -> 2 through word-size auxiliary primes (sm_100a).
//
// What the reference computes (Evaluator::relinearize_one_step, SEAL/seal/evaluator.cpp:934-1069): for every
// coefficient prime q_j and output polynomial p in {0,1}
//     out_p[j] = c_p[j] + INTT_j( sum_d NTT_j(digit_d) (.) key_(d,p)[j] )   mod q_j ,
// digit_d = the d-th dbc-bit digit polynomial of c2_i * (q/q_i)^-1 mod q_i.  That is D*K forward and 2K inverse
// 64-bit transforms per ciphertext (72 at n = 8192) -- 25 % of the whole forward pass on B200.
//
// The sum is a negacyclic product sum over Z_(q_j); lifted to the integers (digits in [0, 2^dbc), key
// coefficients in [0, q_j)) every coefficient W of  sum_d digit_d (*) key_(d,p)  satisfies
//     |W| <= D * n * (2^dbc - 1) * (q_j - 1)            (< 2^89 at n = 8192, dbc = 16)
// so W is determined by its residues modulo three (four for n = 16384) NTT-friendly primes below 2^30 and
// out_p[j] = c_p[j] + (W mod q_j) -- the same canonical residue SEAL stores, bit for bit.  In the small fields
//   * a digit polynomial is transformed ONCE per small prime (it does not depend on j): D*S3 transforms,
//   * the key products accumulate in 64-bit words,
//   * 2K*S3 inverse transforms, then Garner's mixed-radix reconstruction folded with "mod q_j" and "+ c_p".
// Same count of transforms (72 at n = 8192) but on 32-bit words: a Harvey butterfly is one IMAD.HI + two IMAD + four
// ALU instructions instead of ~18, shared memory holds half the bytes, and the key-product stage reads 4 B words.
// The keys are converted once at upload (inverse 64-bit NTT, reduce mod p_s, forward 32-bit NTT, n^-1 folded in).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "params.h"

namespace crcnn {

constexpr int R32_MAXP = 4;

struct Relin32Consts {  // one device copy per key set (Relin32::dc); the kernels index its arrays dynamically
    int S3;             // auxiliary primes in use
    int D;              // total digits = sum_i digits_i
    int dbc;
    uint32_t p[R32_MAXP];
    uint64_t mu[R32_MAXP];                      // floor(2^64 / p_s)
    const uint2 *w[R32_MAXP];                   // (psi^bitrev(i), floor(. * 2^32 / p)) forward pairs, n entries
    const uint2 *iw[R32_MAXP];                  // (psi^-bitrev(i), companion) inverse pairs
    const uint2 *wl[R32_MAXP], *iwl[R32_MAXP];  // the pairs of the contiguous 5-stage pass, transposed (relin32.cu: fwd_last32)
    uint32_t ginv[R32_MAXP][R32_MAXP];          // p_k^-1 mod p_s for k < s (Garner), and Shoup companions
    uint32_t ginvp[R32_MAXP][R32_MAXP];
    uint32_t half[R32_MAXP];                    // (p_s - 1) / 2: mixed-radix digits of (P - 1) / 2
    uint64_t cmodq[MAXK][R32_MAXP];             // prod_(k<s) p_k mod q_j
    uint64_t cmodq_sh[MAXK][R32_MAXP];          // floor(cmodq * 2^64 / q_j)
    uint64_t Pmodq[MAXK];                       // P = prod p_s mod q_j
    unsigned char dprime[32], dshift[32];       // digit d = (scaled c2 of prime dprime[d] >> dshift[d]) & (2^dbc - 1)
    int dfirst[MAXK];                           // first digit of prime i
};

struct Relin32 {
    Relin32Consts c;
    uint32_t *keys = nullptr;    // [S3][D][2K][n], NTT form mod p_s, scaled by n^-1; output index o = p*K + j
    uint2 *tables = nullptr;     // 4 * S3 tables of n pairs
    Relin32Consts *dc = nullptr; // device copy of c (the kernels index its arrays dynamically)
};

// Is the word-size path exact for these parameters?  (digits below every auxiliary prime, bound below the product
// of at most R32_MAXP of them, n <= 16384.)  Fills c.S3 and the digit map on success.
bool relin32_applicable(int n, int K, const uint64_t *q, const int *digits, int dbc, Relin32Consts &c);

// Builds tables and converts the evaluation keys.  evk_dev: canonical NTT-form keys [sum sizes][K][n] on the device,
// key_off[i] = word offset of prime i's keys.  `c` must come from relin32_applicable.
cudaError_t relin32_build(const DeviceParams *dP, int logn, int K, const uint64_t *q, const uint64_t *evk_dev, const long *key_off,
                          long total_polys, Relin32 &r, cudaStream_t stream);
void relin32_free(Relin32 &r);

// scratch bytes one ciphertext needs: digit transforms + key-product accumulators + 16-bit digit planes
size_t relin32_scratch_bytes(int n, int K, const Relin32Consts &c);

// in3: [count][3][K][n] coefficient form; out: [count][2][K][n]
cudaError_t relin32_run(const DeviceParams *dP, int logn, int K, const Relin32 &r, const uint64_t *in3, uint64_t *out, long count,
                        void *scratch, cudaStream_t stream);

}  // namespace crcnn
