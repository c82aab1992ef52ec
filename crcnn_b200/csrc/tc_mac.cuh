// Tensor-core realisation of the plaintext-weight weighted sum (conv / fully connected layers),
// sm_100a tcgen05 kind::i8 -- the "dense integer GEMM" option of BASELINE.json's north_star.
//
// Why this is exact.  Every CrCNN weight is FractionalEncoder(base 3) output (CrCNN/src/globals.cpp:52,
// SEAL/seal/encoder.cpp:1013-1076): for |w| < 1/2 the plaintext is  sum_{i=1..32} a_i x^(n-i)  with
// a_i in {0, 1, t-1}.  Evaluator::transform_to_ntt lifts t-1 to q_j-1 = -1 (evaluator.cpp:1465-1486), so
// multiply_plain by it is, in the COEFFICIENT domain and mod every q_j,
//       (ct (*) w)[c] = - sum_i s_i * X~[c + i],      s_i in {0,+1,-1},
// where X~ is the ciphertext polynomial extended by 32 negated wrap-around coefficients
// (x^n = -1).  A whole layer  Y_m = sum_r X_r (*) w_(m,r)  is therefore the integer GEMM
//       D[(m,i), c] = sum_r A[(m,i), r] * X~_r[c],     A = -s in {0,+1,-1}  (int8),
// followed by the diagonal sum  Y_m[c] = sum_i D[(m,i), c+i].  The 55-bit residues X~ are split into
// 7 (8) unsigned byte planes, each plane is its own u8 x s8 -> s32 GEMM (|D| <= 255*R < 2^31), the planes
// are recombined with shifts and reduced mod q_j once.  Residues are canonical, hence byte-identical to
// the reference's multiply_plain_ntt + transform_from_ntt + add_many; no NTT
// of inputs, weights or outputs is needed at all.
//
// Kernel shape (tc_mac_kernel): one CTA per SM, persistent over work items (4 outputs x one limb-polynomial
// of one output position).  A tile is 128 rows = 4 outputs x 32 taps by N = planes x 32 coefficients; the
// CTA walks the 257 coefficient blocks of its polynomial in order, carrying the half-finished diagonal sums.
//   warp 0      TMA producer: A (weights) and B (byte planes, K-major, 128B swizzle) into a 4-stage ring
//   warp 1      tcgen05.mma.cta_group::1.kind::i8 issuer, accumulators double-buffered in TMEM
//   warps 2-5   epilogue: tcgen05.ld -> skewed shared-memory transpose (the diagonal sum) -> plane
//               recombination -> Barrett -> coalesced 256 B stores
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "params.h"

namespace crcnn {

constexpr int TC_TAPS = 32;    // fractional digits of the encoder = taps per weight
constexpr int TC_BM = 128;     // GEMM rows per tile (4 outputs x 32 taps)
constexpr int TC_BK = 128;     // fan-in bytes per pipeline stage (= 4 UMMA K-steps of 32)
constexpr int TC_CB = 32;      // coefficients per N tile

// Byte planes the residues of this context need (7 for <= 56-bit primes, else 8).
inline int tc_planes_for(const DeviceParams &d) {
    int maxbits = 0;
    for (int j = 0; j < d.K; j++) {
        int b = 0;
        for (uint64_t v = d.tab[j].mod.q; v; v >>= 1) b++;
        maxbits = b > maxbits ? b : maxbits;
    }
    return maxbits <= 56 ? 7 : 8;
}

struct TcMacArgs {
    const int8_t *A;     // [Mpad][Kpad]: row m*32 + (i-1) holds -s_i of weight (m, r) at column r; zero padded
    uint8_t *B;          // scratch [npos*2*K][planes][n+32][Kpad]: byte planes of the gathered inputs, r contiguous
    const uint64_t *x;   // input ciphertexts [num_in][2][K][n], COEFFICIENT form
    const int *in_index; // [npos][R] input ciphertext of (local) output position p, term r
    const uint64_t *bias;// [Mtotal][K][n] Delta-scaled bias in coefficient form, added to poly 0 (may be null)
    uint64_t *out;       // output ct index = (pg / Pimg) * (Mtotal * Pimg) + (m0 + m) * Pimg + pg % Pimg, pg = p0 + p
    int M, Mpad, R, Kpad, planes;
    int npos, p0, Pimg, Mtotal, m0;   // npos positions of this launch, the first one being position p0 of the layer
    int n, K;
};

size_t tc_b_bytes(const TcMacArgs &a);                       // size of the B scratch
cudaError_t launch_umma_i8_probe(int blocks, int iters, int variant, double *macs, cudaStream_t stream);  // tensor-pipe roofline probe
cudaError_t tc_mac_available();                              // driver entry point for tensor maps resolved?
cudaError_t launch_tc_split(const DeviceParams *P, const TcMacArgs &a, cudaStream_t stream);
cudaError_t launch_tc_mac(const DeviceParams *P, const TcMacArgs &a, int sm_count, cudaStream_t stream);

}  // namespace crcnn
