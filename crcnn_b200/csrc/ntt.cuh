// Negacyclic NTT / inverse NTT of whole limb-polynomials resident in shared memory.
//
// One CTA transforms one limb-polynomial (n residues of one prime).  The transform is the same
// map as the reference's ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey
// (SEAL/seal/util/smallntt.cpp:195-375, smallntt.h:210-258): Cooley-Tukey forward with
// bit-reversed output, Gentleman-Sande inverse with n^-1 folded into pre-halved inverse roots,
// using the reference's own table order (root_powers[bitrev(i)] = psi^i), so NTT-form data made by
// SEAL (evaluation keys, NTT plaintexts) is interchangeable.  Outputs are canonical [0, q).
//
// Schedule: log2(n) stages are grouped into passes of 4 or 3 stages done in registers
// (16 or 8 residues per thread); between passes the polynomial lives in shared memory with one
// pad word every 16 residues, which makes both the strided passes and the final contiguous pass
// bank-conflict free.  The first forward pass reads HBM directly (coalesced, stride n/16), the
// last inverse pass writes HBM directly; the contiguous end is staged through shared memory so
// every global access is a full-warp contiguous 256 B segment.
#pragma once
#include "modarith.cuh"
#include "params.h"

namespace crcnn {

__device__ __forceinline__ int ntt_pad(int i) { return i + (i >> 4); }

template <int LOGN>
struct NttPlan {
    static constexpr int N = 1 << LOGN;
    static constexpr int THREADS = N / 32;  // 2 groups of 16 (4 of 8) residues per thread and pass
    // CTAs per SM the shared-memory footprint allows (227 KB usable), capped so 85 registers per thread suffice
    static constexpr int MIN_CTAS = (227 * 1024) / ((N + N / 16) * 8) > 768 / THREADS ? 768 / THREADS : ((227 * 1024) / ((N + N / 16) * 8) < 1 ? 1 : (227 * 1024) / ((N + N / 16) * 8));
    static constexpr int PASSES = (LOGN + 3) / 4;
    static constexpr int WIDE = LOGN - 3 * PASSES;  // how many passes take 4 stages (the rest take 3)
    static constexpr int SMEM_WORDS = N + N / 16;
    // wide passes go last so the gaps seen by shared memory are >=32, 16 and 1 only (conflict free)
    __host__ __device__ static constexpr int bits(int pass) { return pass >= PASSES - WIDE ? 4 : 3; }
};

__device__ __forceinline__ void ct_butterfly(uint64_t &x, uint64_t &y, uint64_t W, uint64_t Wp, uint64_t q, uint64_t twoq) {
    // Harvey butterfly: x, y in [0,4q) -> [0,4q)
    uint64_t X = x >= twoq ? x - twoq : x;
    uint64_t Q = mulshoup_lazy(y, W, Wp, q);
    x = X + Q;
    y = X + twoq - Q;
}

__device__ __forceinline__ void gs_butterfly(uint64_t &u, uint64_t &v, uint64_t W, uint64_t Wp, uint64_t q, uint64_t twoq) {
    // u, v in [0,2q) -> [0,2q); W = (psi^-k)/2 so every stage also halves
    uint64_t T = u + twoq - v;
    uint64_t S = u + v;
    S = S >= twoq ? S - twoq : S;
    u = (S + ((S & 1) ? q : 0)) >> 1;
    v = mulshoup_lazy(T, W, Wp, q);
}

// B forward stages on 2^B residues spaced g apart; m0 = number of blocks at the first stage,
// blk = this group's block index at that stage.
template <int B>
__device__ __forceinline__ void fwd_group(uint64_t (&x)[1 << B], const uint64_t *__restrict__ w,
                                          const uint64_t *__restrict__ wp, uint64_t q, uint64_t twoq, int m0, int blk) {
#pragma unroll
    for (int s = 0; s < B; s++) {
        // stage s: pair distance 2^(B-1-s) (local index), 2^s local blocks each with its own twiddle
#pragma unroll
        for (int pr = 0; pr < (1 << (B - 1)); pr++) {
            const int lb = pr >> (B - 1 - s), a = pr & ((1 << (B - 1 - s)) - 1);
            const int ia = (lb << (B - s)) + a, ib = ia + (1 << (B - 1 - s));
            const int tw = ((m0 + blk) << s) + lb;
            ct_butterfly(x[ia], x[ib], __ldg(w + tw), __ldg(wp + tw), q, twoq);
        }
    }
}

// B inverse stages; h0 = n / (2 * distance of the first stage), blk = group index i.
template <int B>
__device__ __forceinline__ void inv_group(uint64_t (&x)[1 << B], const uint64_t *__restrict__ iw,
                                          const uint64_t *__restrict__ iwp, uint64_t q, uint64_t twoq, int h0, int blk) {
#pragma unroll
    for (int s = 0; s < B; s++) {
        // stage s: pair distance 2^s, 2^(B-1-s) local blocks
#pragma unroll
        for (int pr = 0; pr < (1 << (B - 1)); pr++) {
            const int lb = pr >> s, a = pr & ((1 << s) - 1);
            const int ia = (lb << (s + 1)) + a, ib = ia + (1 << s);
            const int tw = (h0 >> s) + (blk << (B - 1 - s)) + lb;
            gs_butterfly(x[ia], x[ib], __ldg(iw + tw), __ldg(iwp + tw), q, twoq);
        }
    }
}

// One forward pass over the whole polynomial.  SRC_GLOBAL: read `gsrc` (unpadded) instead of smem.
template <int LOGN, int B, bool SRC_GLOBAL>
__device__ __forceinline__ void fwd_pass(uint64_t *sm, const uint64_t *__restrict__ gsrc, const NttTable &tb, int m0, int g) {
    constexpr int N = 1 << LOGN;
    const uint64_t q = tb.mod.q, twoq = 2 * q;
    for (int G = threadIdx.x; G < (N >> B); G += blockDim.x) {
        int blk = G / g, o = G - blk * g;
        int base = blk * (g << B) + o;
        uint64_t x[1 << B];
#pragma unroll
        for (int k = 0; k < (1 << B); k++) x[k] = SRC_GLOBAL ? gsrc[base + k * g] : sm[ntt_pad(base + k * g)];
        fwd_group<B>(x, tb.w, tb.wp, q, twoq, m0, blk);
#pragma unroll
        for (int k = 0; k < (1 << B); k++) sm[ntt_pad(base + k * g)] = x[k];
    }
}

template <int LOGN, int B, bool DST_GLOBAL>
__device__ __forceinline__ void inv_pass(uint64_t *sm, uint64_t *__restrict__ gdst, const NttTable &tb, int g) {
    constexpr int N = 1 << LOGN;
    const uint64_t q = tb.mod.q, twoq = 2 * q;
    const int h0 = N / (2 * g);
    for (int G = threadIdx.x; G < (N >> B); G += blockDim.x) {
        int blk = G / g, o = G - blk * g;
        int base = blk * (g << B) + o;
        uint64_t x[1 << B];
#pragma unroll
        for (int k = 0; k < (1 << B); k++) x[k] = sm[ntt_pad(base + k * g)];
        inv_group<B>(x, tb.iw, tb.iwp, q, twoq, h0, blk);
#pragma unroll
        for (int k = 0; k < (1 << B); k++) {
            if (DST_GLOBAL) {
                uint64_t v = x[k];
                gdst[base + k * g] = v >= q ? v - q : v;
            } else {
                sm[ntt_pad(base + k * g)] = x[k];
            }
        }
    }
}

// Forward transform of one polynomial: src (global, n words, values < 4q) -> sm (padded, lazy [0,4q)).
template <int LOGN>
__device__ __forceinline__ void ntt_forward_to_smem(uint64_t *sm, const uint64_t *__restrict__ src, const NttTable &tb) {
    using P = NttPlan<LOGN>;
    int m0 = 1, g = P::N;
    g >>= P::bits(0);
    if (P::bits(0) == 4) fwd_pass<LOGN, 4, true>(sm, src, tb, m0, g); else fwd_pass<LOGN, 3, true>(sm, src, tb, m0, g);
    m0 <<= P::bits(0);
    __syncthreads();
#pragma unroll
    for (int p = 1; p < P::PASSES; p++) {
        g >>= P::bits(p);
        if (P::bits(p) == 4) fwd_pass<LOGN, 4, false>(sm, nullptr, tb, m0, g); else fwd_pass<LOGN, 3, false>(sm, nullptr, tb, m0, g);
        m0 <<= P::bits(p);
        __syncthreads();
    }
}

// Forward transform when the polynomial is already in padded shared memory (values < 4q).
template <int LOGN>
__device__ __forceinline__ void ntt_forward_in_smem(uint64_t *sm, const NttTable &tb) {
    using P = NttPlan<LOGN>;
    int m0 = 1, g = P::N;
#pragma unroll
    for (int p = 0; p < P::PASSES; p++) {
        g >>= P::bits(p);
        if (P::bits(p) == 4) fwd_pass<LOGN, 4, false>(sm, nullptr, tb, m0, g); else fwd_pass<LOGN, 3, false>(sm, nullptr, tb, m0, g);
        m0 <<= P::bits(p);
        __syncthreads();
    }
}

// Inverse transform of the polynomial in padded shared memory (values < 2q) -> dst (global, canonical).
// Pass order mirrors the forward plan (narrow passes first so the last, HBM-writing pass is strided).
template <int LOGN>
__device__ __forceinline__ void ntt_inverse_from_smem(uint64_t *sm, uint64_t *__restrict__ dst, const NttTable &tb) {
    using P = NttPlan<LOGN>;
    int g = 1;
#pragma unroll
    for (int p = P::PASSES - 1; p >= 1; p--) {
        if (P::bits(p) == 4) inv_pass<LOGN, 4, false>(sm, nullptr, tb, g); else inv_pass<LOGN, 3, false>(sm, nullptr, tb, g);
        g <<= P::bits(p);
        __syncthreads();
    }
    if (P::bits(0) == 4) inv_pass<LOGN, 4, true>(sm, dst, tb, g); else inv_pass<LOGN, 3, true>(sm, dst, tb, g);
}

// Inverse transform leaving the result (lazy, [0,2q)) in shared memory.
template <int LOGN>
__device__ __forceinline__ void ntt_inverse_in_smem(uint64_t *sm, const NttTable &tb) {
    using P = NttPlan<LOGN>;
    int g = 1;
#pragma unroll
    for (int p = P::PASSES - 1; p >= 0; p--) {
        if (P::bits(p) == 4) inv_pass<LOGN, 4, false>(sm, nullptr, tb, g); else inv_pass<LOGN, 3, false>(sm, nullptr, tb, g);
        g <<= P::bits(p);
        __syncthreads();
    }
}

// Coalesced copies between global (unpadded) and shared (padded).
template <int LOGN>
__device__ __forceinline__ void smem_load_poly(uint64_t *sm, const uint64_t *__restrict__ src) {
    for (int i = threadIdx.x; i < (1 << LOGN); i += blockDim.x) sm[ntt_pad(i)] = src[i];
}
template <int LOGN>
__device__ __forceinline__ void smem_store_poly_canonical(const uint64_t *sm, uint64_t *__restrict__ dst, uint64_t q) {
    // input lazy in [0,4q)
    const uint64_t twoq = 2 * q;
    for (int i = threadIdx.x; i < (1 << LOGN); i += blockDim.x) {
        uint64_t v = sm[ntt_pad(i)];
        v = v >= twoq ? v - twoq : v;
        dst[i] = v >= q ? v - q : v;
    }
}

}  // namespace crcnn
