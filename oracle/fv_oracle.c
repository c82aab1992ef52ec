/* TEST INFRASTRUCTURE ONLY -- see fv_oracle.h.  Plain C restatement of the reference algorithm
 * (SEAL 2.3.1 full-RNS FV evaluator + CrCNN layer forwards).  Not used by the product path.
 * All file:line citations are relative to /root/reference.  S/ = SEAL_2.3.1/SEAL/seal/,
 * SU/ = SEAL_2.3.1/SEAL/seal/util/, C/ = CrCNN/src/. */
#include "fv_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ---------------------------------------------------------------- small modulus */
typedef struct {
    uint64_t q;
    uint64_t r0, r1; /* const_ratio = floor(2^128 / q), S/smallmodulus.cpp:62-73 */
    int bits;
} smod;

static int bit_count(uint64_t v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

static smod smod_make(uint64_t q) {
    smod m; m.q = q; m.bits = bit_count(q);
    u128 all = ~(u128)0;
    u128 quo = all / q, rem = all % q;
    if (rem == (u128)(q - 1)) quo += 1; /* q | 2^128 (only m_tilde = 2^32) */
    m.r0 = (uint64_t)quo; m.r1 = (uint64_t)(quo >> 64);
    return m;
}

/* SU/uintarithsmallmod.h:137-176 -- word 2 of the 256-bit product z * const_ratio, then one
 * conditional subtraction. */
static inline uint64_t barrett128(uint64_t z0, uint64_t z1, const smod *m) {
    uint64_t carry = (uint64_t)(((u128)z0 * m->r0) >> 64);
    u128 t2 = (u128)z0 * m->r1;
    uint64_t tmp1 = (uint64_t)t2 + carry;
    uint64_t tmp3 = (uint64_t)(t2 >> 64) + (tmp1 < carry);
    t2 = (u128)z1 * m->r0;
    uint64_t s = tmp1 + (uint64_t)t2;
    carry = (uint64_t)(t2 >> 64) + (s < tmp1);
    tmp1 = z1 * m->r1 + tmp3 + carry;
    tmp3 = z0 - tmp1 * m->q;
    return tmp3 - (m->q & (uint64_t)(-(int64_t)(tmp3 >= m->q)));
}
static inline uint64_t barrett_u128(u128 z, const smod *m) { return barrett128((uint64_t)z, (uint64_t)(z >> 64), m); }
/* SU/uintarithsmallmod.h:178-190 */
static inline uint64_t mulmod(uint64_t a, uint64_t b, const smod *m) { return barrett_u128((u128)a * b, m); }
/* SU/uintarithsmallmod.h:92-114 */
static inline uint64_t addmod(uint64_t a, uint64_t b, const smod *m) { a += b; return a - (m->q & (uint64_t)(-(int64_t)(a >= m->q))); }
/* SU/uintarithsmallmod.h:116-135 */
static inline uint64_t submod(uint64_t a, uint64_t b, const smod *m) { uint64_t r = a - b; return r + (m->q & (uint64_t)(-(int64_t)(a < b))); }
/* SU/uintarithsmallmod.h:50-64 */
static inline uint64_t negmod(uint64_t a, const smod *m) { return a ? m->q - a : 0; }

static uint64_t powmod(uint64_t a, uint64_t e, const smod *m) {
    uint64_t r = 1 % m->q; a %= m->q;
    while (e) { if (e & 1) r = mulmod(r, a, m); a = mulmod(a, a, m); e >>= 1; }
    return r;
}
/* try_invert_uint_mod (extended Euclid); works for the non-prime m_tilde = 2^32 as well */
static uint64_t invmod(uint64_t a, uint64_t q) {
    __int128 t = 0, nt = 1; __int128 r = q, nr = a % q;
    while (nr != 0) { __int128 qu = r / nr; __int128 tmp = t - qu * nt; t = nt; nt = tmp; tmp = r - qu * nr; r = nr; nr = tmp; }
    if (t < 0) t += q;
    return (uint64_t)t;
}
/* SU/uintarithsmallmod.h:66-90 */
static uint64_t div2mod(uint64_t a, const smod *m) {
    if (a & 1) { u128 s = (u128)a + m->q; return (uint64_t)(s >> 1); }
    return a >> 1;
}

/* SU/uintarithsmallmod.cpp:83-108: smallest primitive degree-th root (degree = 2n).  SEAL starts from
 * a random primitive root and walks all odd powers keeping the minimum, so the result is
 * deterministic; we start from the first primitive root found by trial of g = 2,3,... */
int orc_try_minimal_primitive_root(uint64_t degree, uint64_t q, uint64_t *root_out) {
    if ((q - 1) % degree != 0) return 0;
    smod m = smod_make(q);
    uint64_t quo = (q - 1) / degree, root = 0;
    for (uint64_t g = 2; g < 4096; g++) {
        uint64_t r = powmod(g, quo, &m);
        if (r != 0 && powmod(r, degree >> 1, &m) == q - 1) { root = r; break; }
    }
    if (!root) return 0;
    uint64_t gsq = mulmod(root, root, &m), cur = root;
    for (uint64_t i = 0; i < degree; i++) {
        if (cur < root) root = cur;
        cur = mulmod(cur, gsq, &m);
    }
    *root_out = root;
    return 1;
}

uint64_t orc_barrett_reduce_128(uint64_t lo, uint64_t hi, uint64_t q) { smod m = smod_make(q); return barrett128(lo, hi, &m); }
uint64_t orc_mulmod(uint64_t a, uint64_t b, uint64_t q) { smod m = smod_make(q); return mulmod(a, b, &m); }

/* ---------------------------------------------------------------- NTT tables (SU/smallntt.cpp:37-92,162-184) */
typedef struct {
    smod m; int logn, n; uint64_t root;
    uint64_t *rp, *srp, *irp2, *sirp2;
} ntt_tab;

static uint32_t bitrev(uint32_t x, int bits) { uint32_t r = 0; for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

static void powers_bitrev(uint64_t root, const ntt_tab *t, uint64_t *dst) { /* SU/smallntt.cpp:162-172 */
    uint64_t prev = 1; dst[0] = 1;
    for (int i = 1; i < t->n; i++) { prev = mulmod(prev, root, &t->m); dst[bitrev((uint32_t)i, t->logn)] = prev; }
}
static void scale_powers(const uint64_t *in, const ntt_tab *t, uint64_t *dst) { /* SU/smallntt.cpp:175-184 */
    for (int i = 0; i < t->n; i++) dst[i] = (uint64_t)((((u128)in[i]) << 64) / t->m.q);
}
static int ntt_tab_make(ntt_tab *t, int logn, uint64_t q) {
    memset(t, 0, sizeof(*t));
    t->m = smod_make(q); t->logn = logn; t->n = 1 << logn;
    if (!orc_try_minimal_primitive_root((uint64_t)2 << logn, q, &t->root)) return 0;
    uint64_t inv_root = invmod(t->root, q);
    size_t bytes = (size_t)t->n * 8;
    t->rp = malloc(bytes); t->srp = malloc(bytes); t->irp2 = malloc(bytes); t->sirp2 = malloc(bytes);
    uint64_t *irp = malloc(bytes);
    powers_bitrev(t->root, t, t->rp); scale_powers(t->rp, t, t->srp);
    powers_bitrev(inv_root, t, irp);
    for (int i = 0; i < t->n; i++) t->irp2[i] = div2mod(irp[i], &t->m); /* SU/smallntt.cpp:76-79 */
    scale_powers(t->irp2, t, t->sirp2);
    free(irp);
    return 1;
}
static void ntt_tab_free(ntt_tab *t) { free(t->rp); free(t->srp); free(t->irp2); free(t->sirp2); }

/* SU/smallntt.cpp:195-273 -- Harvey CT butterflies, output in [0,4q), bit-reversed order */
static void ntt_lazy(uint64_t *a, const ntt_tab *tb) {
    uint64_t q = tb->m.q, two_q = 2 * q; int n = tb->n, t = n >> 1;
    for (int m = 1; m < n; m <<= 1) {
        for (int i = 0; i < m; i++) {
            int j1 = 2 * i * t, j2 = j1 + t;
            uint64_t W = tb->rp[m + i], Wp = tb->srp[m + i];
            for (int j = j1; j < j2; j++) {
                uint64_t X = a[j], Y = a[j + t];
                uint64_t cx = X - (two_q & (uint64_t)(-(int64_t)(X >= two_q)));
                uint64_t Q = (uint64_t)(((u128)Wp * Y) >> 64);
                Q = Y * W - Q * q;
                a[j] = cx + Q; a[j + t] = cx + (two_q - Q);
            }
        }
        t >>= 1;
    }
}
/* SU/smallntt.h:210-234 */
static void ntt_full(uint64_t *a, const ntt_tab *tb) {
    ntt_lazy(a, tb);
    uint64_t q = tb->m.q, two_q = 2 * q;
    for (int i = 0; i < tb->n; i++) {
        if (a[i] >= two_q) a[i] -= two_q;
        if (a[i] >= q) a[i] -= q;
    }
}
/* SU/smallntt.cpp:276-375 -- Gentleman-Sande with n^-1 folded in, output in [0,2q) */
static void intt_lazy(uint64_t *a, const ntt_tab *tb) {
    uint64_t q = tb->m.q, two_q = 2 * q; int n = tb->n, t = 1;
    for (int m = n; m > 1; m >>= 1) {
        int j1 = 0, h = m >> 1;
        for (int i = 0; i < h; i++) {
            int j2 = j1 + t;
            uint64_t W = tb->irp2[h + i], Wp = tb->sirp2[h + i];
            for (int j = j1; j < j2; j++) {
                uint64_t U = a[j], V = a[j + t];
                uint64_t T = two_q - V + U;
                uint64_t cu = U + V - (two_q & (uint64_t)(-(int64_t)((U << 1) >= T)));
                a[j] = (cu + (q & (uint64_t)(-(int64_t)(T & 1)))) >> 1;
                uint64_t H = (uint64_t)(((u128)Wp * T) >> 64);
                a[j + t] = T * W - H * q;
            }
            j1 += t << 1;
        }
        t <<= 1;
    }
}
/* SU/smallntt.h:239-258 */
static void intt_full(uint64_t *a, const ntt_tab *tb) {
    intt_lazy(a, tb);
    uint64_t q = tb->m.q;
    for (int i = 0; i < tb->n; i++) if (a[i] >= q) a[i] -= q;
}

int orc_ntt_single(uint64_t *poly, int logn, uint64_t q, int inverse) {
    ntt_tab t; if (!ntt_tab_make(&t, logn, q)) return 0;
    if (inverse) intt_full(poly, &t); else ntt_full(poly, &t);
    ntt_tab_free(&t); return 1;
}
/* SU/polyarithsmallmod.h:401-464 */
void orc_dyadic_product(const uint64_t *a, const uint64_t *b, int count, uint64_t q, uint64_t *out) {
    smod m = smod_make(q); for (int i = 0; i < count; i++) out[i] = mulmod(a[i], b[i], &m);
}
/* SU/polyarithsmallmod.h:173-232 */
void orc_multiply_poly_scalar(const uint64_t *a, int count, uint64_t s, uint64_t q, uint64_t *out) {
    smod m = smod_make(q); for (int i = 0; i < count; i++) out[i] = mulmod(a[i], s, &m);
}

/* ---------------------------------------------------------------- context */
#define MAXK 16
#define MAXS 18
static const uint64_t AUX_MODS[] = { /* SU/globals.cpp:330-333 (first entries of aux_small_mods) */
    0x1fffffffffb40001ULL, 0x1fffffffff500001ULL, 0x1fffffffff380001ULL, 0x1fffffffff000001ULL,
    0x1ffffffffef00001ULL, 0x1ffffffffee80001ULL, 0x1ffffffffeb40001ULL, 0x1ffffffffe780001ULL,
    0x1ffffffffe600001ULL, 0x1ffffffffe4c0001ULL, 0x1ffffffffdf40001ULL, 0x1ffffffffdac0001ULL,
    0x1ffffffffda40001ULL, 0x1ffffffffc680001ULL, 0x1ffffffffc000001ULL, 0x1ffffffffb880001ULL,
    0x1ffffffffb7c0001ULL };
#define M_SK 0x1fffffffffe00001ULL   /* SU/globals.cpp:324 */
#define M_TILDE (1ULL << 32)         /* SU/globals.cpp:327 */

struct orc_ctx {
    int n, logn, K, L, S; uint64_t t;
    ntt_tab qt[MAXK], bt[MAXS];
    smod mtilde, msk;
    /* Evaluator ctor, S/evaluator.cpp:66-105 */
    uint64_t delta[MAXK], rho[MAXK], half, lift_inc[MAXK];
    /* BaseConverter ctor, SU/baseconverter.cpp:95-349 */
    uint64_t inv_qhat[MAXK];              /* inv_coeff_base_products_mod_coeff_array_ */
    uint64_t mt_inv_qhat[MAXK];           /* mtilde_inv_coeff_base_products_mod_coeff_array_ */
    uint64_t qhat_mod_bsk[MAXS][MAXK];    /* coeff_base_products_mod_aux_bsk_array_ */
    uint64_t qhat_mod_mt[MAXK];           /* coeff_base_products_mod_mtilde_array_ */
    uint64_t inv_q_mod_mt;                /* inv_coeff_products_mod_mtilde_ */
    uint64_t q_mod_bsk[MAXS];             /* coeff_products_all_mod_bsk_array_ */
    uint64_t inv_mt_mod_bsk[MAXS];        /* inv_mtilde_mod_bsk_array_ */
    uint64_t inv_q_mod_bsk[MAXS];         /* inv_coeff_products_all_mod_aux_bsk_array_ */
    uint64_t inv_Mhat[MAXS];              /* inv_aux_base_products_mod_aux_array_ */
    uint64_t Mhat_mod_q[MAXK][MAXS];      /* aux_base_products_mod_coeff_array_ */
    uint64_t Mhat_mod_msk[MAXS];          /* aux_base_products_mod_msk_array_ */
    uint64_t inv_M_mod_msk;               /* inv_aux_products_mod_msk_ */
    uint64_t M_mod_q[MAXK];               /* aux_products_all_mod_coeff_array_ */
};

static uint64_t prod_mod_except(const uint64_t *v, int cnt, int skip, const smod *m) {
    uint64_t r = 1 % m->q;
    for (int i = 0; i < cnt; i++) if (i != skip) r = mulmod(r, v[i] % m->q, m);
    return r;
}

orc_ctx *orc_create(int n, int K, const uint64_t *primes, uint64_t t) {
    if (K < 1 || K > MAXK - 1) return NULL;
    orc_ctx *c = calloc(1, sizeof(orc_ctx));
    c->n = n; c->K = K; c->t = t; c->logn = 0;
    while ((1 << c->logn) < n) c->logn++;
    if ((1 << c->logn) != n) { free(c); return NULL; }
    for (int i = 0; i < K; i++) if (!ntt_tab_make(&c->qt[i], c->logn, primes[i])) { free(c); return NULL; }
    /* aux base size, SU/baseconverter.cpp:47-58 */
    int total_bits = 0; for (int i = 0; i < K; i++) total_bits += c->qt[i].m.bits;
    c->L = K; if (32 + bit_count(t) + total_bits >= 61 * K + 61) c->L++;
    c->S = c->L + 1;
    uint64_t aux[MAXS], bsk[MAXS];
    for (int i = 0; i < c->L; i++) aux[i] = bsk[i] = AUX_MODS[i];
    bsk[c->L] = M_SK;
    for (int i = 0; i < c->S; i++) if (!ntt_tab_make(&c->bt[i], c->logn, bsk[i])) { free(c); return NULL; }
    c->mtilde = smod_make(M_TILDE); c->msk = smod_make(M_SK);

    /* Evaluator constants: Delta = floor(Q/t), rho = Q - t*Delta = Q mod t, each reduced mod q_j
     * (S/evaluator.cpp:66-105).  Delta mod q_j = -(rho) * t^-1 mod q_j because t*Delta = Q - rho. */
    c->half = (t + 1) >> 1; /* S/evaluator.cpp:73 */
    uint64_t rho;
    { smod tm = smod_make(t); rho = prod_mod_except(primes, K, -1, &tm); }
    for (int j = 0; j < K; j++) {
        const smod *m = &c->qt[j].m;
        c->rho[j] = rho % m->q;
        c->delta[j] = mulmod(negmod(rho % m->q, m), invmod(t % m->q, m->q), m);
        c->lift_inc[j] = m->q - t; /* S/evaluator.cpp:82-86 (fast plain lift) */
    }
    /* BaseConverter constants */
    for (int i = 0; i < K; i++) {
        const smod *m = &c->qt[i].m;
        c->inv_qhat[i] = invmod(prod_mod_except(primes, K, i, m), m->q);
        c->mt_inv_qhat[i] = mulmod(c->inv_qhat[i], M_TILDE, m);
        c->qhat_mod_mt[i] = prod_mod_except(primes, K, i, &c->mtilde);
        c->M_mod_q[i] = prod_mod_except(aux, c->L, -1, m);
        for (int j = 0; j < c->L; j++) c->Mhat_mod_q[i][j] = prod_mod_except(aux, c->L, j, m);
    }
    c->inv_q_mod_mt = invmod(prod_mod_except(primes, K, -1, &c->mtilde), M_TILDE);
    for (int k = 0; k < c->S; k++) {
        const smod *m = &c->bt[k].m;
        for (int i = 0; i < K; i++) c->qhat_mod_bsk[k][i] = prod_mod_except(primes, K, i, m);
        c->q_mod_bsk[k] = prod_mod_except(primes, K, -1, m);
        c->inv_q_mod_bsk[k] = invmod(c->q_mod_bsk[k], m->q);
        c->inv_mt_mod_bsk[k] = invmod(M_TILDE % m->q, m->q);
    }
    for (int i = 0; i < c->L; i++) {
        const smod *m = &c->bt[i].m;
        c->inv_Mhat[i] = invmod(prod_mod_except(aux, c->L, i, m), m->q);
        c->Mhat_mod_msk[i] = prod_mod_except(aux, c->L, i, &c->msk);
    }
    c->inv_M_mod_msk = invmod(prod_mod_except(aux, c->L, -1, &c->msk), M_SK);
    return c;
}
void orc_destroy(orc_ctx *c) {
    if (!c) return;
    for (int i = 0; i < c->K; i++) ntt_tab_free(&c->qt[i]);
    for (int i = 0; i < c->S; i++) ntt_tab_free(&c->bt[i]);
    free(c);
}
int orc_n(const orc_ctx *c) { return c->n; }
int orc_K(const orc_ctx *c) { return c->K; }
int orc_bsk_count(const orc_ctx *c) { return c->S; }
const uint64_t *orc_ntt_table(const orc_ctx *c, int base, int idx, int which) {
    const ntt_tab *t = base ? &c->bt[idx] : &c->qt[idx];
    return which == 0 ? t->rp : which == 1 ? t->srp : which == 2 ? t->irp2 : t->sirp2;
}
uint64_t orc_modulus(const orc_ctx *c, int base, int idx) { return base ? c->bt[idx].m.q : c->qt[idx].m.q; }
uint64_t orc_minimal_root(const orc_ctx *c, int base, int idx) { return base ? c->bt[idx].root : c->qt[idx].root; }

/* ---------------------------------------------------------------- FractionalEncoder (balanced, base 3) */
/* BalancedEncoder::encode(int64_t), S/encoder.cpp:408-481, base 3 */
static int encode_int_b3(uint64_t t, int64_t value, uint64_t *dst, int cap) {
    const uint64_t base = 3;
    int coeff_count;
    if (value < 0) {
        uint64_t pos = (uint64_t)(-value);
        coeff_count = (int)(ceil((double)bit_count((uint64_t)value) / log2((double)base)) + 1);
        int idx = 0;
        while (pos && idx < cap) {
            uint64_t rem = pos % base;
            if (0 < rem && rem <= (base - 1) / 2) dst[idx] = t - rem;
            else if (rem > (base - 1) / 2) dst[idx] = base - rem;
            pos = (pos + ((base - 1) / 2)) / base;
            idx++;
        }
    } else {
        uint64_t v = (uint64_t)value;
        coeff_count = (int)(ceil((double)bit_count(v) / log2((double)base)) + 1);
        int idx = 0;
        while (v && idx < cap) {
            uint64_t rem = v % base;
            if (0 < rem && rem <= (base - 1) / 2) dst[idx] = rem;
            else if (rem > (base - 1) / 2) dst[idx] = t - base + rem;
            v = (v + base / 2) / base;
            idx++;
        }
    }
    return coeff_count;
}
/* BalancedFractionalEncoder::encode_odd, S/encoder.cpp:1013-1076, with integer_coeff_count 64,
 * fraction_coeff_count 32 as in C/globals.cpp:52 */
int orc_encode_fractional(const orc_ctx *c, double value, uint64_t *out) {
    const int frac = 32; int coeff_count = c->n + 1;
    memset(out, 0, (size_t)coeff_count * 8);
    int64_t vi = (int64_t)round(value);
    int int_cc = encode_int_b3(c->t, vi, out, coeff_count);
    value -= (double)vi;
    if (value == 0) return int_cc;
    /* fractional digit i (0-based, most significant first) lands at coefficient n-1-i */
    for (int i = 0; i < frac; i++) {
        value *= 3.0;
        int sign = (value >= 0 ? 1 : -1);
        vi = (int64_t)(sign * ceil(fabs(value) - 0.5));
        value -= (double)vi;
        int neg = 0; if (vi < 0) { neg = 1; vi = -vi; }
        uint64_t coef = (uint64_t)vi;
        if (!neg && vi != 0) coef = c->t - coef;
        out[c->n - 1 - i] = coef;
    }
    return coeff_count;
}

/* ---------------------------------------------------------------- evaluator ops */
#define STRIDE(c) ((size_t)(c)->n + 1)

/* S/evaluator.cpp:1495-1539 */
void orc_ct_transform(const orc_ctx *c, uint64_t *cts, int count, int size, int inverse) {
    for (size_t p = 0; p < (size_t)count * size; p++)
        for (int j = 0; j < c->K; j++) {
            uint64_t *a = cts + (p * c->K + j) * STRIDE(c);
            if (inverse) intt_full(a, &c->qt[j]); else ntt_full(a, &c->qt[j]);
        }
}
/* lift of S/evaluator.cpp:1465-1486 (fast plain lift) into K limb-polys of stride n+1 */
static void plain_lift(const orc_ctx *c, const uint64_t *plain, int coeff_count, uint64_t *out) {
    memset(out, 0, (size_t)c->K * STRIDE(c) * 8);
    for (int j = 0; j < c->K; j++)
        for (int i = 0; i < coeff_count && i < c->n + 1; i++)
            out[j * STRIDE(c) + i] = plain[i] >= c->half ? plain[i] + c->lift_inc[j] : plain[i];
}
/* S/evaluator.cpp:1418-1493 */
void orc_plain_to_ntt(const orc_ctx *c, const uint64_t *plain, int coeff_count, uint64_t *out) {
    plain_lift(c, plain, coeff_count, out);
    for (int j = 0; j < c->K; j++) ntt_full(out + j * STRIDE(c), &c->qt[j]);
}
/* S/evaluator.cpp:1541-1585 */
void orc_multiply_plain_ntt(const orc_ctx *c, uint64_t *cts, int count, int size, const uint64_t *pn) {
    for (size_t p = 0; p < (size_t)count * size; p++)
        for (int j = 0; j < c->K; j++) {
            uint64_t *a = cts + (p * c->K + j) * STRIDE(c);
            const uint64_t *w = pn + j * STRIDE(c);
            for (int i = 0; i < c->n; i++) a[i] = mulmod(a[i], w[i], &c->qt[j].m);
        }
}
/* multiply_plain S/evaluator.cpp:1243-1416; add_plain :1145-1192; sub_plain :1194-1241 */
void orc_plain_op(const orc_ctx *c, uint64_t *cts, int count, int size, const uint64_t *plain, int cc, int op) {
    if (op == 0) {
        if (cc == 1) { /* constant branch :1278-1341 */
            for (size_t p = 0; p < (size_t)count * size; p++)
                for (int j = 0; j < c->K; j++) {
                    uint64_t s = plain[0] >= c->half ? plain[0] + c->lift_inc[j] : plain[0];
                    uint64_t *a = cts + (p * c->K + j) * STRIDE(c);
                    for (int i = 0; i < c->n + 1; i++) a[i] = mulmod(a[i], s, &c->qt[j].m);
                }
            return;
        }
        uint64_t *pn = malloc((size_t)c->K * STRIDE(c) * 8);
        orc_plain_to_ntt(c, plain, cc, pn); /* :1366-1396 */
        for (size_t p = 0; p < (size_t)count * size; p++)
            for (int j = 0; j < c->K; j++) { /* :1398-1415 */
                uint64_t *a = cts + (p * c->K + j) * STRIDE(c);
                ntt_lazy(a, &c->qt[j]);
                for (int i = 0; i < c->n + 1; i++) a[i] = mulmod(a[i], pn[j * STRIDE(c) + i], &c->qt[j].m);
                intt_full(a, &c->qt[j]);
            }
        free(pn);
        return;
    }
    for (int k = 0; k < count; k++) {
        uint64_t *c0 = cts + (size_t)k * size * c->K * STRIDE(c);
        for (int i = 0; i < cc; i++)
            for (int j = 0; j < c->K; j++) {
                const smod *m = &c->qt[j].m; uint64_t sc;
                if (plain[i] >= c->half) sc = barrett_u128((u128)c->delta[j] * plain[i] + c->rho[j], m);
                else sc = mulmod(c->delta[j], plain[i], m);
                uint64_t *x = c0 + j * STRIDE(c) + i;
                *x = (op == 1) ? addmod(*x, sc, m) : submod(*x, sc, m);
            }
    }
}
/* S/evaluator.cpp:254-308 */
void orc_add_many(const orc_ctx *c, const uint64_t *cts, int count, int size, uint64_t *out) {
    size_t w = (size_t)size * c->K * STRIDE(c);
    memcpy(out, cts, w * 8);
    for (int k = 1; k < count; k++)
        for (int p = 0; p < size; p++)
            for (int j = 0; j < c->K; j++) {
                size_t off = ((size_t)p * c->K + j) * STRIDE(c);
                for (int i = 0; i < c->n + 1; i++) out[off + i] = addmod(out[off + i], cts[k * w + off + i], &c->qt[j].m);
            }
}

/* ---- BEHZ pieces, per polynomial; buffers are [limbs][n+1] ---- */
/* SU/baseconverter.cpp:663-742 */
static void fastbconv_mtilde(const orc_ctx *c, const uint64_t *in, uint64_t *out /* [S+1][n+1] */) {
    size_t st = STRIDE(c);
    for (size_t k = 0; k < st; k++) {
        uint64_t y[MAXK];
        for (int i = 0; i < c->K; i++) y[i] = mulmod(in[i * st + k], c->mt_inv_qhat[i], &c->qt[i].m);
        for (int j = 0; j < c->S; j++) {
            u128 acc = 0; for (int i = 0; i < c->K; i++) acc += (u128)y[i] * c->qhat_mod_bsk[j][i];
            out[j * st + k] = barrett_u128(acc, &c->bt[j].m);
        }
        u128 acc = 0; for (int i = 0; i < c->K; i++) acc += (u128)y[i] * c->qhat_mod_mt[i];
        out[(size_t)c->S * st + k] = barrett_u128(acc, &c->mtilde);
    }
}
/* SU/baseconverter.cpp:581-622 */
static void mont_rq(const orc_ctx *c, const uint64_t *in /* [S+1][n+1] */, uint64_t *out /* [S][n+1] */) {
    size_t st = STRIDE(c);
    for (int k = 0; k < c->S; k++)
        for (size_t i = 0; i < st; i++) {
            uint64_t r = mulmod(in[(size_t)c->S * st + i], c->inv_q_mod_mt, &c->mtilde);
            r = negmod(r, &c->mtilde);
            u128 tmp = (u128)c->q_mod_bsk[k] * r + in[k * st + i];
            uint64_t v = barrett_u128(tmp, &c->bt[k].m);
            out[k * st + i] = mulmod(v, c->inv_mt_mod_bsk[k], &c->bt[k].m);
        }
}
/* SU/baseconverter.cpp:388-446 then :624-661; in = [K+S][n+1] */
static void fast_floor(const orc_ctx *c, const uint64_t *in, uint64_t *out /* [S][n+1] */) {
    size_t st = STRIDE(c);
    for (size_t k = 0; k < st; k++) {
        uint64_t u[MAXK];
        for (int i = 0; i < c->K; i++) u[i] = mulmod(in[i * st + k], c->inv_qhat[i], &c->qt[i].m);
        for (int j = 0; j < c->S; j++) {
            u128 acc = 0; for (int i = 0; i < c->K; i++) acc += (u128)u[i] * c->qhat_mod_bsk[j][i];
            uint64_t v = barrett_u128(acc, &c->bt[j].m);
            uint64_t x = in[((size_t)c->K + j) * st + k];
            out[j * st + k] = mulmod(x + c->bt[j].m.q - v, c->inv_q_mod_bsk[j], &c->bt[j].m);
        }
    }
}
/* SU/baseconverter.cpp:448-579; in = [S][n+1] (aux limbs then m_sk), out = [K][n+1] */
static void fastbconv_sk(const orc_ctx *c, const uint64_t *in, uint64_t *out) {
    size_t st = STRIDE(c);
    uint64_t msk = c->msk.q, msk_half = msk >> 1;
    for (size_t k = 0; k < st; k++) {
        uint64_t g[MAXS];
        for (int i = 0; i < c->L; i++) g[i] = mulmod(in[i * st + k], c->inv_Mhat[i], &c->bt[i].m);
        u128 acc = 0; for (int i = 0; i < c->L; i++) acc += (u128)g[i] * c->Mhat_mod_msk[i];
        uint64_t s = barrett_u128(acc, &c->msk);
        uint64_t alpha = mulmod(s + (msk - in[(size_t)c->L * st + k]), c->inv_M_mod_msk, &c->msk);
        for (int j = 0; j < c->K; j++) {
            const smod *m = &c->qt[j].m;
            acc = 0; for (int i = 0; i < c->L; i++) acc += (u128)g[i] * c->Mhat_mod_q[j][i];
            uint64_t e = barrett_u128(acc, m);
            if (alpha > msk_half) e = barrett_u128((u128)c->M_mod_q[j] * (msk - alpha) + e, m);
            else e = barrett_u128((u128)(m->q - c->M_mod_q[j]) * alpha + e, m);
            out[j * st + k] = e;
        }
    }
}

/* S/evaluator.cpp:702-884 (size-2 input) */
void orc_square(const orc_ctx *c, const uint64_t *in, int count, uint64_t *out3) {
    size_t st = STRIDE(c); int K = c->K, S = c->S;
    uint64_t *bsk_mt = malloc((size_t)(S + 1) * st * 8);
    uint64_t *xq = malloc((size_t)2 * K * st * 8), *xb = malloc((size_t)2 * S * st * 8);
    uint64_t *dq = malloc((size_t)3 * K * st * 8), *db = malloc((size_t)3 * S * st * 8);
    uint64_t *tog = malloc((size_t)(K + S) * st * 8), *flo = malloc((size_t)S * st * 8);
    for (int ci = 0; ci < count; ci++) {
        const uint64_t *ct = in + (size_t)ci * 2 * K * st;
        uint64_t *res = out3 + (size_t)ci * 3 * K * st;
        for (int p = 0; p < 2; p++) { /* :742-751 */
            fastbconv_mtilde(c, ct + (size_t)p * K * st, bsk_mt);
            mont_rq(c, bsk_mt, xb + (size_t)p * S * st);
        }
        memcpy(xq, ct, (size_t)2 * K * st * 8);
        for (int p = 0; p < 2; p++) { /* :769-779 */
            for (int j = 0; j < K; j++) ntt_lazy(xq + ((size_t)p * K + j) * st, &c->qt[j]);
            for (int j = 0; j < S; j++) ntt_lazy(xb + ((size_t)p * S + j) * st, &c->bt[j]);
        }
        for (int j = 0; j < K; j++) { /* :783-834 in base q */
            const smod *m = &c->qt[j].m; const uint64_t *a = xq + (size_t)j * st, *b = xq + ((size_t)K + j) * st;
            for (size_t i = 0; i < st; i++) {
                dq[(size_t)j * st + i] = mulmod(a[i], a[i], m);
                dq[((size_t)2 * K + j) * st + i] = mulmod(b[i], b[i], m);
                uint64_t x = mulmod(a[i], b[i], m); dq[((size_t)K + j) * st + i] = addmod(x, x, m);
            }
        }
        for (int j = 0; j < S; j++) { /* in base Bsk */
            const smod *m = &c->bt[j].m; const uint64_t *a = xb + (size_t)j * st, *b = xb + ((size_t)S + j) * st;
            for (size_t i = 0; i < st; i++) {
                db[(size_t)j * st + i] = mulmod(a[i], a[i], m);
                db[((size_t)2 * S + j) * st + i] = mulmod(b[i], b[i], m);
                uint64_t x = mulmod(a[i], b[i], m); db[((size_t)S + j) * st + i] = addmod(x, x, m);
            }
        }
        for (int p = 0; p < 3; p++) { /* :837-883 */
            for (int j = 0; j < K; j++) {
                uint64_t *a = dq + ((size_t)p * K + j) * st; intt_lazy(a, &c->qt[j]);
                for (size_t i = 0; i < st; i++) tog[(size_t)j * st + i] = mulmod(a[i], c->t, &c->qt[j].m);
            }
            for (int j = 0; j < S; j++) {
                uint64_t *a = db + ((size_t)p * S + j) * st; intt_lazy(a, &c->bt[j]);
                for (size_t i = 0; i < st; i++) tog[((size_t)K + j) * st + i] = mulmod(a[i], c->t, &c->bt[j].m);
            }
            fast_floor(c, tog, flo);
            fastbconv_sk(c, flo, res + (size_t)p * K * st);
        }
    }
    free(bsk_mt); free(xq); free(xb); free(dq); free(db); free(tog); free(flo);
}

/* S/evaluator.cpp:886-1069 (one relinearize step 3 -> 2) */
void orc_relinearize(const orc_ctx *c, const uint64_t *in3, int count, const uint64_t *evk, const int *sizes,
                     int dbc, uint64_t *out2) {
    size_t st = STRIDE(c); int K = c->K;
    u128 *acc0 = malloc((size_t)K * st * sizeof(u128)), *acc1 = malloc((size_t)K * st * sizeof(u128));
    uint64_t *d = malloc(st * 8), *dig = malloc(st * 8), *tmp = malloc(st * 8);
    for (int ci = 0; ci < count; ci++) {
        const uint64_t *ct = in3 + (size_t)ci * 3 * K * st;
        uint64_t *res = out2 + (size_t)ci * 2 * K * st;
        memcpy(res, ct, (size_t)2 * K * st * 8);
        memset(acc0, 0, (size_t)K * st * sizeof(u128)); memset(acc1, 0, (size_t)K * st * sizeof(u128));
        const uint64_t *key = evk;
        for (int i = 0; i < K; i++) {
            const uint64_t *c2 = ct + ((size_t)2 * K + i) * st;
            for (size_t m = 0; m < st; m++) d[m] = mulmod(c2[m], c->inv_qhat[i], &c->qt[i].m); /* :984-985 */
            int shift = 0;
            for (int k = 0; k < sizes[i]; k += 2) {
                const uint64_t *k0 = key + (size_t)k * K * st, *k1 = key + (size_t)(k + 1) * K * st;
                for (size_t m = 0; m < st; m++) dig[m] = (d[m] >> shift) & ((1ULL << dbc) - 1); /* :997-1001 */
                for (int j = 0; j < K; j++) {
                    memcpy(tmp, dig, st * 8);
                    ntt_lazy(tmp, &c->qt[j]); /* :1011 */
                    for (size_t m = 0; m < st; m++) { /* :1015-1030 */
                        acc0[(size_t)j * st + m] += (u128)tmp[m] * k0[(size_t)j * st + m];
                        acc1[(size_t)j * st + m] += (u128)tmp[m] * k1[(size_t)j * st + m];
                    }
                }
                shift += dbc;
            }
            key += (size_t)sizes[i] * K * st;
        }
        for (int p = 0; p < 2; p++) { /* :1041-1068 */
            u128 *acc = p ? acc1 : acc0;
            for (int j = 0; j < K; j++) {
                for (size_t m = 0; m < st; m++) tmp[m] = barrett_u128(acc[(size_t)j * st + m], &c->qt[j].m);
                intt_full(tmp, &c->qt[j]);
                uint64_t *dst = res + ((size_t)p * K + j) * st;
                for (size_t m = 0; m < st; m++) dst[m] = addmod(dst[m], tmp[m], &c->qt[j].m);
            }
        }
    }
    free(acc0); free(acc1); free(d); free(dig); free(tmp);
}

/* ---------------------------------------------------------------- CrCNN layers */
/* Layer::computeBoundaries, C/layer.cpp:12-26 */
static void boundaries(int xd, int yd, int xs, int ys, int xf, int yf, int *xl, int *yl) {
    *xl = (xf > xs) ? xd - xf + 1 : xd - xs + 1;
    *yl = (yf > ys) ? yd - yf + 1 : yd - ys + 1;
}

/* One output neuron exactly as the reference's inner loops do it
 * (C/convolutionalLayer.cpp:72-88, C/fullyConnectedLayer.cpp:123-140): per term copy the NTT-form
 * input, multiply_plain_ntt, transform_from_ntt; add_plain(bias) on the first term; add_many. */
static void weighted_sum(const orc_ctx *c, const uint64_t *in_ntt, const int *in_idx, const uint64_t *w_ntt,
                         const int *w_idx, int terms, const uint64_t *bias, uint64_t *out, uint64_t *scratch) {
    size_t ctw = (size_t)2 * c->K * STRIDE(c), pw = (size_t)c->K * STRIDE(c);
    for (int r = 0; r < terms; r++) {
        uint64_t *tmp = scratch + (size_t)r * ctw;
        memcpy(tmp, in_ntt + (size_t)in_idx[r] * ctw, ctw * 8);
        orc_multiply_plain_ntt(c, tmp, 1, 2, w_ntt + (size_t)w_idx[r] * pw);
        orc_ct_transform(c, tmp, 1, 2, 1);
    }
    orc_plain_op(c, scratch, 1, 2, bias, c->n + 1, 1);
    orc_add_many(c, scratch, terms, 2, out);
}

/* C/convolutionalLayer.cpp:159-197 + :56-93 */
void orc_conv_forward(const orc_ctx *c, const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf,
                      int nf, const uint64_t *w, const uint64_t *b, uint64_t *out) {
    size_t ctw = (size_t)2 * c->K * STRIDE(c), pw = (size_t)c->K * STRIDE(c), st = STRIDE(c);
    int xo = (xd - xf) / xs + 1, yo = (yd - yf) / ys + 1, R = zd * xf * yf, nin = zd * xd * yd;
    uint64_t *xin = malloc((size_t)nin * ctw * 8);
    memcpy(xin, in, (size_t)nin * ctw * 8);
    orc_ct_transform(c, xin, nin, 2, 0); /* transform_input_to_ntt :95-148 */
    uint64_t *wn = malloc((size_t)nf * R * pw * 8);
    for (int i = 0; i < nf * R; i++) orc_plain_to_ntt(c, w + (size_t)i * st, c->n + 1, wn + (size_t)i * pw); /* :151-156 */
    uint64_t *scratch = malloc((size_t)R * ctw * 8);
    int *ii = malloc(sizeof(int) * R), *wi = malloc(sizeof(int) * R);
    int xl, yl; boundaries(xd, yd, xs, ys, xf, yf, &xl, &yl);
    memset(out, 0, (size_t)nf * xo * yo * ctw * 8);
    for (int k = 0; k < nf; k++)
        for (int i = 0; i < xl; i += xs)
            for (int j = 0; j < yl; j += ys) {
                int p = 0;
                for (int z = 0; z < zd; z++) for (int kx = 0; kx < xf; kx++) for (int ky = 0; ky < yf; ky++) {
                    ii[p] = (z * xd + i + kx) * yd + j + ky; wi[p] = k * R + p; p++;
                }
                weighted_sum(c, xin, ii, wn, wi, R, b + (size_t)k * st,
                             out + (((size_t)k * xo + i / xs) * yo + j / ys) * ctw, scratch);
            }
    free(xin); free(wn); free(scratch); free(ii); free(wi);
}
/* C/fullyConnectedLayer.cpp:113-168; input already flattened row-major (reshapeInput :38-56) */
void orc_fc_forward(const orc_ctx *c, const uint64_t *in, int in_dim, int out_dim, const uint64_t *w,
                    const uint64_t *b, uint64_t *out) {
    size_t ctw = (size_t)2 * c->K * STRIDE(c), pw = (size_t)c->K * STRIDE(c), st = STRIDE(c);
    uint64_t *xin = malloc((size_t)in_dim * ctw * 8);
    memcpy(xin, in, (size_t)in_dim * ctw * 8);
    orc_ct_transform(c, xin, in_dim, 2, 0);
    uint64_t *wn = malloc((size_t)in_dim * pw * 8), *scratch = malloc((size_t)in_dim * ctw * 8);
    int *ii = malloc(sizeof(int) * in_dim);
    for (int j = 0; j < in_dim; j++) ii[j] = j;
    for (int i = 0; i < out_dim; i++) {
        for (int j = 0; j < in_dim; j++) orc_plain_to_ntt(c, w + ((size_t)i * in_dim + j) * st, c->n + 1, wn + (size_t)j * pw);
        weighted_sum(c, xin, ii, wn, ii, in_dim, b + (size_t)i * st, out + (size_t)i * ctw, scratch);
    }
    free(xin); free(wn); free(scratch); free(ii);
}
/* C/poolingLayer.cpp:22-44, C/avgPoolingLayer.cpp:16-45 */
void orc_pool_forward(const orc_ctx *c, const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf,
                      const uint64_t *div_plain, int div_cc, uint64_t *out) {
    size_t ctw = (size_t)2 * c->K * STRIDE(c);
    int xo = (xd - xf) / xs + 1, yo = (yd - yf) / ys + 1, xl, yl;
    boundaries(xd, yd, xs, ys, xf, yf, &xl, &yl);
    uint64_t *pix = malloc((size_t)xf * yf * ctw * 8);
    memset(out, 0, (size_t)zd * xo * yo * ctw * 8);
    for (int z = 0; z < zd; z++)
        for (int i = 0; i < xl; i += xs)
            for (int j = 0; j < yl; j += ys) {
                int p = 0;
                for (int kx = 0; kx < xf; kx++) for (int ky = 0; ky < yf; ky++)
                    memcpy(pix + (size_t)(p++) * ctw, in + (((size_t)z * xd + i + kx) * yd + j + ky) * ctw, ctw * 8);
                uint64_t *o = out + (((size_t)z * xo + i / xs) * yo + j / ys) * ctw;
                orc_add_many(c, pix, xf * yf, 2, o);
                if (div_plain) orc_plain_op(c, o, 1, 2, div_plain, div_cc, 0);
            }
    free(pix);
}
/* C/batchNormLayer.cpp:29-40; mean/invstd are [zd][n+1] plaintext words */
void orc_bn_forward(const orc_ctx *c, const uint64_t *in, int zd, int xd, int yd, const uint64_t *mean,
                    const uint64_t *invstd, uint64_t *out) {
    size_t ctw = (size_t)2 * c->K * STRIDE(c), st = STRIDE(c);
    memcpy(out, in, (size_t)zd * xd * yd * ctw * 8);
    for (int z = 0; z < zd; z++) {
        uint64_t *o = out + (size_t)z * xd * yd * ctw;
        orc_plain_op(c, o, xd * yd, 2, mean + (size_t)z * st, c->n + 1, 2);
        orc_plain_op(c, o, xd * yd, 2, invstd + (size_t)z * st, c->n + 1, 0);
    }
}
/* C/squareLayer.cpp:22-71 */
void orc_square_forward(const orc_ctx *c, const uint64_t *in, int count, const uint64_t *evk, const int *sizes,
                        int dbc, uint64_t *out) {
    size_t w3 = (size_t)3 * c->K * STRIDE(c);
    uint64_t *tmp = malloc(w3 * 8);
    for (int i = 0; i < count; i++) {
        orc_square(c, in + (size_t)i * 2 * c->K * STRIDE(c), 1, tmp);
        orc_relinearize(c, tmp, 1, evk, sizes, dbc, out + (size_t)i * 2 * c->K * STRIDE(c));
    }
    free(tmp);
}

/* ================================================================ client-side steps of the re-encryption
 * Network::forward decrypts and re-encrypts the whole tensor before layer 6 (C/network.cpp:30-33: decryptImage -> encryptImage,
 * C/globals.cpp:127-142, 207-226).  These restate Decryptor::decrypt, FractionalEncoder::decode and Encryptor::encrypt for that step. */
#define GAMMA 0x1fffffffffc80001ULL   /* SU/globals.cpp:330 (internal_mods::gamma) */

/* Decryptor::decrypt for size-2 ciphertexts, S/decryptor.cpp:107-234.  sk_ntt: the secret key in NTT form [K][n+1] as the
 * Decryptor keeps it (secret_key_array_, power 1); plain_out: n+1 words per ciphertext (values < t, pad word 0). */
void orc_decrypt(const orc_ctx *c, const uint64_t *cts, int count, const uint64_t *sk_ntt, uint64_t *plain_out) {
    const int n = c->n, K = c->K;
    const size_t st = STRIDE(c), ctw = (size_t)2 * K * st;
    smod tm = smod_make(c->t), gm = smod_make(GAMMA);
    uint64_t primes[MAXK];
    for (int i = 0; i < K; i++) primes[i] = c->qt[i].m.q;
    /* BaseConverter constants, SU/baseconverter.cpp:315-349 */
    uint64_t qhat_t[MAXK], qhat_g[MAXK], tg_mod_q[MAXK];
    for (int i = 0; i < K; i++) {
        qhat_t[i] = prod_mod_except(primes, K, i, &tm);
        qhat_g[i] = prod_mod_except(primes, K, i, &gm);
        tg_mod_q[i] = mulmod(c->t % c->qt[i].m.q, GAMMA % c->qt[i].m.q, &c->qt[i].m);   /* plain_gamma_product_mod_coeff_array_ :345-349 */
    }
    const uint64_t neg_inv_q_t = invmod(negmod(prod_mod_except(primes, K, -1, &tm), &tm), c->t);   /* :325-335 */
    const uint64_t neg_inv_q_g = invmod(negmod(prod_mod_except(primes, K, -1, &gm), &gm), GAMMA);
    const uint64_t inv_gamma_t = invmod(GAMMA % c->t, c->t);                                      /* :337-343 */
    uint64_t *y = malloc((size_t)K * st * 8), *tmp = malloc(st * 8);
    for (int ct = 0; ct < count; ct++) {
        const uint64_t *c0 = cts + (size_t)ct * ctw, *c1 = c0 + (size_t)K * st;
        for (int i = 0; i < K; i++) {
            const smod *m = &c->qt[i].m;
            memcpy(tmp, c1 + (size_t)i * st, st * 8);
            ntt_lazy(tmp, &c->qt[i]);                                       /* :158 */
            for (int k = 0; k < n; k++) tmp[k] = mulmod(tmp[k], sk_ntt[(size_t)i * st + k], m);   /* dyadic_product_coeffmod :160 */
            intt_full(tmp, &c->qt[i]);                                      /* :169 */
            for (int k = 0; k < n; k++)                                     /* lazy "+ c0" then x |gamma t|_qi, :179-187 */
                y[(size_t)i * st + k] = mulmod(tmp[k] + c0[(size_t)i * st + k], tg_mod_q[i], m);
        }
        uint64_t *out = plain_out + (size_t)ct * st;
        for (int k = 0; k < n; k++) {
            u128 st_ = 0, sg = 0;
            for (int i = 0; i < K; i++) {                                   /* fastbconv_plain_gamma, SU/baseconverter.cpp:744-795 */
                uint64_t v = mulmod(y[(size_t)i * st + k], c->inv_qhat[i], &c->qt[i].m);
                st_ += (u128)v * qhat_t[i];
                sg += (u128)v * qhat_g[i];
            }
            uint64_t at = mulmod(barrett_u128(st_, &tm), neg_inv_q_t, &tm);   /* S/decryptor.cpp:196-201 */
            uint64_t ag = mulmod(barrett_u128(sg, &gm), neg_inv_q_g, &gm);
            uint64_t w;
            if (ag > (GAMMA >> 1)) w = addmod(at, (GAMMA - ag) % c->t, &tm);  /* :207-224 */
            else w = submod(at, ag % c->t, &tm);
            out[k] = mulmod(w, inv_gamma_t, &tm);                             /* :233-234 */
        }
        out[n] = 0;
    }
    free(y); free(tmp);
}

/* BalancedFractionalEncoder::decode with (64, 32, base 3), S/encoder.cpp (BalancedFractionalEncoder::decode,
 * BalancedEncoder::decode_int64): plain has n+1 words. */
double orc_decode_fractional(const orc_ctx *c, const uint64_t *plain) {
    const uint64_t t = c->t, neg = (t + 1) >> 1;   /* coeff_neg_threshold_ = (plain_modulus + 1) >> 1, S/encoder.cpp (BalancedEncoder ctor) */
    int64_t ip = 0;
    int top = 63;
    while (top >= 0 && plain[top] == 0) top--;
    for (int i = top; i >= 0; i--) {
        int64_t v = plain[i] >= neg ? -(int64_t)(t - plain[i]) : (int64_t)plain[i];
        ip = (int64_t)((uint64_t)ip * 3u) + v;
    }
    double frac = 0;
    const uint64_t *f = plain + c->n - 32;        /* plain_copy + coeff_count - 1 - fraction_coeff_count_, coeff_count = n+1 */
    for (int i = 0; i < 32; i++) {
        int64_t v = f[i] >= neg ? -(int64_t)(t - f[i]) : (int64_t)f[i];
        frac += (double)v;
        frac /= 3.0;
    }
    return (double)ip - frac;
}

/* Encryptor::encrypt, S/encryptor.cpp:95-166, with the three sampled polynomials given explicitly (u in {-1,0,1}; e0, e1 the
 * rounded clipped-normal noise, as signed small integers, n entries each): c0 = pk0*u + e0 + Delta*m (preencrypt :168-200),
 * c1 = pk1*u + e1.  pk: the public key as SEAL stores it, [2][K][n+1] in NTT form. */
void orc_encrypt(const orc_ctx *c, const uint64_t *plain, int coeff_count, const uint64_t *pk, const int8_t *u, const int8_t *e0,
                 const int8_t *e1, uint64_t *out) {
    const int n = c->n, K = c->K;
    const size_t st = STRIDE(c);
    uint64_t *un = malloc(st * 8), *tmp = malloc(st * 8);
    memset(out, 0, (size_t)2 * K * st * 8);
    for (int i = 0; i < K; i++) {
        const smod *m = &c->qt[i].m;
        for (int k = 0; k < n; k++) un[k] = u[k] > 0 ? 1 : (u[k] < 0 ? m->q - 1 : 0);
        un[n] = 0;
        ntt_full(un, &c->qt[i]);    /* ntt_double_multiply_poly_nttpoly, SU/polyarithsmallmod.h: NTT(u) once, two dyadic products, two inverse NTTs */
        for (int p = 0; p < 2; p++) {
            const int8_t *e = p ? e1 : e0;
            for (int k = 0; k < n; k++) tmp[k] = mulmod(un[k], pk[((size_t)p * K + i) * st + k], m);
            tmp[n] = 0;
            intt_full(tmp, &c->qt[i]);
            uint64_t *dst = out + ((size_t)p * K + i) * st;
            for (int k = 0; k < n; k++) {
                uint64_t v = tmp[k];
                if (p == 0 && k < coeff_count) {                         /* preencrypt */
                    uint64_t pm = plain[k];
                    u128 z = (u128)c->delta[i] * pm;
                    if (pm >= c->half) z += c->rho[i];
                    v = addmod(v, barrett_u128(z, m), m);
                }
                uint64_t en = e[k] > 0 ? (uint64_t)e[k] : (e[k] < 0 ? m->q - (uint64_t)(-(int)e[k]) : 0);
                dst[k] = addmod(en, v, m);
            }
        }
    }
    free(un); free(tmp);
}
