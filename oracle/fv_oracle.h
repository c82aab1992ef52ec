/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference algorithm for the
 * CrCNN encrypted-forward hot path.  Never linked into or called from the product path.
 *
 * Plain C restatement of SEAL 2.3.1's full-RNS FV evaluator operations and of the CrCNN layer
 * forwards built on them.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference).  Parity of this oracle is PINNED: tests/test_oracle_*.py
 * check it byte-for-byte against (a) SEAL's own known-answer vectors (SEALTest/util/smallntt.cpp,
 * uintarithsmallmod.cpp, polyarithsmallmod.cpp), (b) outputs of the unmodified reference compiled
 * into oracle/_ref/libcrcnn_ref.so, and (c) the committed fixtures under tests/golden/ which were
 * generated from that reference by tests/golden/make_golden.py.
 *
 * Buffer conventions (SEAL layout, SEAL/seal/ciphertext.h:448-452, :647-660):
 *   ciphertext  = uint64[size][K][n+1]   (trailing pad word of every limb-poly is 0)
 *   plaintext   = uint64[coeff_count], coeff_count <= n+1, values < t
 *   NTT plain   = uint64[K][n+1]
 */
#ifndef FV_ORACLE_H
#define FV_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

/* SEALContext + Evaluator + BaseConverter constant derivation
 * (SEAL/seal/context.cpp:23-165, evaluator.cpp:19-121, util/baseconverter.cpp:20-349). */
orc_ctx *orc_create(int n, int K, const uint64_t *primes, uint64_t t);
void orc_destroy(orc_ctx *c);
int orc_n(const orc_ctx *c);
int orc_K(const orc_ctx *c);
int orc_bsk_count(const orc_ctx *c);
/* which: 0 root_powers, 1 scaled_root_powers, 2 inv_root_powers_div_two, 3 scaled twin;
 * base: 0 = coeff prime `idx`, 1 = Bsk prime `idx`. Returns pointer to n words. */
const uint64_t *orc_ntt_table(const orc_ctx *c, int base, int idx, int which);
uint64_t orc_modulus(const orc_ctx *c, int base, int idx);
uint64_t orc_minimal_root(const orc_ctx *c, int base, int idx);

/* scalar primitives (SEAL/seal/util/uintarithsmallmod.h:137-190) exposed for known-answer tests */
uint64_t orc_barrett_reduce_128(uint64_t lo, uint64_t hi, uint64_t q);
uint64_t orc_mulmod(uint64_t a, uint64_t b, uint64_t q);
int orc_try_minimal_primitive_root(uint64_t degree, uint64_t q, uint64_t *root);
/* stand-alone transform of one limb-poly for an arbitrary NTT prime (KAT for smallntt) */
int orc_ntt_single(uint64_t *poly, int logn, uint64_t q, int inverse);
void orc_dyadic_product(const uint64_t *a, const uint64_t *b, int count, uint64_t q, uint64_t *out);
void orc_multiply_poly_scalar(const uint64_t *a, int count, uint64_t s, uint64_t q, uint64_t *out);

/* FractionalEncoder(t, x^n+1, 64, 32, base 3).encode (SEAL/seal/encoder.cpp:1013-1076, :441-481).
 * out has n+1 words (zero padded). Returns SEAL's coeff_count of the encoding. */
int orc_encode_fractional(const orc_ctx *c, double value, uint64_t *out);

/* Evaluator operations, in place on `count` ciphertexts of `size` polys */
void orc_ct_transform(const orc_ctx *c, uint64_t *cts, int count, int size, int inverse);
void orc_plain_to_ntt(const orc_ctx *c, const uint64_t *plain, int coeff_count, uint64_t *out);
void orc_multiply_plain_ntt(const orc_ctx *c, uint64_t *cts, int count, int size, const uint64_t *plain_ntt);
/* op: 0 multiply_plain, 1 add_plain, 2 sub_plain */
void orc_plain_op(const orc_ctx *c, uint64_t *cts, int count, int size, const uint64_t *plain, int coeff_count, int op);
void orc_add_many(const orc_ctx *c, const uint64_t *cts, int count, int size, uint64_t *out);
void orc_square(const orc_ctx *c, const uint64_t *in, int count, uint64_t *out3);
/* evk: keys_[0][i] back to back, each [sizes[i]][K][n+1]; dbc = decomposition bit count */
void orc_relinearize(const orc_ctx *c, const uint64_t *in3, int count, const uint64_t *evk, const int *sizes,
                     int dbc, uint64_t *out2);

/* CrCNN layers; tensors [z][x][y] of size-2 cts; plaintext parameters as [.. ][n+1] words */
void orc_conv_forward(const orc_ctx *c, const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf,
                      int nf, const uint64_t *w, const uint64_t *b, uint64_t *out);
void orc_fc_forward(const orc_ctx *c, const uint64_t *in, int in_dim, int out_dim, const uint64_t *w,
                    const uint64_t *b, uint64_t *out);
void orc_pool_forward(const orc_ctx *c, const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf,
                      const uint64_t *div_plain /* NULL = sum pool */, int div_coeff_count, uint64_t *out);
void orc_bn_forward(const orc_ctx *c, const uint64_t *in, int zd, int xd, int yd, const uint64_t *mean,
                    const uint64_t *invstd, uint64_t *out);
void orc_square_forward(const orc_ctx *c, const uint64_t *in, int count, const uint64_t *evk, const int *sizes,
                        int dbc, uint64_t *out);

/* Client-side steps of the reference's re-encryption inside Network::forward (CrCNN/src/network.cpp:30-33), SURVEY 8(f) N4:
 * Decryptor::decrypt (SEAL/seal/decryptor.cpp:107-234), FractionalEncoder::decode, Encryptor::encrypt (encryptor.cpp:95-166)
 * with the sampled polynomials u, e0, e1 supplied by the caller (signed small integers, n entries each). */
void orc_decrypt(const orc_ctx *c, const uint64_t *cts, int count, const uint64_t *sk_ntt, uint64_t *plain_out);
double orc_decode_fractional(const orc_ctx *c, const uint64_t *plain);
void orc_encrypt(const orc_ctx *c, const uint64_t *plain, int coeff_count, const uint64_t *pk, const int8_t *u, const int8_t *e0,
                 const int8_t *e1, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
