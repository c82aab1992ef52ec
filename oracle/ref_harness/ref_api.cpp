// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the product path.
//
// extern "C" harness around the UNMODIFIED reference: SEAL 2.3.1 (Evaluator,
// Encryptor, Decryptor, KeyGenerator, FractionalEncoder) and the CrCNN layer
// classes, all compiled from /root/reference by oracle/Makefile.ref into
// oracle/_ref/libcrcnn_ref.so.  It lets the tests (ctypes) and bench.py's
// `--impl reference` / cpu_baseline legs drive the real reference on raw
// SEAL-layout word buffers: ciphertexts are uint64[count][size][K][n+1]
// (reference: SEAL/seal/ciphertext.h:448-452), plaintexts uint64[coeff_count].
//
// Everything here is our own glue; no reference code is copied.  The only
// deviation from the reference's `setParameters` (CrCNN/src/globals.cpp:25-56)
// is a deterministic UniformRandomGeneratorFactory so that keys and
// encryptions are reproducible (EncryptionParameters::set_random_generator,
// SEAL/seal/encryptionparams.h:223).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <atomic>

#include "seal/seal.h"
#include "globals.h"
#include "layer.h"
#include "convolutionalLayer.h"
#include "fullyConnectedLayer.h"
#include "poolingLayer.h"
#include "avgPoolingLayer.h"
#include "batchNormLayer.h"
#include "squareLayer.h"
#include "network.h"

using namespace seal;
using namespace std;

namespace {

// One independently seeded mt19937_64 per create() call: SEAL calls create()
// once per encryption (SEAL/seal/encryptor.cpp:103) and once in key generation.
class SeededGen : public UniformRandomGenerator {
public:
    explicit SeededGen(uint64_t seed) : eng_(seed) {}
    uint32_t generate() override { return static_cast<uint32_t>(eng_()); }
private:
    std::mt19937_64 eng_;
};

class SeededFactory : public UniformRandomGeneratorFactory {
public:
    explicit SeededFactory(uint64_t seed) : seed_(seed), ctr_(0) {}
    UniformRandomGenerator *create() override {
        uint64_t c = ctr_.fetch_add(1);
        return new SeededGen(seed_ * 0x9E3779B97F4A7C15ULL + c * 0xD1B54A32D192ED03ULL + 12345);
    }
private:
    uint64_t seed_;
    std::atomic<uint64_t> ctr_;
};

std::string g_err;
SeededFactory *g_factory = nullptr;
int g_n = 0, g_K = 0;

template <class F> int guarded(F &&f) {
    try { f(); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
    catch (...) { g_err = "unknown exception"; return -1; }
}

inline size_t ct_words(int size) { return size_t(size) * g_K * (g_n + 1); }

Ciphertext make_ct(const uint64_t *words, int size) {
    // alias ctor sets the hash block (SEAL/seal/ciphertext.h:119-123); copy un-aliases
    Ciphertext alias(*parms, size, const_cast<uint64_t *>(words));
    return Ciphertext(alias);
}

void put_ct(const Ciphertext &ct, uint64_t *words, int expect_size) {
    if (ct.size() != expect_size) throw std::logic_error("unexpected ciphertext size " + to_string(ct.size()));
    memcpy(words, ct.data(), ct_words(expect_size) * 8);
}

Plaintext make_pt(const uint64_t *words, int coeff_count) {
    Plaintext p(coeff_count);
    memcpy(p.data(), words, size_t(coeff_count) * 8);
    return p;
}

ciphertext3D make_tensor(const uint64_t *in, int zd, int xd, int yd, int size = 2) {
    ciphertext3D t(zd, ciphertext2D(xd, vector<Ciphertext>(yd)));
    size_t w = ct_words(size);
    for (int z = 0; z < zd; z++)
        for (int x = 0; x < xd; x++)
            for (int y = 0; y < yd; y++)
                t[z][x][y] = make_ct(in + ((size_t(z) * xd + x) * yd + y) * w, size);
    return t;
}

void put_tensor(const ciphertext3D &t, uint64_t *out, int size = 2) {
    size_t w = ct_words(size);
    size_t i = 0;
    for (auto &plane : t)
        for (auto &row : plane)
            for (auto &ct : row) { put_ct(ct, out + i * w, size); i++; }
}

} // namespace

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }

// Mirrors setParameters (CrCNN/src/globals.cpp:25-56); primes==NULL selects
// coeff_modulus_128(n) exactly as the reference does.
int ref_init(int n, uint64_t t, uint64_t seed, const uint64_t *primes, int nprimes) {
    return guarded([&] {
        if (parms) {  // re-init: drop the previous globals (delParameters order, globals.cpp:113-124)
            delete ev_keys16; delete fraencoder; delete evaluator; delete decryptor;
            delete encryptor; delete keygen; delete context; delete parms;
            parms = nullptr;
        }
        delete g_factory;
        g_factory = new SeededFactory(seed);
        parms = new EncryptionParameters();
        parms->set_poly_modulus("1x^" + to_string(n) + " + 1");
        if (primes && nprimes > 0) {
            vector<SmallModulus> mods;
            for (int i = 0; i < nprimes; i++) mods.emplace_back(primes[i]);
            parms->set_coeff_modulus(mods);
        } else {
            parms->set_coeff_modulus(coeff_modulus_128(n));
        }
        parms->set_plain_modulus(t);
        parms->set_random_generator(g_factory);
        context = new SEALContext(*parms);
        if (!context->qualifiers().parameters_set || !context->qualifiers().enable_ntt)
            throw std::invalid_argument("parameters not valid / NTT not enabled");
        keygen = new KeyGenerator(*context);
        encryptor = new Encryptor(*context, keygen->public_key());
        decryptor = new Decryptor(*context, keygen->secret_key());
        evaluator = new Evaluator(*context);
        fraencoder = new FractionalEncoder(context->plain_modulus(), context->poly_modulus(), 64, 32, 3);
        ev_keys16 = new EvaluationKeys();
        keygen->generate_evaluation_keys(16, *ev_keys16);
        g_n = n;
        g_K = static_cast<int>(context->coeff_modulus().size());
    });
}

int ref_n() { return g_n; }
int ref_K() { return g_K; }
uint64_t ref_t() { return context ? context->plain_modulus().value() : 0; }
int ref_primes(uint64_t *out) {
    for (int i = 0; i < g_K; i++) out[i] = context->coeff_modulus()[i].value();
    return g_K;
}
int ref_fast_plain_lift() { return context->qualifiers().enable_fast_plain_lift ? 1 : 0; }

// Evaluation keys: keys_[0][i] is a Ciphertext of size 2*digits_i
// (SEAL/seal/keygenerator.cpp:198-215).  Copies them back to back, each in
// SEAL layout [size_i][K][n+1].  sizes[i] receives size_i.  out may be NULL.
long ref_evk(uint64_t *out, int *sizes) {
    long words = 0;
    const auto &row = ev_keys16->data()[0];
    for (int i = 0; i < (int)row.size(); i++) {
        int sz = row[i].size();
        if (sizes) sizes[i] = sz;
        if (out) memcpy(out + words, row[i].data(), ct_words(sz) * 8);
        words += (long)ct_words(sz);
    }
    return words;
}
int ref_evk_dbc() { return ev_keys16->decomposition_bit_count(); }

// Replace ev_keys16 by caller-supplied key material (same layout ref_evk() returns).  Built by
// writing SEAL's own EvaluationKeys stream format (SEAL/seal/evaluationkeys.cpp:8-39,
// ciphertext.cpp:103-113) and calling EvaluationKeys::load, so no private member is touched.
int ref_set_evk(const uint64_t *words, const int *sizes, int dbc) {
    return guarded([&] {
        std::stringstream ss(std::ios::in | std::ios::out | std::ios::binary);
        auto hash = parms->hash_block();
        ss.write(reinterpret_cast<const char *>(&hash), sizeof(hash));
        int32_t dbc32 = dbc, dim1 = 1, dim2 = g_K;
        ss.write(reinterpret_cast<const char *>(&dbc32), 4);
        ss.write(reinterpret_cast<const char *>(&dim1), 4);
        ss.write(reinterpret_cast<const char *>(&dim2), 4);
        size_t off = 0;
        for (int i = 0; i < g_K; i++) {
            Ciphertext ct = make_ct(words + off, sizes[i]);
            ct.save(ss);
            off += ct_words(sizes[i]);
        }
        ss.seekg(0);
        ev_keys16->load(ss);
    });
}

// FractionalEncoder(t, x^n+1, 64, 32, 3).encode(v)  (CrCNN/src/globals.cpp:52)
int ref_encode(double v, uint64_t *out /* n+1 words */, int *coeff_count) {
    return guarded([&] {
        Plaintext p = fraencoder->encode(v);
        memset(out, 0, size_t(g_n + 1) * 8);
        memcpy(out, p.data(), size_t(p.coeff_count()) * 8);
        if (coeff_count) *coeff_count = p.coeff_count();
    });
}
int ref_decode(const uint64_t *words, int coeff_count, double *v) {
    return guarded([&] { *v = fraencoder->decode(make_pt(words, coeff_count)); });
}

// encryptImage's inner statement (CrCNN/src/globals.cpp:134) for `count` values.
int ref_encrypt_values(const float *vals, int count, uint64_t *out) {
    return guarded([&] {
        for (int i = 0; i < count; i++) {
            Ciphertext ct;
            encryptor->encrypt(fraencoder->encode(vals[i]), ct);
            put_ct(ct, out + size_t(i) * ct_words(2), 2);
        }
    });
}

// decryptImage's inner statements (CrCNN/src/globals.cpp:221-222) + noise budget.
int ref_decrypt_values(const uint64_t *cts, int count, int size, double *vals, int *budgets,
                       uint64_t *plain_out /* optional count*(n+1) */) {
    return guarded([&] {
        for (int i = 0; i < count; i++) {
            Ciphertext ct = make_ct(cts + size_t(i) * ct_words(size), size);
            Plaintext p;
            decryptor->decrypt(ct, p);
            if (vals) vals[i] = fraencoder->decode(p);
            if (budgets) budgets[i] = decryptor->invariant_noise_budget(ct);
            if (plain_out) {
                memset(plain_out + size_t(i) * (g_n + 1), 0, size_t(g_n + 1) * 8);
                memcpy(plain_out + size_t(i) * (g_n + 1), p.data(), size_t(p.coeff_count()) * 8);
            }
        }
    });
}

// ---- Evaluator-level operations on raw buffers (in place unless noted) ----

int ref_ct_transform(uint64_t *cts, int count, int size, int inverse) {
    return guarded([&] {
        for (int i = 0; i < count; i++) {
            uint64_t *w = cts + size_t(i) * ct_words(size);
            Ciphertext ct = make_ct(w, size);
            if (inverse) evaluator->transform_from_ntt(ct); else evaluator->transform_to_ntt(ct);
            put_ct(ct, w, size);
        }
    });
}

// Evaluator::transform_to_ntt(Plaintext&): out receives K*(n+1) words.
int ref_plain_to_ntt(const uint64_t *plain, int coeff_count, uint64_t *out) {
    return guarded([&] {
        Plaintext p = make_pt(plain, coeff_count);
        evaluator->transform_to_ntt(p);
        memcpy(out, p.data(), size_t(g_K) * (g_n + 1) * 8);
    });
}

int ref_multiply_plain_ntt(uint64_t *cts, int count, int size, const uint64_t *plain_ntt) {
    return guarded([&] {
        Plaintext p = make_pt(plain_ntt, g_K * (g_n + 1));
        for (int i = 0; i < count; i++) {
            uint64_t *w = cts + size_t(i) * ct_words(size);
            Ciphertext ct = make_ct(w, size);
            evaluator->multiply_plain_ntt(ct, p);
            put_ct(ct, w, size);
        }
    });
}

// op: 0 = multiply_plain (generic / constant branch chosen by coeff_count as SEAL does),
//     1 = add_plain, 2 = sub_plain
int ref_plain_op(uint64_t *cts, int count, int size, const uint64_t *plain, int coeff_count, int op) {
    return guarded([&] {
        Plaintext p = make_pt(plain, coeff_count);
        for (int i = 0; i < count; i++) {
            uint64_t *w = cts + size_t(i) * ct_words(size);
            Ciphertext ct = make_ct(w, size);
            if (op == 0) evaluator->multiply_plain(ct, p);
            else if (op == 1) evaluator->add_plain(ct, p);
            else evaluator->sub_plain(ct, p);
            put_ct(ct, w, size);
        }
    });
}

int ref_add_many(const uint64_t *cts, int count, int size, uint64_t *out) {
    return guarded([&] {
        vector<Ciphertext> v;
        for (int i = 0; i < count; i++) v.push_back(make_ct(cts + size_t(i) * ct_words(size), size));
        Ciphertext dst;
        evaluator->add_many(v, dst);
        put_ct(dst, out, size);
    });
}

// Evaluator::square on size-2 inputs -> size-3 outputs.
int ref_square(const uint64_t *in, int count, uint64_t *out) {
    return guarded([&] {
        for (int i = 0; i < count; i++) {
            Ciphertext ct = make_ct(in + size_t(i) * ct_words(2), 2);
            evaluator->square(ct);
            put_ct(ct, out + size_t(i) * ct_words(3), 3);
        }
    });
}

// Evaluator::relinearize(size 3 -> 2) with ev_keys16.
int ref_relinearize(const uint64_t *in, int count, uint64_t *out) {
    return guarded([&] {
        for (int i = 0; i < count; i++) {
            Ciphertext ct = make_ct(in + size_t(i) * ct_words(3), 3);
            evaluator->relinearize(ct, *ev_keys16);
            put_ct(ct, out + size_t(i) * ct_words(2), 2);
        }
    });
}

// ---- CrCNN layers.  Tensors are [z][x][y] of size-2 cts; weights are floats
// encoded exactly as CnnBuilder does (CrCNN/src/cnnBuilder.cpp:25-105). ----

int ref_conv_forward(const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf, int nf,
                     int th_count, const float *w, const float *b, uint64_t *out) {
    return guarded([&] {
        plaintext4D ew(nf, plaintext3D(zd, plaintext2D(xf, vector<Plaintext>(yf))));
        vector<Plaintext> eb(nf);
        size_t k = 0;
        for (int n = 0; n < nf; n++) {
            for (int z = 0; z < zd; z++)
                for (int i = 0; i < xf; i++)
                    for (int j = 0; j < yf; j++) ew[n][z][i][j] = fraencoder->encode(w[k++]);
            eb[n] = fraencoder->encode(b[n]);
        }
        ConvolutionalLayer layer("conv", xd, yd, zd, xs, ys, xf, yf, nf, th_count, ew, eb);
        put_tensor(layer.forward(make_tensor(in, zd, xd, yd)), out);
    });
}

int ref_fc_forward(const uint64_t *in, int in_dim, int out_dim, int th_count, const float *w, const float *b,
                   uint64_t *out) {
    return guarded([&] {
        plaintext2D ew(out_dim, vector<Plaintext>(in_dim));
        vector<Plaintext> eb(out_dim);
        size_t k = 0;
        for (int i = 0; i < out_dim; i++) {
            for (int j = 0; j < in_dim; j++) ew[i][j] = fraencoder->encode(w[k++]);
            eb[i] = fraencoder->encode(b[i]);
        }
        FullyConnectedLayer layer("fc", in_dim, out_dim, th_count, ew, eb);
        put_tensor(layer.forward(make_tensor(in, 1, in_dim, 1)), out);
    });
}

// FullyConnectedLayer::reshapeInput on a [zd][xd][yd] tensor followed by forward
// (CrCNN/src/fullyConnectedLayer.cpp:38-56, 113-168).
int ref_fc_forward_3d(const uint64_t *in, int zd, int xd, int yd, int out_dim, int th_count, const float *w,
                      const float *b, uint64_t *out) {
    return guarded([&] {
        int in_dim = zd * xd * yd;
        plaintext2D ew(out_dim, vector<Plaintext>(in_dim));
        vector<Plaintext> eb(out_dim);
        size_t k = 0;
        for (int i = 0; i < out_dim; i++) {
            for (int j = 0; j < in_dim; j++) ew[i][j] = fraencoder->encode(w[k++]);
            eb[i] = fraencoder->encode(b[i]);
        }
        FullyConnectedLayer layer("fc", in_dim, out_dim, th_count, ew, eb);
        put_tensor(layer.forward(make_tensor(in, zd, xd, yd)), out);
    });
}

int ref_pool_forward(const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf, int avg,
                     uint64_t *out) {
    return guarded([&] {
        if (avg) {
            AvgPoolingLayer layer("avgpool", xd, yd, zd, xs, ys, xf, yf);
            put_tensor(layer.forward(make_tensor(in, zd, xd, yd)), out);
        } else {
            PoolingLayer layer("pool", xd, yd, zd, xs, ys, xf, yf);
            put_tensor(layer.forward(make_tensor(in, zd, xd, yd)), out);
        }
    });
}

// mean / invstd are the already-transformed floats CnnBuilder would encode
// (invstd = 1/sqrt(var+1e-5), CrCNN/src/cnnBuilder.cpp:97-99).
int ref_bn_forward(const uint64_t *in, int zd, int xd, int yd, const float *mean, const float *invstd,
                   uint64_t *out) {
    return guarded([&] {
        vector<Plaintext> em(zd), ev(zd);
        for (int i = 0; i < zd; i++) { em[i] = fraencoder->encode(mean[i]); ev[i] = fraencoder->encode(invstd[i]); }
        BatchNormLayer layer("bn", zd, em, ev);
        put_tensor(layer.forward(make_tensor(in, zd, xd, yd)), out);
    });
}

int ref_square_forward(const uint64_t *in, int zd, int xd, int yd, int th_count, uint64_t *out) {
    return guarded([&] {
        SquareLayer layer("square", th_count);
        put_tensor(layer.forward(make_tensor(in, zd, xd, yd)), out);
    });
}

// ---- steady-state timing of the weighted-sum layers.  The reference encodes its weights once (CnnBuilder) and transforms them to
// NTT form lazily inside the FIRST forward (transform_kernel_to_ntt, CrCNN/src/convolutionalLayer.cpp:149-168, 195;
// fullyConnectedLayer.cpp:129-131); every later image skips both.  times[0] = FractionalEncoder::encode of all weights,
// times[1] = first forward (includes the weight NTTs), times[2] = mean of `reps` further forwards of the SAME layer object
// (what the reference pays per image).  Seconds, wall clock.  out may be null.
static double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

int ref_conv_forward_timed(const uint64_t *in, int xd, int yd, int zd, int xs, int ys, int xf, int yf, int nf,
                           int th_count, const float *w, const float *b, int reps, double *times, uint64_t *out) {
    return guarded([&] {
        double t0 = now_s();
        plaintext4D ew(nf, plaintext3D(zd, plaintext2D(xf, vector<Plaintext>(yf))));
        vector<Plaintext> eb(nf);
        size_t k = 0;
        for (int n = 0; n < nf; n++) {
            for (int z = 0; z < zd; z++)
                for (int i = 0; i < xf; i++)
                    for (int j = 0; j < yf; j++) ew[n][z][i][j] = fraencoder->encode(w[k++]);
            eb[n] = fraencoder->encode(b[n]);
        }
        ConvolutionalLayer layer("conv", xd, yd, zd, xs, ys, xf, yf, nf, th_count, ew, eb);
        double t1 = now_s();
        ciphertext3D x = make_tensor(in, zd, xd, yd);
        ciphertext3D y = layer.forward(x);
        double t2 = now_s();
        for (int r = 0; r < reps; r++) y = layer.forward(x);
        double t3 = now_s();
        times[0] = t1 - t0; times[1] = t2 - t1; times[2] = reps > 0 ? (t3 - t2) / reps : 0.0;
        if (out) put_tensor(y, out);
    });
}

int ref_fc_forward_timed(const uint64_t *in, int in_dim, int out_dim, int th_count, const float *w, const float *b,
                         int reps, double *times, uint64_t *out) {
    return guarded([&] {
        double t0 = now_s();
        plaintext2D ew(out_dim, vector<Plaintext>(in_dim));
        vector<Plaintext> eb(out_dim);
        size_t k = 0;
        for (int i = 0; i < out_dim; i++) {
            for (int j = 0; j < in_dim; j++) ew[i][j] = fraencoder->encode(w[k++]);
            eb[i] = fraencoder->encode(b[i]);
        }
        FullyConnectedLayer layer("fc", in_dim, out_dim, th_count, ew, eb);
        double t1 = now_s();
        ciphertext3D x = make_tensor(in, 1, in_dim, 1);
        ciphertext3D y = layer.forward(x);
        double t2 = now_s();
        for (int r = 0; r < reps; r++) y = layer.forward(x);
        double t3 = now_s();
        times[0] = t1 - t0; times[1] = t2 - t1; times[2] = reps > 0 ? (t3 - t2) / reps : 0.0;
        if (out) put_tensor(y, out);
    });
}

// The key holder's material for the re-encryption step: the secret key as the Decryptor keeps it (NTT form,
// [K][n+1] words, SEAL/seal/decryptor.cpp:60-70) and the public key ([2][K][n+1], NTT form, SEAL/seal/encryptor.cpp:60-75).
int ref_keys(uint64_t *sk_out, uint64_t *pk_out) {
    return guarded([&] {
        const size_t w = size_t(g_K) * (g_n + 1);
        if (sk_out) memcpy(sk_out, keygen->secret_key().data().data(), w * 8);
        if (pk_out) memcpy(pk_out, keygen->public_key().data().data(), 2 * w * 8);
    });
}

// What Network::forward does to ONE ciphertext before layer 6 (CrCNN/src/network.cpp:30-33 through decryptImage / encryptImage,
// CrCNN/src/globals.cpp:127-142, 207-226): decrypt, decode to a float (floatCube), encode again -- the plaintext the fresh
// ciphertext is made of.  plain_out: n+1 words.
int ref_reencode(const uint64_t *ct, uint64_t *plain_out, double *value) {
    return guarded([&] {
        Ciphertext c = make_ct(ct, 2);
        Plaintext p;
        decryptor->decrypt(c, p);
        float f = (float)fraencoder->decode(p);
        Plaintext q = fraencoder->encode(f);
        memset(plain_out, 0, size_t(g_n + 1) * 8);
        memcpy(plain_out, q.data(), size_t(q.coeff_count()) * 8);
        if (value) *value = f;
    });
}

} // extern "C"
