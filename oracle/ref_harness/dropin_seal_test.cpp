// TEST INFRASTRUCTURE ONLY.  The strongest drop-in check: ONE process holds both the UNMODIFIED
// reference layer classes (global namespace, CrCNN/src/*.cpp linked from /root/reference) and the
// B200 layer classes (namespace crcnn_b200, -DCRCNN_WITH_SEAL so they take real seal::Ciphertext /
// seal::Plaintext / seal::EvaluationKeys), feeds both the same SEAL-encrypted image and the same
// FractionalEncoder-encoded weights, and memcmp()s every layer's output ciphertexts, then decrypts
// with SEAL and compares scores, argmax and invariant noise budgets.
//
// Built by oracle/Makefile.ref into oracle/_ref/dropin_seal_test (needs /root/reference at build
// time; the binary travels to the GPU box).  Run:  dropin_seal_test [n] [t]
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <random>

// reference classes (global namespace)
#include "globals.h"
#include "convolutionalLayer.h"
#include "fullyConnectedLayer.h"
#include "poolingLayer.h"
#include "avgPoolingLayer.h"
#include "batchNormLayer.h"
#include "squareLayer.h"
#include "network.h"

// B200 classes (namespace crcnn_b200)
#define CRCNN_WITH_SEAL
#include "../../crcnn_b200/cpp/crcnn_b200.hpp"

extern "C" int ref_init(int n, uint64_t t, uint64_t seed, const uint64_t *primes, int nprimes);
extern "C" const char *ref_last_error();

namespace gpu = crcnn_b200;

static bool same(const ciphertext3D &a, const ciphertext3D &b, const char *what) {
    if (a.size() != b.size() || a[0].size() != b[0].size() || a[0][0].size() != b[0][0].size()) {
        std::cout << what << ": shape differs\n";
        return false;
    }
    for (size_t z = 0; z < a.size(); z++)
        for (size_t x = 0; x < a[z].size(); x++)
            for (size_t y = 0; y < a[z][x].size(); y++) {
                const Ciphertext &p = a[z][x][y], &q = b[z][x][y];
                size_t words = (size_t)p.size() * p.coeff_mod_count() * p.poly_coeff_count();
                if (p.size() != q.size() || std::memcmp(p.data(), q.data(), words * 8) != 0 ||
                    const_cast<Ciphertext &>(p).hash_block() != const_cast<Ciphertext &>(q).hash_block()) {
                    std::cout << what << ": ciphertext [" << z << "][" << x << "][" << y << "] differs\n";
                    return false;
                }
            }
    std::cout << what << ": bit-identical (" << a.size() * a[0].size() * a[0][0].size() << " ciphertexts)\n";
    return true;
}

int main(int argc, char **argv) {
    int n = argc > 1 ? atoi(argv[1]) : 4096;
    uint64_t t = argc > 2 ? strtoull(argv[2], nullptr, 0) : (1ULL << 20);
    if (ref_init(n, t, 2024, nullptr, 0) != 0) { std::cout << "ref_init: " << ref_last_error() << "\n"; return 1; }
    try {
        gpu::Runtime::get().init(*context);
        gpu::Runtime::get().setEvaluationKeys(*ev_keys16);

        std::mt19937 rng(7);
        std::uniform_real_distribution<float> pix(-0.4242f, 2.8215f), wd(-0.5f, 0.5f);
        // a small network with every layer type: 1x7x7 -> conv(3 filters 3x3, stride 1) -> avgpool 2x2/1 -> bn
        //   -> conv(4 filters 2x2, stride 1) -> square -> sum-pool 2x2/1 -> fc(16 -> 10)... sized to run in seconds on the CPU
        std::vector<float> image(49);
        for (auto &p : image) p = pix(rng);
        ciphertext3D x = encryptImage(image, 1, 7, 7);  // CrCNN/src/globals.cpp:127-142 (SEAL encryption, client side)

        auto enc4 = [&](int nf, int zd, int xf, int yf) {
            plaintext4D w(nf, plaintext3D(zd, plaintext2D(xf, std::vector<Plaintext>(yf))));
            for (auto &a : w) for (auto &b : a) for (auto &c : b) for (auto &d : c) d = fraencoder->encode(wd(rng));
            return w;
        };
        auto enc1 = [&](int k, float lo, float hi) {
            std::uniform_real_distribution<float> d(lo, hi);
            std::vector<Plaintext> v(k);
            for (auto &p : v) p = fraencoder->encode(d(rng));
            return v;
        };
        plaintext4D w1 = enc4(3, 1, 3, 3), w2 = enc4(4, 3, 2, 2);
        std::vector<Plaintext> b1 = enc1(3, -0.5f, 0.5f), b2 = enc1(4, -0.5f, 0.5f), mean = enc1(3, -0.3f, 0.3f), invstd = enc1(3, 0.7f, 1.5f);
        plaintext2D wf(10, std::vector<Plaintext>(16));
        for (auto &r : wf) for (auto &p : r) p = fraencoder->encode(wd(rng));
        std::vector<Plaintext> bf = enc1(10, -0.5f, 0.5f);

        // reference network (deep copies of the parameters: the reference NTT-transforms its own in place)
        plaintext4D w1r = w1, w2r = w2; plaintext2D wfr = wf;
        std::vector<Plaintext> b1r = b1, b2r = b2, meanr = mean, invstdr = invstd, bfr = bf;
        Network ref;
        ref.getLayers().push_back(std::shared_ptr<Layer>(new ConvolutionalLayer("conv1", 7, 7, 1, 1, 1, 3, 3, 3, 3, w1r, b1r)));
        ref.getLayers().push_back(std::shared_ptr<Layer>(new AvgPoolingLayer("pool1", 5, 5, 3, 1, 1, 2, 2)));
        ref.getLayers().push_back(std::shared_ptr<Layer>(new BatchNormLayer("bn1", 3, meanr, invstdr)));
        ref.getLayers().push_back(std::shared_ptr<Layer>(new ConvolutionalLayer("conv2", 4, 4, 3, 1, 1, 2, 2, 4, 4, w2r, b2r)));
        ref.getLayers().push_back(std::shared_ptr<Layer>(new SquareLayer("act1", 4)));
        ref.getLayers().push_back(std::shared_ptr<Layer>(new PoolingLayer("pool2", 3, 3, 4, 1, 1, 2, 2)));
        ref.getLayers().push_back(std::shared_ptr<Layer>(new FullyConnectedLayer("fc", 16, 10, 10, wfr, bfr)));

        gpu::Network net;
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::ConvolutionalLayer("conv1", 7, 7, 1, 1, 1, 3, 3, 3, 3, w1, b1)));
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::AvgPoolingLayer("pool1", 5, 5, 3, 1, 1, 2, 2)));
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::BatchNormLayer("bn1", 3, mean, invstd)));
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::ConvolutionalLayer("conv2", 4, 4, 3, 1, 1, 2, 2, 4, 4, w2, b2)));
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::SquareLayer("act1", 4)));
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::PoolingLayer("pool2", 3, 3, 4, 1, 1, 2, 2)));
        net.getLayers().push_back(std::shared_ptr<gpu::Layer>(new gpu::FullyConnectedLayer("fc", 16, 10, 10, wf, bf)));

        // layer by layer (the reference's Network::forward re-encrypts before layer 6 unconditionally,
        // network.cpp:30; both sides are therefore driven through getLayer(i)->forward)
        bool ok = true;
        ciphertext3D a = x, b = x;
        for (int i = 0; i < ref.getNumLayers(); i++) {
            a = ref.getLayer(i)->forward(a);
            b = net.getLayer(i)->forward(b);
            ok = same(a, b, ref.getLayer(i)->getName().c_str()) && ok;
        }
        // whole network device-resident.  Seven layers > layer_before_reenc = 6: without a re-encryption callback and without an
        // explicit opt-out Network::forward refuses (the reference would re-encrypt here, network.cpp:30)
        bool refused = false;
        try { net.forward(x); } catch (const std::logic_error &) { refused = true; }
        std::cout << "forward without re-encryption policy refused: " << refused << "\n";
        ok = ok && refused;
        net.skip_reencryption = true;
        ciphertext3D c = net.forward(x);
        ok = same(a, c, "Network::forward (device resident, re-encryption skipped)") && ok;
        // the reference's behaviour: decrypt + encrypt before layer 6 (network.cpp:30-33), done by the key holder in the callback.
        // Encryption is randomised, so the re-encrypted tensor is captured and pushed through the REFERENCE's layer 6 for comparison.
        net.skip_reencryption = false;
        ciphertext3D captured;
        int calls = 0;
        net.reencrypt = [&](gpu::ciphertext3D mid) {
            calls++;
            floatCube image = decryptImage(mid);   // CrCNN/src/globals.cpp (SEAL Decryptor + FractionalEncoder::decode)
            captured = encryptImage(image);        // fresh ciphertexts, full noise budget
            return captured;
        };
        ciphertext3D d = net.forward(x);
        ciphertext3D e = captured;
        for (int i = 6; i < ref.getNumLayers(); i++) e = ref.getLayer(i)->forward(e);
        ok = same(e, d, "Network::forward with the re-encryption callback (segments [0,6) + [6,7))") && ok;
        ok = ok && calls == 1 && captured.size() == 4 && captured[0].size() == 2 && captured[0][0].size() == 2;
        int nb_d = decryptor->invariant_noise_budget(d[0][0][0]);
        std::cout << "re-encryption callback calls " << calls << ", budget after re-encrypted tail " << nb_d << " bits\n";
        net.reencrypt = nullptr;
        // the same noise reset ON THE DEVICE: keys uploaded by the key holder, decrypt -> decode -> float -> encode ->
        // encrypt on the GPU.  Ciphertexts are randomised, plaintexts are not: the decrypted scores must equal, coefficient by
        // coefficient, those of the REFERENCE's own Network::forward (which re-encrypts before layer 6 with SEAL on the host).
        net.use_device_reencryption(keygen->secret_key(), keygen->public_key(), 2024);
        ciphertext3D g = net.forward(x);
        ciphertext3D h = ref.forward(x);            // CrCNN/src/network.cpp:22-47, unmodified
        bool plain_equal = true;
        for (int i = 0; i < 10; i++) {
            Plaintext pg, ph;
            decryptor->decrypt(g[0][i][0], pg);
            decryptor->decrypt(h[0][i][0], ph);
            plain_equal = plain_equal && pg == ph;
        }
        int nb_g = decryptor->invariant_noise_budget(g[0][0][0]), nb_h = decryptor->invariant_noise_budget(h[0][0][0]);
        std::cout << "device re-encryption: decrypted plaintexts equal to the reference's Network::forward: " << plain_equal
                  << ", noise budget " << nb_g << " vs " << nb_h << " bits\n";
        ok = ok && plain_equal && nb_g > 0 && std::abs(nb_g - nb_h) <= 1;
        net.reencrypt_dev = nullptr;
        net.skip_reencryption = true;

        floatCube sa = decryptImage(a), sc = decryptImage(c);  // SEAL decryption, client side
        int arg_a = 0, arg_c = 0;
        for (int i = 0; i < 10; i++) {
            if (sa[0][i][0] > sa[0][arg_a][0]) arg_a = i;
            if (sc[0][i][0] > sc[0][arg_c][0]) arg_c = i;
            if (sa[0][i][0] != sc[0][i][0]) ok = false;
        }
        int nb_a = decryptor->invariant_noise_budget(a[0][0][0]), nb_c = decryptor->invariant_noise_budget(c[0][0][0]);
        std::cout << "scores equal, label " << arg_a << " == " << arg_c << ", noise budget " << nb_a << " == " << nb_c << " bits\n";
        ok = ok && arg_a == arg_c && nb_a == nb_c && nb_a > 0;

        // error behaviour: a ciphertext from other parameters is rejected with std::invalid_argument
        bool threw = false;
        try {
            ciphertext3D bad = x;
            const_cast<Ciphertext &>(bad[0][0][0]).hash_block()[0] ^= 1;
            net.getLayer(0)->forward(bad);
        } catch (const std::invalid_argument &) { threw = true; }
        std::cout << "foreign ciphertext rejected: " << threw << "\n";
        ok = ok && threw;
        std::cout << (ok ? "DROPIN OK\n" : "DROPIN FAILED\n");
        return ok ? 0 : 1;
    } catch (const std::exception &e) {
        std::cout << "EXCEPTION " << e.what() << "\n";
        return 1;
    }
}
