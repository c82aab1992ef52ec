// Drives the C++17 layer classes (crcnn_b200/cpp/crcnn_b200.hpp, stand-in value types) exactly the
// way a CrCNN program drives the reference's: build layers from plaintext parameters, push them into
// a Network, call forward on a ciphertext3D.  Reads the case from a binary file written by
// tests/test_gpu_cpp_dropin.py and writes the outputs next to it; Python byte-compares with the oracle.
//
// file format (little endian): int32 n, K, u64 t, u64 q[K], then the evk (int32 dbc, int32 sizes[K], words),
// then input ciphertexts (2 x 4 x 4 of size 2), then plaintext records in Plaintext::save format.
#include <fstream>
#include <iostream>
#include <sstream>
#include "../../crcnn_b200/cpp/crcnn_b200.hpp"

using namespace crcnn_b200;
using namespace std;

template <class T> T rd(istream &s) { T v; s.read(reinterpret_cast<char *>(&v), sizeof(T)); return v; }

int main(int argc, char **argv) {
    if (argc != 3) { cerr << "usage: dropin_compat_test <case.bin> <out.bin>\n"; return 2; }
    ifstream in(argv[1], ios::binary);
    int n = rd<int32_t>(in), K = rd<int32_t>(in);
    uint64_t t = rd<uint64_t>(in);
    vector<uint64_t> q(K);
    for (auto &x : q) x = rd<uint64_t>(in);
    try {
        Runtime &rt = Runtime::get();
        rt.init(n, q, t);
        int dbc = rd<int32_t>(in);
        vector<int> sizes(K);
        size_t words = 0;
        for (auto &s : sizes) { s = rd<int32_t>(in); words += (size_t)s * K * (n + 1); }
        vector<uint64_t> evk(words);
        in.read(reinterpret_cast<char *>(evk.data()), words * 8);
        rt.setEvaluationKeys(evk.data(), sizes.data(), dbc);

        ciphertext3D x(2, ciphertext2D(4, vector<Ciphertext>(4)));
        for (auto &pl : x) for (auto &row : pl) for (auto &ct : row) ct.load(in);

        // conv(2ch 4x4 -> 3 filters 2x2) -> avgpool 2x2/1 -> bn -> square -> fc(12 -> 4), parameters from the stream
        Network net;
        net.getLayers().push_back(shared_ptr<Layer>(new ConvolutionalLayer("conv", 4, 4, 2, 1, 1, 2, 2, 3, 40, &in)));
        net.getLayers().push_back(shared_ptr<Layer>(new AvgPoolingLayer("pool", 3, 3, 3, 1, 1, 2, 2)));
        net.getLayers().push_back(shared_ptr<Layer>(new BatchNormLayer("bn", 3, &in)));
        net.getLayers().push_back(shared_ptr<Layer>(new SquareLayer("act", 50)));
        net.getLayers().push_back(shared_ptr<Layer>(new FullyConnectedLayer("fc", 12, 4, 40, &in)));

        ofstream out(argv[2], ios::binary);
        // 1. whole network, device resident
        ciphertext3D y = net.forward(x);
        for (auto &pl : y) for (auto &row : pl) for (auto &ct : row) ct.save(out);
        // 2. layer by layer through the reference's by-value signature
        ciphertext3D z = x;
        for (int i = 0; i < net.getNumLayers(); i++) z = net.getLayer(i)->forward(z);
        for (auto &pl : z) for (auto &row : pl) for (auto &ct : row) ct.save(out);
        // 3. save/load round trip of the encoded parameters, then forward again
        stringstream ss(ios::in | ios::out | ios::binary);
        net.getLayer(0)->savePlaintextParameters(&ss);
        ConvolutionalLayer conv2("conv", 4, 4, 2, 1, 1, 2, 2, 3, 40, &ss);
        ciphertext3D c1 = net.getLayer(0)->forward(x), c2 = conv2.forward(x);
        bool same = true;
        for (size_t a = 0; a < c1.size(); a++) for (size_t b = 0; b < c1[a].size(); b++) for (size_t c = 0; c < c1[a][b].size(); c++)
            same = same && memcmp(c1[a][b][c].data(), c2[a][b][c].data(), rt.ct_words() * 8) == 0;
        // 4. error behaviour: wrong geometry must throw std::invalid_argument like SEAL's checks do
        bool threw = false;
        try { ConvolutionalLayer bad("conv", 5, 5, 2, 1, 1, 2, 2, 3, 40, conv2.filters, conv2.biases); bad.forward(x); }
        catch (const invalid_argument &) { threw = true; }
        cout << "avgpool_div_nonzero " << (static_cast<AvgPoolingLayer *>(net.getLayer(1).get())->div_factor[n - 1] != 0) << "\n";
        cout << "saveload_same " << same << "\nbad_geometry_throws " << threw << "\nOK\n";
    } catch (const exception &e) {
        cout << "EXCEPTION " << e.what() << "\n";
        return 1;
    }
    return 0;
}
