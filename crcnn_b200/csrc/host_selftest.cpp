// Host-only dump of everything derive_params() computes, for the CPU test-suite to compare with the
// oracle (tests/test_host_params.py).  No CUDA involved: modarith.cuh compiles as plain C++.
//   host_selftest <n> <t> <q_1> ... <q_K>   -> text on stdout
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "params.h"

using namespace crcnn;

static uint64_t fnv(const std::vector<uint64_t> &v) {
    uint64_t h = 1469598103934665603ULL;
    for (uint64_t x : v) { h ^= x; h *= 1099511628211ULL; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: host_selftest n t q...\n"); return 2; }
    int n = atoi(argv[1]);
    uint64_t t = strtoull(argv[2], nullptr, 0);
    std::vector<uint64_t> q;
    for (int i = 3; i < argc; i++) q.push_back(strtoull(argv[i], nullptr, 0));
    try {
        HostParams hp = derive_params(n, (int)q.size(), q.data(), t);
        const DeviceParams &d = hp.d;
        printf("K %d L %d S %d half %" PRIu64 "\n", d.K, d.L, d.S, d.half);
        for (int s = 0; s < d.K + d.S; s++) {
            printf("slot %d q %" PRIu64 " r0 %" PRIu64 " r1 %" PRIu64 " root %" PRIu64 " w %" PRIu64 " wp %" PRIu64 " iw %" PRIu64 " iwp %" PRIu64 "\n",
                   s, d.tab[s].mod.q, d.tab[s].mod.r0, d.tab[s].mod.r1, hp.roots[s], fnv(hp.w[s]), fnv(hp.wp[s]), fnv(hp.iw[s]), fnv(hp.iwp[s]));
        }
        for (int j = 0; j < d.K; j++)
            printf("prime %d delta %" PRIu64 " rho %" PRIu64 " lift %" PRIu64 "\n", j, d.delta[j], d.rho[j], d.lift_inc[j]);
        // modular arithmetic spot checks against unsigned __int128
        uint64_t x = 0x123456789abcdef1ULL, y = 0xfedcba9876543211ULL;
        for (int s = 0; s < d.K + d.S; s++) {
            const Mod &m = d.tab[s].mod;
            for (int it = 0; it < 1000; it++) {
                x = x * 6364136223846793005ULL + 1442695040888963407ULL;
                y = y * 2862933555777941757ULL + 3037000493ULL;
                uint64_t a = x % (4 * m.q), b = y % m.q;
                uint64_t want = (uint64_t)(((unsigned __int128)a * b) % m.q);
                if (mulmod(a, b, m) != want) { printf("MULMOD MISMATCH\n"); return 1; }
                uint64_t wp = (uint64_t)((((unsigned __int128)b) << 64) / m.q);
                uint64_t lazy = mulshoup_lazy(a, b, wp, m.q);
                if (lazy >= 2 * m.q || lazy % m.q != want) { printf("SHOUP MISMATCH\n"); return 1; }
            }
        }
        // folded reduction of the limb-split GEMM's class sums (modarith.cuh: tcn_fold_reduce) against unsigned __int128
        for (int j = 0; j < d.K; j++) {
            const Mod &m = d.tab[j].mod;
            const TcnFold f = tcn_fold_make(m.q);
            printf("fold %d ok %u T %u delta %u\n", j, f.ok, f.T, f.delta);
            if (!f.ok) continue;
            for (int it = 0; it < 200000; it++) {
                uint32_t s[13];
                const int mode = it % 8;
                for (int w = 0; w < 13; w++) {
                    x = x * 6364136223846793005ULL + 1442695040888963407ULL;
                    const uint32_t r = (uint32_t)(x >> 33);                 // < 2^31
                    s[w] = mode == 0 ? 0x7fffffffu : mode == 1 ? 0u : mode == 2 ? ((x >> 20) & 1 ? 0x7fffffffu : 0u)
                         : mode == 3 ? 7u * 4096u * 255u * 255u - (r & 3) : mode == 4 ? (r >> (x & 31)) : r;
                }
                y = y * 2862933555777941757ULL + 3037000493ULL;
                const uint64_t bias = mode == 0 ? m.q - 1 : (it & 1) ? y % m.q : 0;
                unsigned __int128 z = 0;
                for (int w = 12; w >= 0; w--) z = (z << 8) + s[w];
                const uint64_t want = (uint64_t)((z + bias) % m.q);
                if (tcn_fold_reduce(s, bias, f, m.q) != want) { printf("FOLD MISMATCH\n"); return 1; }
            }
        }
        // three-fold reduction of arbitrary 128-bit values (modarith.cuh: reduce128_fold) for every modulus, against barrett128 and __int128
        for (int sidx = 0; sidx < d.K + d.S; sidx++) {
            const Mod &m = d.tab[sidx].mod;
            const Fold128 f = fold128_make(m.q);
            printf("fold128 %d ok %u delta %u\n", sidx, f.ok, f.delta);
            if (!f.ok) continue;
            for (int it = 0; it < 200000; it++) {
                x = x * 6364136223846793005ULL + 1442695040888963407ULL;
                y = y * 2862933555777941757ULL + 3037000493ULL;
                U128 z;
                const int mode = it % 8;
                z.lo = mode == 0 ? ~0ull : mode == 1 ? 0 : mode == 2 ? m.q - 1 : mode == 3 ? m.q : x;
                z.hi = mode == 0 ? ~0ull : mode == 1 ? 0 : mode == 2 ? 0 : mode == 3 ? (y >> (y & 63)) : mode == 4 ? (y >> 40) : y;
                const uint64_t want = (uint64_t)(((((unsigned __int128)z.hi) << 64) | z.lo) % m.q);
                if (reduce128_fold(z, f) != want || barrett128(z, m) != want) { printf("FOLD128 MISMATCH\n"); return 1; }
            }
        }
        // fractional encoder
        for (int i = 4; i >= 0; i--) {
            double vals[] = {0.0867, -3.25, 0.0, 1.0, 2.8215};
            std::vector<uint32_t> idx; std::vector<uint64_t> val;
            encode_fractional_sparse(vals[i], n, t, idx, val);
            printf("enc %.4f nnz %zu", vals[i], idx.size());
            for (size_t e = 0; e < idx.size(); e++) printf(" %u:%" PRIu64, idx[e], val[e]);
            printf("\n");
        }
        printf("OK\n");
    } catch (const std::exception &e) {
        printf("ERROR %s\n", e.what());
        return 1;
    }
    return 0;
}
