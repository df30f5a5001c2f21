// BASELINE.json config 5 from C++ (the twin of profiles/mixed_pipeline.py): the mixed enc / ct_mul / dec job over ALL GPUs of the box
// through pvacb::Group (include/pvacb.hpp over the C ABI) -- one process, one host thread per device inside the library, keys
// replicated over NVLink, items sharded by contiguous index ranges with the RNG streams of the global indices.
//
//   group_pipeline [items = 65536] [tile pairs (global) = 4096] [gpus = all]
//
// Items come in pairs (2j, 2j+1): plaintexts derived from the item index, both encrypted, multiplied, decrypted; only the 16-byte
// decrypts leave the devices. Every product is checked against a*b mod p on the host, and a digest of all decrypts plus the commit_ct
// digests of the first tile's products (= every ciphertext byte) is printed: it is the same for any number of GPUs.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pvacb.hpp"

static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main(int argc, char** argv) {
    const uint64_t items = argc > 1 ? strtoull(argv[1], nullptr, 10) : 65536;
    const uint64_t tile = argc > 2 ? strtoull(argv[2], nullptr, 10) : 4096;
    int ngpu = argc > 3 ? atoi(argv[3]) : 0;
    if (ngpu <= 0) {
        // as many as the box has: probe by creating single-device groups until one fails
        for (ngpu = 0; ngpu < 16; ngpu++) { pvacb_ctx* c = nullptr; if (pvacb_ctx_create(ngpu, &c)) break; pvacb_ctx_destroy(c); }
    }
    if (ngpu < 1) { fprintf(stderr, "no sm_100-class GPU\n"); return 2; }
    std::vector<int> devs;
    for (int k = 0; k < ngpu; k++) devs.push_back(k);
    try {
        pvacb::Group g(devs, PVACB_PRF_LIVE);
        uint8_t key[32], seed[32];
        for (int i = 0; i < 32; i++) { key[i] = (uint8_t)(37 * i + 11); seed[i] = (uint8_t)(i + 1); }
        g.set_tape(PVACB_TAPE_CHACHA20, key);        // fixed tape key and keygen seed: the printed digest is reproducible
        g.keygen(nullptr, seed);
        auto rates = g.tune_export();
        const uint64_t pairs = items / 2;
        uint64_t bad = 0, checked = 0, digest = 0;
        double t_enc = 0, t_mul = 0, t_dec = 0, t_host = 0;
        auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
        auto t0 = std::chrono::steady_clock::now();
        for (uint64_t f0 = 0; f0 < pairs; f0 += tile) {                      // a tile is a GLOBAL batch: the group shards it over its GPUs
            const uint64_t c = std::min<uint64_t>(tile, pairs - f0);
            std::vector<uint64_t> va(c), vb(c);
            for (uint64_t j = 0; j < c; j++) { va[j] = mix64(2 * (f0 + j) + 0x1234); vb[j] = mix64(2 * (f0 + j) + 0x1235); }
            // one seed per tile and operand: (seed, global item index inside the tile) never repeats across tiles
            double q0 = now();
            pvacb::ShardedCiphers A = g.enc_value(va, 3 * (f0 / tile) + 1), B = g.enc_value(vb, 3 * (f0 / tile) + 2);
            double q1 = now();
            pvacb::ShardedCiphers P = g.ct_mul(A, B, 3 * (f0 / tile) + 3);
            double q2 = now();
            std::vector<pvacb::Fp> d = g.dec_value(P);
            double q3 = now();
            t_enc += q1 - q0; t_mul += q2 - q1; t_dec += q3 - q2;
            if (f0 == 0)                                                    // ciphertext bytes too: commit_ct digests of the first tile's products
                for (const auto& cm : g.commit_ct(P))
                    for (int q = 0; q < 4; q++) { uint64_t w; memcpy(&w, cm.data() + 8 * q, 8); digest = mix64(digest ^ w); }
            for (uint64_t j = 0; j < c; j++) {
                unsigned __int128 p = (unsigned __int128)va[j] * vb[j];
                uint64_t lo = (uint64_t)p, hi = (uint64_t)(p >> 64);
                const uint64_t top = hi >> 63;           // fold bit 127: 2^127 = 1 (mod p); a product of two 64-bit values is < 2p
                hi &= 0x7FFFFFFFFFFFFFFFull;
                lo += top;
                if (lo < top) hi++;
                bad += (d[j].lo != lo || d[j].hi != hi);
                digest = mix64(digest ^ d[j].lo) + mix64(d[j].hi + j + f0);
            }
            checked += c;
            t_host += now() - q3;
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("{\"gpus\": %d, \"items\": %llu, \"tile_pairs\": %llu, \"products_checked\": %llu, \"mismatches\": %llu, \"seconds\": %.3f, \"items_per_s\": %.0f, "
               "\"decrypt_digest\": \"%016llx\", \"export_gbs_direct\": %.1f, \"export_gbs_relayed\": %.1f, \"seconds_enc\": %.3f, \"seconds_mul\": %.3f, \"seconds_dec\": %.3f, \"seconds_host_check\": %.3f}\n",
               ngpu, (unsigned long long)items, (unsigned long long)tile, (unsigned long long)checked, (unsigned long long)bad, secs, items / secs,
               (unsigned long long)digest, rates.first, rates.second, t_enc, t_mul, t_dec, t_host);
        return bad ? 1 : 0;
    } catch (const pvacb::Error& e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
}
