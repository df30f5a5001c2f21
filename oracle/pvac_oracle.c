/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the pvac-hfhe hot path.
 *
 * Never linked into, imported by, or called from the product path (pvac_hfhe_cppbyv_b200/); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Plain C11 (+ unsigned __int128). Written from the behaviour of the reference, each function citing the
 * reference file:line it restates (paths relative to /root/reference/include/pvac/). Parity is PINNED:
 * tests/test_oracle_vs_ref.py checks every function here against the unmodified reference compiled in
 * oracle/_ref (when /root/reference is present) and tests/test_oracle_golden.py checks it against the
 * committed fixtures in tests/golden/ (generated from that reference build by oracle/make_golden.py) and the
 * reference repository's own golden file bounty2_data/{a,b,sum}.ct.
 *
 * Third-party behaviour restated: libstdc++ (GCC 13.3.0) std::unordered_map iteration order, which fixes the
 * edge emission order of ct_mul (ops/arithmetic.hpp:75-101). Its bucket count after reserve(n) comes from
 * std::__detail::_Prime_rehash_policy::_M_next_bkt, queried through oracle/buckets.cpp.
 *
 * Randomness: the reference draws 64-bit words from the OS CSPRNG (core/random.hpp:106-110). Here every
 * draw comes from an explicit SplitMix64 word tape (same stream definition as oracle/ref_shim.cpp and
 * include/pvacb.h) so that results are reproducible.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t lo, hi; } fp_t;

#define MASK63 0x7FFFFFFFFFFFFFFFull
#define ORC_B 337
#define ORC_M_BITS 8192
#define ORC_M_WORDS 128
#define ORC_N_BITS 16384
#define ORC_H_COL_WT 192
#define ORC_X_COL_WT 128
#define ORC_ERR_WT 128
#define ORC_LPN_N 4096
#define ORC_LPN_WORDS 64
#define ORC_LPN_T 16384

extern uint64_t orc_next_bkt(uint64_t n); /* oracle/buckets.cpp */

/* ------------------------------------------------------------------ word tape (core/random.hpp:106-110 replaced) */
static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* Tape kinds (same definitions as pvac_hfhe_cppbyv_b200/csrc/common.cuh): 0 SplitMix64 of a 64-bit state; 1 ChaCha20 -- word k =
 * 64-bit little-endian word (k mod 8) of block (k div 8) under a 256-bit key, block input words 12..15 = (block index, lane,
 * stream id lo, stream id hi), the stream id being the `tape_state` argument of the entry points below; 2 explicit words.
 * The kind, key, lane and words are process-wide settings of this test library (orc_set_tape / _lane / _words). */
static int g_tape_kind = 0;
static uint32_t g_tape_key[8];
static uint32_t g_tape_lane = 0;
static const uint64_t* g_tape_words = 0;
static uint64_t g_tape_nwords = 0;
void orc_set_tape(int kind, const uint8_t* key32) {
    g_tape_kind = kind;
    if (key32) for (int i = 0; i < 8; i++) g_tape_key[i] = (uint32_t)key32[4 * i] | ((uint32_t)key32[4 * i + 1] << 8) | ((uint32_t)key32[4 * i + 2] << 16) | ((uint32_t)key32[4 * i + 3] << 24);
}
void orc_set_tape_lane(uint32_t lane) { g_tape_lane = lane; }
void orc_set_tape_words(const uint64_t* words, uint64_t n) { g_tape_words = words; g_tape_nwords = n; }
static uint32_t rotl32c(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define ORC_QR(a, b, c, d) a += b; d ^= a; d = rotl32c(d, 16); c += d; b ^= c; b = rotl32c(b, 12); a += b; d ^= a; d = rotl32c(d, 8); c += d; b ^= c; b = rotl32c(b, 7);
/* the ChaCha20 block function (D. J. Bernstein; RFC 8439 section 2.3): 20 rounds over constants | key | words 12..15 */
void orc_chacha20_block(const uint32_t key[8], const uint32_t c[4], uint32_t out[16]) {
    uint32_t in[16] = { 0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7], c[0], c[1], c[2], c[3] };
    uint32_t x[16];
    memcpy(x, in, sizeof x);
    for (int r = 0; r < 10; r++) {
        ORC_QR(x[0], x[4], x[8], x[12]) ORC_QR(x[1], x[5], x[9], x[13]) ORC_QR(x[2], x[6], x[10], x[14]) ORC_QR(x[3], x[7], x[11], x[15])
        ORC_QR(x[0], x[5], x[10], x[15]) ORC_QR(x[1], x[6], x[11], x[12]) ORC_QR(x[2], x[7], x[8], x[13]) ORC_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}
typedef struct { uint64_t state; uint64_t draws; } tape_t;
static uint64_t tape_u64(tape_t* t) {
    const uint64_t k = t->draws++;
    if (g_tape_kind == 1) {
        uint32_t c[4] = { (uint32_t)(k >> 3), g_tape_lane ^ (uint32_t)((k >> 3) >> 32), (uint32_t)t->state, (uint32_t)(t->state >> 32) }, o[16];
        orc_chacha20_block(g_tape_key, c, o);
        return (uint64_t)o[2 * (k & 7)] | ((uint64_t)o[2 * (k & 7) + 1] << 32);
    }
    if (g_tape_kind == 2) return k < g_tape_nwords ? g_tape_words[k] : 0;
    return mix64(t->state + (k + 1) * 0x9E3779B97F4A7C15ull);
}
uint64_t orc_item_stream_state(uint64_t batch_seed, uint64_t item) {
    return mix64(batch_seed + 0xD1342543DE82EF95ull * (item + 1));
}

/* ------------------------------------------------------------------ Fp = GF(2^127-1)  (core/field.hpp) */
static const u128 P127 = (((u128)1) << 127) - 1;

static fp_t fp_pack(u128 x) { fp_t r = { (uint64_t)x, (uint64_t)(x >> 64) }; return r; }
static u128 fp_val(fp_t a) { return ((u128)a.hi << 64) | a.lo; }

/* core/field.hpp:26-48 -- any 128-bit value to the canonical residue; p itself maps to 0 */
static fp_t fp_from_words(uint64_t lo, uint64_t hi) {
    u128 x = ((u128)hi << 64) | lo;
    x = (x & P127) + (x >> 127);
    if (x >= P127) x -= P127;
    return fp_pack(x);
}
static fp_t fp_from_u64(uint64_t x) { fp_t r = { x, 0 }; return r; }
/* core/field.hpp:50-56 */
static fp_t fp_add(fp_t a, fp_t b) {
    u128 s = fp_val(a) + fp_val(b); /* < 2^128 for canonical inputs */
    return fp_from_words((uint64_t)s, (uint64_t)(s >> 64));
}
/* core/field.hpp:58-67 */
static fp_t fp_neg(fp_t a) {
    u128 s = P127 - fp_val(a);
    return fp_from_words((uint64_t)s, (uint64_t)(s >> 64));
}
/* core/field.hpp:69-71 */
static fp_t fp_sub(fp_t a, fp_t b) { return fp_add(a, fp_neg(b)); }

/* core/field.hpp:113-213 -- 256-bit schoolbook product, Mersenne folding, canonical result */
static fp_t fp_mul(fp_t a, fp_t b) {
    a = fp_from_words(a.lo, a.hi); b = fp_from_words(b.lo, b.hi); /* the reference result is the canonical residue for any 128-bit inputs */
    u128 p00 = (u128)a.lo * b.lo, p01 = (u128)a.lo * b.hi, p10 = (u128)a.hi * b.lo, p11 = (u128)a.hi * b.hi;
    uint64_t z0 = (uint64_t)p00;
    u128 m1 = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    uint64_t z1 = (uint64_t)m1;
    u128 m2 = (p01 >> 64) + (p10 >> 64) + (uint64_t)p11 + (m1 >> 64);
    uint64_t z2 = (uint64_t)m2;
    uint64_t z3 = (uint64_t)(p11 >> 64) + (uint64_t)(m2 >> 64);
    /* value = low127 + 2^127 * high ; 2^127 == 1 (mod p) */
    u128 low = (((u128)(z1 & MASK63)) << 64) | z0;
    u128 high = ((u128)((z3 << 1) | (z2 >> 63)) << 64) | ((z2 << 1) | (z1 >> 63)); /* bits 127..254 (z3 bit 63 is 0 for canonical inputs) */
    u128 s = low + (high & P127) + (high >> 127);
    s = (s & P127) + (s >> 127);
    if (s >= P127) s -= P127;
    return fp_pack(s);
}
static int fp_is_zero(fp_t a) { return (a.lo | a.hi) == 0; }
static int fp_eq(fp_t a, fp_t b) { return a.lo == b.lo && a.hi == b.hi; }

/* core/field.hpp:229-273 -- a^(p-2); any exponentiation schedule gives the same canonical value */
static fp_t fp_inv(fp_t a) {
    u128 e = P127 - 2;
    fp_t r = fp_from_u64(1), base = a;
    while (e) {
        if (e & 1) r = fp_mul(r, base);
        base = fp_mul(base, base);
        e >>= 1;
    }
    return r;
}
static fp_t fp_pow_u128(fp_t a, u128 e) {
    fp_t r = fp_from_u64(1);
    while (e) {
        if (e & 1) r = fp_mul(r, a);
        a = fp_mul(a, a);
        e >>= 1;
    }
    return r;
}

/* crypto/lpn.hpp:25-37 */
static fp_t hash_to_fp_nonzero(uint64_t lo, uint64_t hi) {
    fp_t r = fp_from_words(lo, hi & MASK63);
    if (fp_is_zero(r)) return fp_from_u64(1);
    return r;
}

/* core/types.hpp:145-155 -- lo word first, then hi word, retry on zero */
static fp_t rand_fp_nonzero(tape_t* t) {
    for (;;) {
        uint64_t lo = tape_u64(t);
        uint64_t hi = tape_u64(t) & MASK63;
        fp_t x = fp_from_words(lo, hi);
        if (!fp_is_zero(x)) return x;
    }
}

/* ------------------------------------------------------------------ SHA-256 (core/hash.hpp:24-191, FIPS 180-4) */
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2 };

typedef struct { uint32_t h[8]; uint64_t len; uint8_t buf[64]; size_t ptr; } sha_t;
static uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha_init(sha_t* s) {
    static const uint32_t iv[8] = { 0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19 };
    memcpy(s->h, iv, sizeof iv);
    s->len = 0; s->ptr = 0;
}
static void sha_block(sha_t* s, const uint8_t* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = s->h[0], b = s->h[1], c = s->h[2], d = s->h[3], e = s->h[4], f = s->h[5], g = s->h[6], h = s->h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
        uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s->h[0] += a; s->h[1] += b; s->h[2] += c; s->h[3] += d; s->h[4] += e; s->h[5] += f; s->h[6] += g; s->h[7] += h;
}
static void sha_update(sha_t* s, const void* data, size_t n) {
    const uint8_t* p = (const uint8_t*)data;
    s->len += n;
    while (n) {
        size_t take = 64 - s->ptr; if (take > n) take = n;
        memcpy(s->buf + s->ptr, p, take);
        s->ptr += take; p += take; n -= take;
        if (s->ptr == 64) { sha_block(s, s->buf); s->ptr = 0; }
    }
}
static void sha_u64le(sha_t* s, uint64_t x) { /* core/hash.hpp:187-191 */
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(x >> (8 * i));
    sha_update(s, b, 8);
}
static void sha_final(sha_t* s, uint8_t out[32]) {
    uint64_t bits = s->len * 8;
    uint8_t pad = 0x80, z = 0, be[8];
    sha_update(s, &pad, 1);
    while (s->ptr != 56) sha_update(s, &z, 1);
    for (int i = 0; i < 8; i++) be[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha_update(s, be, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = s->h[i] >> 24; out[4 * i + 1] = s->h[i] >> 16; out[4 * i + 2] = s->h[i] >> 8; out[4 * i + 3] = s->h[i]; }
}
static uint64_t le64(const uint8_t* p) { uint64_t x = 0; for (int i = 0; i < 8; i++) x |= (uint64_t)p[i] << (8 * i); return x; }

void orc_sha256(const uint8_t* p, size_t n, uint8_t out[32]) { sha_t s; sha_init(&s); sha_update(&s, p, n); sha_final(&s, out); }

/* ------------------------------------------------------------------ AES-256-CTR word stream (crypto/lpn.hpp:41-149, FIPS 197) */
static uint8_t AES_SBOX[256];
static uint32_t AES_T0[256]; /* little-endian column (2s, s, s, 3s) */
static int aes_ready = 0;
static uint8_t gmul(uint8_t a, uint8_t b) { uint8_t r = 0; while (b) { if (b & 1) r ^= a; a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); b >>= 1; } return r; }
static void aes_tables(void) {
    if (aes_ready) return;
    for (int x = 0; x < 256; x++) {
        uint8_t inv = 0;
        if (x) for (int y = 1; y < 256; y++) if (gmul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
        uint8_t s = inv, r = inv;
        for (int k = 0; k < 4; k++) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
        s ^= 0x63;
        AES_SBOX[x] = s;
        AES_T0[x] = (uint32_t)gmul(s, 2) | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)gmul(s, 3) << 24);
    }
    aes_ready = 1;
}
static uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
typedef struct { uint32_t rk[60]; uint64_t ctr; uint64_t buf[2]; int has_buf; uint64_t nout; int patched; } aesctr_t;
/* test hook (the engine's pvacb_debug_set(ctx, 1, word, or_mask)): keystream word `word` of every LPN stream is OR-ed with or_mask, which
 * lets a test reach the rejection branch of bounded() (crypto/lpn.hpp:141-148, p = 2^-61 per row) */
static uint64_t g_patch_word = ~0ull, g_patch_or = 0;
void orc_set_prf_patch(uint64_t word, uint64_t or_mask) { g_patch_word = word; g_patch_or = or_mask; }
static uint32_t subword(uint32_t w) {
    return (uint32_t)AES_SBOX[w & 0xff] | ((uint32_t)AES_SBOX[(w >> 8) & 0xff] << 8) | ((uint32_t)AES_SBOX[(w >> 16) & 0xff] << 16) | ((uint32_t)AES_SBOX[w >> 24] << 24);
}
/* key schedule, words are little-endian loads of the key bytes (column c = bytes 4c..4c+3) */
static void aesctr_init(aesctr_t* a, const uint8_t key[32], uint64_t nonce) {
    aes_tables();
    for (int i = 0; i < 8; i++) a->rk[i] = (uint32_t)key[4 * i] | ((uint32_t)key[4 * i + 1] << 8) | ((uint32_t)key[4 * i + 2] << 16) | ((uint32_t)key[4 * i + 3] << 24);
    uint32_t rcon = 1;
    for (int i = 8; i < 60; i++) {
        uint32_t t = a->rk[i - 1];
        if (i % 8 == 0) { t = subword((t >> 8) | (t << 24)) ^ rcon; rcon = gmul((uint8_t)rcon, 2); }
        else if (i % 8 == 4) t = subword(t);
        a->rk[i] = a->rk[i - 8] ^ t;
    }
    a->ctr = nonce; a->has_buf = 0; a->nout = 0; a->patched = 0;
}
/* one block: input = LE64(ctr) || 0^8 (counter in the low lane only, crypto/lpn.hpp:84,104) */
static void aesctr_block(aesctr_t* a, uint64_t out[2]) {
    uint32_t s0 = (uint32_t)a->ctr ^ a->rk[0], s1 = (uint32_t)(a->ctr >> 32) ^ a->rk[1], s2 = a->rk[2], s3 = a->rk[3];
    a->ctr++;
    for (int r = 1; r < 14; r++) {
        const uint32_t* k = a->rk + 4 * r;
        uint32_t t0 = AES_T0[s0 & 0xff] ^ rotl32(AES_T0[(s1 >> 8) & 0xff], 8) ^ rotl32(AES_T0[(s2 >> 16) & 0xff], 16) ^ rotl32(AES_T0[s3 >> 24], 24) ^ k[0];
        uint32_t t1 = AES_T0[s1 & 0xff] ^ rotl32(AES_T0[(s2 >> 8) & 0xff], 8) ^ rotl32(AES_T0[(s3 >> 16) & 0xff], 16) ^ rotl32(AES_T0[s0 >> 24], 24) ^ k[1];
        uint32_t t2 = AES_T0[s2 & 0xff] ^ rotl32(AES_T0[(s3 >> 8) & 0xff], 8) ^ rotl32(AES_T0[(s0 >> 16) & 0xff], 16) ^ rotl32(AES_T0[s1 >> 24], 24) ^ k[2];
        uint32_t t3 = AES_T0[s3 & 0xff] ^ rotl32(AES_T0[(s0 >> 8) & 0xff], 8) ^ rotl32(AES_T0[(s1 >> 16) & 0xff], 16) ^ rotl32(AES_T0[s2 >> 24], 24) ^ k[3];
        s0 = t0; s1 = t1; s2 = t2; s3 = t3;
    }
    const uint32_t* k = a->rk + 56;
    uint32_t o0 = ((uint32_t)AES_SBOX[s0 & 0xff] | ((uint32_t)AES_SBOX[(s1 >> 8) & 0xff] << 8) | ((uint32_t)AES_SBOX[(s2 >> 16) & 0xff] << 16) | ((uint32_t)AES_SBOX[s3 >> 24] << 24)) ^ k[0];
    uint32_t o1 = ((uint32_t)AES_SBOX[s1 & 0xff] | ((uint32_t)AES_SBOX[(s2 >> 8) & 0xff] << 8) | ((uint32_t)AES_SBOX[(s3 >> 16) & 0xff] << 16) | ((uint32_t)AES_SBOX[s0 >> 24] << 24)) ^ k[1];
    uint32_t o2 = ((uint32_t)AES_SBOX[s2 & 0xff] | ((uint32_t)AES_SBOX[(s3 >> 8) & 0xff] << 8) | ((uint32_t)AES_SBOX[(s0 >> 16) & 0xff] << 16) | ((uint32_t)AES_SBOX[s1 >> 24] << 24)) ^ k[2];
    uint32_t o3 = ((uint32_t)AES_SBOX[s3 & 0xff] | ((uint32_t)AES_SBOX[(s0 >> 8) & 0xff] << 8) | ((uint32_t)AES_SBOX[(s1 >> 16) & 0xff] << 16) | ((uint32_t)AES_SBOX[s2 >> 24] << 24)) ^ k[3];
    out[0] = (uint64_t)o0 | ((uint64_t)o1 << 32);
    out[1] = (uint64_t)o2 | ((uint64_t)o3 << 32);
}
/* word FIFO: word w of the stream = half (w&1) of block (w>>1)  (crypto/lpn.hpp:108-139) */
static uint64_t aesctr_next(aesctr_t* a) {
    uint64_t x;
    if (a->has_buf) { a->has_buf = 0; x = a->buf[1]; }
    else { aesctr_block(a, a->buf); a->has_buf = 1; x = a->buf[0]; }
    if (a->patched && a->nout == g_patch_word) x |= g_patch_or;
    a->nout++;
    return x;
}
/* crypto/lpn.hpp:141-148 -- strict '<' acceptance */
static uint64_t aesctr_bounded(aesctr_t* a, uint64_t M) {
    if (M <= 1) return 0;
    uint64_t lim = UINT64_MAX - (UINT64_MAX % M);
    for (;;) { uint64_t x = aesctr_next(a); if (x < lim) return x % M; }
}
/* a mixed sequence of draws from ONE stream: moduli[i] == 0 -> next_u64(), else bounded(moduli[i]); pins the rejection branch of
 * bounded() and its interplay with the word FIFO against the reference (moduli just above 2^63 reject every other word) */
void orc_aes_ctr_draws(const uint8_t key[32], uint64_t nonce, const uint64_t* moduli, uint64_t* out, size_t n) {
    aesctr_t a; aesctr_init(&a, key, nonce);
    for (size_t i = 0; i < n; i++) out[i] = moduli[i] ? aesctr_bounded(&a, moduli[i]) : aesctr_next(&a);
}
void orc_aes_ctr_words(const uint8_t key[32], uint64_t nonce, uint64_t* out, size_t n) {
    aesctr_t a; aesctr_init(&a, key, nonce);
    for (size_t i = 0; i < n; i++) out[i] = aesctr_next(&a);
}

/* ------------------------------------------------------------------ keys */
typedef struct {
    uint64_t canon_tag;
    uint8_t h_digest[32];
    uint64_t* H; /* n_bits columns x 128 words */
    fp_t powg[ORC_B];
    uint64_t prf_k[4];
    uint64_t lpn_s[ORC_LPN_WORDS];
    int lpn_rows; /* rows of the LPN sample actually evaluated; ORC_LPN_T = as the reference, >=127 gives identical PRF output */
} orc_keys;

/* crypto/lpn.hpp:157-164 */
uint64_t orc_fnv1a(const char* dom) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (const char* p = dom; *p; ++p) { h ^= (uint8_t)*p; h *= 0x100000001b3ull; }
    return h;
}
/* crypto/lpn.hpp:166-192 */
void orc_derive_aes_key(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom, uint8_t key[32], uint64_t* nonce) {
    sha_t s; sha_init(&s);
    for (int i = 0; i < 4; i++) sha_u64le(&s, k->prf_k[i]);
    sha_u64le(&s, k->canon_tag);
    sha_update(&s, k->h_digest, 32);
    sha_u64le(&s, ztag); sha_u64le(&s, nlo); sha_u64le(&s, nhi);
    uint64_t dh = orc_fnv1a(dom);
    sha_u64le(&s, dh);
    sha_final(&s, key);
    *nonce = dh ^ nlo;
}
static int parity64(uint64_t x) { return __builtin_parityll(x); }
/* crypto/lpn.hpp:194-233 -- rows r < `rows`; ybits has (rows+63)/64 words */
void orc_lpn_make_ybits(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom, int rows, uint64_t* ybits) {
    uint8_t key[32]; uint64_t nonce;
    orc_derive_aes_key(k, ztag, nlo, nhi, dom, key, &nonce);
    aesctr_t prg; aesctr_init(&prg, key, nonce);
    prg.patched = 1;
    memset(ybits, 0, (size_t)((rows + 63) / 64) * 8);
    for (int r = 0; r < rows; r++) {
        uint64_t acc = 0;
        for (int wi = 0; wi < ORC_LPN_WORDS; wi++) acc ^= aesctr_next(&prg) & k->lpn_s[wi];
        int e = aesctr_bounded(&prg, 8) < 1 ? 1 : 0; /* tau = 1/8 */
        ybits[r >> 6] ^= (uint64_t)(parity64(acc) ^ e) << (r & 63);
    }
}
/* crypto/toeplitz.hpp:22-48,121-141 -- bits 0..126 of the GF(2)[x] product ybits*top */
void orc_toep_127(const uint64_t* top, size_t ntop, const uint64_t* y, size_t ny, uint64_t out[2]) {
    out[0] = out[1] = 0;
    for (int j = 0; j < 127; j++) {
        int bit = 0;
        for (int i = 0; i <= j; i++) {
            size_t yi = (size_t)i >> 6, ti = (size_t)(j - i) >> 6;
            int yb = yi < ny ? (int)((y[yi] >> (i & 63)) & 1) : 0;
            int tb = ti < ntop ? (int)((top[ti] >> ((j - i) & 63)) & 1) : 0;
            bit ^= yb & tb;
        }
        out[j >> 6] |= (uint64_t)bit << (j & 63);
    }
}
/* crypto/lpn.hpp:235-261 */
fp_t orc_prf_R_core_fp(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom) {
    int rows = k->lpn_rows;
    size_t ny = (size_t)(rows + 63) / 64;
    uint64_t* y = (uint64_t*)malloc(ny * 8);
    orc_lpn_make_ybits(k, ztag, nlo, nhi, dom, rows, y);
    uint8_t key[32]; uint64_t nonce;
    orc_derive_aes_key(k, ztag, nlo, nhi, "pvac.dom.toeplitz", key, &nonce);
    nonce ^= orc_fnv1a(dom);
    aesctr_t prg; aesctr_init(&prg, key, nonce);
    size_t ntop = ((size_t)rows + 127 + 63) / 64;
    uint64_t* top = (uint64_t*)malloc(ntop * 8);
    for (size_t i = 0; i < ntop; i++) top[i] = aesctr_next(&prg);
    uint64_t o[2];
    orc_toep_127(top, ntop, y, ny, o);
    free(y); free(top);
    return hash_to_fp_nonzero(o[0], o[1]);
}
/* crypto/lpn.hpp:263-275 */
static fp_t prf_triple(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* d1, const char* d2, const char* d3) {
    fp_t r1 = orc_prf_R_core_fp(k, ztag, nlo, nhi, d1);
    fp_t r2 = orc_prf_R_core_fp(k, ztag, nlo, nhi, d2);
    fp_t r3 = orc_prf_R_core_fp(k, ztag, nlo, nhi, d3);
    return fp_mul(fp_mul(r1, r2), r3);
}
static fp_t prf_R(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi) { return prf_triple(k, ztag, nlo, nhi, "pvac.prf.r.1", "pvac.prf.r.2", "pvac.prf.r.3"); }
static fp_t prf_R_noise(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi) { return prf_triple(k, ztag, nlo, nhi, "pvac.prf.noise.1", "pvac.prf.noise.2", "pvac.prf.noise.3"); }
/* ops/encrypt.hpp:114-129 */
static fp_t prf_noise_delta(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint32_t group_id, uint8_t kind) {
    uint64_t g = (uint64_t)group_id + 1, kk = (uint64_t)kind + 1;
    nlo ^= 0x9e3779b97f4a7c15ull * g; nhi ^= 0x94d049bb133111ebull * g; ztag ^= 0x517cc1b727220a95ull * g;
    nlo ^= kk; nhi ^= kk << 32; ztag ^= kk << 48;
    return prf_R_noise(k, ztag, nlo, nhi);
}
/* crypto/matrix.hpp:254-264 */
uint64_t orc_prg_layer_ztag(uint64_t canon_tag, uint64_t nlo, uint64_t nhi) {
    sha_t s; sha_init(&s);
    sha_update(&s, "pvac.dom.ztag", 13);
    sha_u64le(&s, canon_tag); sha_u64le(&s, nlo); sha_u64le(&s, nhi);
    uint8_t out[32]; sha_final(&s, out);
    return le64(out);
}
/* crypto/matrix.hpp:15-92 -- k distinct values of [0,N) in draw order; SHA-256(label|words|LE64(ctr)) gives 4 words */
void orc_prg_choose_k(int k, int N, const char* label, const uint64_t* words, size_t nwords, int32_t* out) {
    uint8_t* used = (uint8_t*)calloc((size_t)N, 1);
    uint64_t ctr = 0; uint8_t buf[32]; int pos = 32, n = 0;
    uint64_t lim = UINT64_MAX - (UINT64_MAX % (uint64_t)N);
    while (n < k) {
        if (pos >= 32) {
            sha_t s; sha_init(&s);
            sha_update(&s, label, strlen(label));
            for (size_t i = 0; i < nwords; i++) sha_u64le(&s, words[i]);
            sha_u64le(&s, ctr++);
            sha_final(&s, buf);
            pos = 0;
        }
        uint64_t x = le64(buf + pos); pos += 8;
        if (N > 1 && x > lim) continue; /* non-strict '<=' acceptance, matrix.hpp:71 */
        int v = N > 1 ? (int)(x % (uint64_t)N) : 0;
        if (!used[v]) { used[v] = 1; out[n++] = v; }
    }
    free(used);
}
/* crypto/matrix.hpp:267-303 */
void orc_sigma_from_H(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint16_t idx, uint8_t ch, uint64_t salt, uint64_t* s) {
    uint64_t words[7] = { k->canon_tag, ztag, nlo, nhi, idx, ch, salt };
    int32_t pick[ORC_X_COL_WT > ORC_ERR_WT ? ORC_X_COL_WT : ORC_ERR_WT];
    memset(s, 0, ORC_M_WORDS * 8);
    orc_prg_choose_k(ORC_X_COL_WT, ORC_N_BITS, "pvac.dom.x_seed", words, 7, pick);
    for (int i = 0; i < ORC_X_COL_WT; i++) {
        const uint64_t* col = k->H + (size_t)pick[i] * ORC_M_WORDS;
        for (int w = 0; w < ORC_M_WORDS; w++) s[w] ^= col[w];
    }
    orc_prg_choose_k(ORC_ERR_WT, ORC_M_BITS, "pvac.dom.noise", words, 7, pick);
    for (int i = 0; i < ORC_ERR_WT; i++) s[pick[i] >> 6] ^= 1ull << (pick[i] & 63);
}
/* ops/encrypt.hpp:16-27 with the default Params (core/types.hpp:36-70) */
void orc_plan_noise(int depth_hint, int* z2o, int* z3o) {
    double budget = 120.0 + 16.0 * (depth_hint > 0 ? depth_hint : 0);
    double per2 = 2.0 * log2((double)ORC_B), per3 = 3.0 * log2((double)ORC_B);
    int z2 = (int)floor((budget * 0.55) / per2), z3 = (int)floor((budget * (1.0 - 0.55)) / per3);
    if (z2 < 0) z2 = 0;
    if (z3 < 0) z3 = 0;
    if (z2 + z3 == 1) { if (z3 > 0) ++z3; else ++z2; }
    *z2o = z2; *z3o = z3;
}

/* crypto/keygen.hpp:35-136 + crypto/matrix.hpp:191-251 (gen_H). gen_ubk_public draws nothing from the tape and is
 * not on the hot path, so it is skipped. rand_fp at keygen.hpp:59 passes two csprng_u64() calls as function
 * arguments; g++ 13.3 evaluates the second (hi) first -- pinned by tests/test_oracle_vs_ref.py. */
orc_keys* orc_keygen(uint64_t tape_state) {
    orc_keys* k = (orc_keys*)calloc(1, sizeof(orc_keys));
    tape_t t = { tape_state, 0 };
    k->lpn_rows = ORC_LPN_T;
    k->canon_tag = tape_u64(&t);
    k->H = (uint64_t*)calloc((size_t)ORC_N_BITS * ORC_M_WORDS, 8);
    sha_t hs; sha_init(&hs);
    sha_update(&hs, "H|v2", 4);
    sha_u64le(&hs, ORC_M_BITS); sha_u64le(&hs, ORC_N_BITS); sha_u64le(&hs, ORC_H_COL_WT);
    for (int c = 0; c < ORC_N_BITS; c++) {
        uint64_t words[5] = { ORC_M_BITS, ORC_N_BITS, ORC_H_COL_WT, (uint64_t)c, k->canon_tag };
        int32_t rows[ORC_H_COL_WT];
        orc_prg_choose_k(ORC_H_COL_WT, ORC_M_BITS, "pvac.dom.h_gen", words, 5, rows);
        uint64_t* col = k->H + (size_t)c * ORC_M_WORDS;
        for (int i = 0; i < ORC_H_COL_WT; i++) col[rows[i] >> 6] |= 1ull << (rows[i] & 63);
    }
    for (size_t i = 0; i < (size_t)ORC_N_BITS * ORC_M_WORDS; i++) sha_u64le(&hs, k->H[i]);
    sha_final(&hs, k->h_digest);
    for (int i = 0; i < 4; i++) k->prf_k[i] = tape_u64(&t);
    u128 E = (P127 - 1) / ORC_B;
    fp_t g;
    for (;;) {
        uint64_t hi = tape_u64(&t) & MASK63; uint64_t lo = tape_u64(&t);
        fp_t h = fp_from_words(lo, hi);
        if (fp_is_zero(h)) continue;
        fp_t acc = fp_pow_u128(h, E);
        if (!fp_eq(acc, fp_from_u64(1))) { g = acc; break; }
    }
    k->powg[0] = fp_from_u64(1);
    for (int i = 1; i < ORC_B; i++) k->powg[i] = fp_mul(k->powg[i - 1], g);
    /* omega_B search (keygen.hpp:99-122): value unused anywhere, but it consumes tape words. 337 is prime, so the
     * loop accepts the first h whose (truncated-exponent) power is not 1. */
    for (;;) {
        uint64_t hi = tape_u64(&t) & MASK63; uint64_t lo = tape_u64(&t);
        fp_t h = fp_from_words(lo, hi);
        if (fp_is_zero(h)) continue;
        fp_t w = fp_pow_u128(h, (u128)(uint64_t)E);
        if (fp_eq(w, fp_from_u64(1))) continue;
        break;
    }
    for (int i = 0; i < ORC_LPN_WORDS; i++) k->lpn_s[i] = tape_u64(&t);
    return k;
}
orc_keys* orc_keys_from_raw(uint64_t canon_tag, const uint8_t* h_digest, const uint64_t* H, const uint64_t* powg, const uint64_t* prf_k, const uint64_t* lpn_s) {
    orc_keys* k = (orc_keys*)calloc(1, sizeof(orc_keys));
    k->lpn_rows = ORC_LPN_T;
    k->canon_tag = canon_tag;
    memcpy(k->h_digest, h_digest, 32);
    k->H = (uint64_t*)calloc((size_t)ORC_N_BITS * ORC_M_WORDS, 8);
    if (H) memcpy(k->H, H, (size_t)ORC_N_BITS * ORC_M_WORDS * 8);
    for (int i = 0; i < ORC_B; i++) { if (powg) { k->powg[i].lo = powg[2 * i]; k->powg[i].hi = powg[2 * i + 1]; } else k->powg[i] = fp_from_u64(1); }
    memcpy(k->prf_k, prf_k, 32);
    memcpy(k->lpn_s, lpn_s, ORC_LPN_WORDS * 8);
    return k;
}
void orc_keys_free(orc_keys* k) { if (k) { free(k->H); free(k); } }
void orc_keys_set_lpn_rows(orc_keys* k, int rows) { k->lpn_rows = rows; }
void orc_keys_export(const orc_keys* k, uint64_t* canon_tag, uint8_t* h_digest, uint64_t* H, uint64_t* powg, uint64_t* prf_k, uint64_t* lpn_s) {
    *canon_tag = k->canon_tag;
    memcpy(h_digest, k->h_digest, 32);
    if (H) memcpy(H, k->H, (size_t)ORC_N_BITS * ORC_M_WORDS * 8);
    for (int i = 0; i < ORC_B; i++) { powg[2 * i] = k->powg[i].lo; powg[2 * i + 1] = k->powg[i].hi; }
    memcpy(prf_k, k->prf_k, 32);
    memcpy(lpn_s, k->lpn_s, ORC_LPN_WORDS * 8);
}

/* ------------------------------------------------------------------ ciphertexts (core/types.hpp:96-119) */
typedef struct { uint8_t rule; uint64_t ztag, nlo, nhi; uint32_t pa, pb; } layer_t;
typedef struct { uint32_t lid; uint16_t idx; uint8_t ch; fp_t w; uint64_t s[ORC_M_WORDS]; } edge_t;
typedef struct { uint32_t nL, nE, capL, capE; layer_t* L; edge_t* E; } orc_ct;

static orc_ct* ct_new(void) { return (orc_ct*)calloc(1, sizeof(orc_ct)); }
void orc_ct_free(orc_ct* c) { if (c) { free(c->L); free(c->E); free(c); } }
static void ct_push_layer(orc_ct* c, layer_t l) {
    if (c->nL == c->capL) { c->capL = c->capL ? c->capL * 2 : 8; c->L = (layer_t*)realloc(c->L, (size_t)c->capL * sizeof(layer_t)); }
    c->L[c->nL++] = l;
}
static edge_t* ct_push_edge(orc_ct* c) {
    if (c->nE == c->capE) { c->capE = c->capE ? c->capE * 2 : 64; c->E = (edge_t*)realloc(c->E, (size_t)c->capE * sizeof(edge_t)); }
    return &c->E[c->nE++];
}

/* ops/encrypt.hpp:73-104 */
static void compact_layers(orc_ct* c) {
    uint32_t L = c->nL;
    if (!L) return;
    uint8_t* used = (uint8_t*)calloc(L, 1);
    for (uint32_t i = 0; i < c->nE; i++) if (c->E[i].lid < L) used[c->E[i].lid] = 1;
    for (int changed = 1; changed;) {
        changed = 0;
        for (uint32_t l = 0; l < L; l++) {
            if (!used[l] || c->L[l].rule != 1) continue;
            uint32_t pa = c->L[l].pa, pb = c->L[l].pb;
            if (pa < L && !used[pa]) { used[pa] = 1; changed = 1; }
            if (pb < L && !used[pb]) { used[pb] = 1; changed = 1; }
        }
    }
    uint32_t* remap = (uint32_t*)malloc((size_t)L * 4);
    uint32_t n = 0;
    for (uint32_t l = 0; l < L; l++) { remap[l] = used[l] ? n : UINT32_MAX; if (used[l]) c->L[n++] = c->L[l]; }
    if (n != L) {
        c->nL = n;
        for (uint32_t l = 0; l < n; l++) if (c->L[l].rule == 1) { c->L[l].pa = remap[c->L[l].pa]; c->L[l].pb = remap[c->L[l].pb]; }
        for (uint32_t i = 0; i < c->nE; i++) c->E[i].lid = remap[c->E[i].lid];
    }
    free(used); free(remap);
}

/* ops/encrypt.hpp:39-71 -- merge equal (layer, idx, sign); drop w==0 && sigma==0; order (lid, idx, P, M) */
static void compact_edges(orc_ct* c) {
    size_t slots = (size_t)c->nL * ORC_B * 2;
    int32_t* first = (int32_t*)malloc(slots * 4);
    for (size_t i = 0; i < slots; i++) first[i] = -1;
    edge_t* out = (edge_t*)malloc((size_t)(c->nE ? c->nE : 1) * sizeof(edge_t));
    edge_t* agg = (edge_t*)malloc((size_t)(c->nE ? c->nE : 1) * sizeof(edge_t));
    uint32_t na = 0;
    for (uint32_t i = 0; i < c->nE; i++) {
        edge_t* e = &c->E[i];
        size_t slot = ((size_t)e->lid * ORC_B + e->idx) * 2 + (e->ch ? 1 : 0);
        if (first[slot] < 0) {
            first[slot] = (int32_t)na;
            agg[na] = *e;
            agg[na].w = fp_add(fp_from_u64(0), e->w);
            na++;
        } else {
            edge_t* a = &agg[first[slot]];
            a->w = fp_add(a->w, e->w);
            for (int w = 0; w < ORC_M_WORDS; w++) a->s[w] ^= e->s[w];
        }
    }
    uint32_t n = 0;
    for (uint32_t l = 0; l < c->nL; l++)
        for (int k = 0; k < ORC_B; k++)
            for (int sg = 0; sg < 2; sg++) {
                int32_t ai = first[((size_t)l * ORC_B + k) * 2 + sg];
                if (ai < 0) continue;
                edge_t* a = &agg[ai];
                int nz = !fp_is_zero(a->w);
                for (int w = 0; w < ORC_M_WORDS && !nz; w++) nz = a->s[w] != 0;
                if (!nz) continue;
                out[n] = *a; out[n].lid = l; out[n].idx = (uint16_t)k; out[n].ch = (uint8_t)sg; n++;
            }
    memcpy(c->E, out, (size_t)n * sizeof(edge_t));
    c->nE = n;
    free(first); free(out); free(agg);
}

static void make_edge(orc_ct* c, const orc_keys* k, const layer_t* L, uint16_t idx, uint8_t ch, fp_t w, tape_t* t) {
    /* ops/encrypt.hpp:150-153 -- the salt is the only draw */
    edge_t* e = ct_push_edge(c);
    e->lid = 0; e->idx = idx; e->ch = ch; e->w = w;
    orc_sigma_from_H(k, L->ztag, L->nlo, L->nhi, idx, ch, tape_u64(t), e->s);
}

/* ops/encrypt.hpp:162-258 */
static orc_ct* enc_fp_depth(const orc_keys* k, tape_t* t, fp_t v, int depth_hint) {
    orc_ct* C = ct_new();
    layer_t L; memset(&L, 0, sizeof L);
    L.rule = 0;
    L.nlo = tape_u64(t); L.nhi = tape_u64(t);
    L.ztag = orc_prg_layer_ztag(k->canon_tag, L.nlo, L.nhi);
    ct_push_layer(C, L);
    enum { S = 8 };
    int idx[S]; uint8_t ch[S]; fp_t r[S];
    for (int j = 0; j < S; j++) {
        for (;;) { /* pick_unique_idx, encrypt.hpp:131-136 */
            int x = (int)(tape_u64(t) % ORC_B), dup = 0;
            for (int q = 0; q < j; q++) dup |= idx[q] == x;
            if (!dup) { idx[j] = x; break; }
        }
        ch[j] = (uint8_t)(tape_u64(t) & 1);
    }
    fp_t sumg = fp_from_u64(0);
    for (int j = 0; j < S - 1; j++) {
        r[j] = rand_fp_nonzero(t);
        fp_t term = fp_mul(r[j], k->powg[idx[j]]);
        sumg = ch[j] == 0 ? fp_add(sumg, term) : fp_sub(sumg, term);
    }
    fp_t r_last = fp_mul(fp_sub(v, sumg), fp_inv(k->powg[idx[S - 1]]));
    r[S - 1] = ch[S - 1] ? fp_neg(r_last) : r_last;
    fp_t R = prf_R(k, L.ztag, L.nlo, L.nhi);
    for (int j = 0; j < S; j++) make_edge(C, k, &L, (uint16_t)idx[j], ch[j], fp_mul(r[j], R), t);

    int Z2, Z3; orc_plan_noise(depth_hint, &Z2, &Z3);
    int total = Z2 + Z3, gid = 0;
    fp_t delta_acc = fp_from_u64(0);
    for (int q = 0; q < Z2; q++, gid++) {
        int i = (int)(tape_u64(t) % ORC_B), j;
        do { j = (int)(tape_u64(t) % ORC_B); } while (j == i);
        uint8_t s1 = (uint8_t)(tape_u64(t) & 1), s2 = s1 ^ 1;
        fp_t Delta;
        if (total - gid <= 1) Delta = fp_neg(delta_acc);
        else { Delta = prf_noise_delta(k, L.ztag, L.nlo, L.nhi, (uint32_t)gid, 0); delta_acc = fp_add(delta_acc, Delta); }
        fp_t Dp = s1 == 0 ? Delta : fp_neg(Delta);
        fp_t ri = rand_fp_nonzero(t);
        fp_t rj = fp_mul(fp_sub(fp_mul(ri, k->powg[i]), Dp), fp_inv(k->powg[j]));
        make_edge(C, k, &L, (uint16_t)i, s1, fp_mul(ri, R), t);
        make_edge(C, k, &L, (uint16_t)j, s2, fp_mul(rj, R), t);
    }
    for (int q = 0; q < Z3; q++, gid++) {
        int i = (int)(tape_u64(t) % ORC_B), j, kk;
        do { j = (int)(tape_u64(t) % ORC_B); } while (j == i);
        do { kk = (int)(tape_u64(t) % ORC_B); } while (kk == i || kk == j);
        uint8_t s1 = (uint8_t)(tape_u64(t) & 1), s2 = (uint8_t)(tape_u64(t) & 1), s3 = (uint8_t)(tape_u64(t) & 1);
        fp_t Delta;
        if (total - gid <= 1) Delta = fp_neg(delta_acc);
        else { Delta = prf_noise_delta(k, L.ztag, L.nlo, L.nhi, (uint32_t)gid, 1); delta_acc = fp_add(delta_acc, Delta); }
        fp_t a = rand_fp_nonzero(t), b = rand_fp_nonzero(t);
        fp_t t1 = fp_mul(a, k->powg[i]), t2 = fp_mul(b, k->powg[j]);
        if (s1) t1 = fp_neg(t1);
        if (s2) t2 = fp_neg(t2);
        fp_t gk = s3 == 0 ? k->powg[kk] : fp_neg(k->powg[kk]);
        fp_t cc = fp_mul(fp_sub(Delta, fp_add(t1, t2)), fp_inv(gk));
        make_edge(C, k, &L, (uint16_t)i, s1, fp_mul(a, R), t);
        make_edge(C, k, &L, (uint16_t)j, s2, fp_mul(b, R), t);
        make_edge(C, k, &L, (uint16_t)kk, s3, fp_mul(cc, R), t);
    }
    compact_edges(C);
    /* guard_budget (encrypt.hpp:106-111) can never fire here: <= 8+2*Z2+3*Z3 edges */
    for (uint32_t i = C->nE; i-- > 1;) { /* shuffle_edges, encrypt.hpp:155-160 */
        uint32_t j = (uint32_t)(tape_u64(t) % (i + 1));
        edge_t tmp = C->E[i]; C->E[i] = C->E[j]; C->E[j] = tmp;
    }
    return C;
}

/* ops/encrypt.hpp:260-279 and ops/arithmetic.hpp:12-31 (identical bodies) */
static orc_ct* concat_ct(const orc_ct* a, const orc_ct* b) {
    orc_ct* C = ct_new();
    for (uint32_t i = 0; i < a->nL; i++) ct_push_layer(C, a->L[i]);
    uint32_t off = a->nL;
    for (uint32_t i = 0; i < b->nL; i++) { layer_t l = b->L[i]; if (l.rule == 1) { l.pa += off; l.pb += off; } ct_push_layer(C, l); }
    for (uint32_t i = 0; i < a->nE; i++) *ct_push_edge(C) = a->E[i];
    for (uint32_t i = 0; i < b->nE; i++) { edge_t* e = ct_push_edge(C); *e = b->E[i]; e->lid += off; }
    if (C->nE > 1200000u) compact_edges(C); /* guard_budget */
    compact_layers(C);
    return C;
}

orc_ct* orc_enc_fp_depth(const orc_keys* k, uint64_t tape_state, const uint64_t* v, int depth, uint64_t* draws) {
    tape_t t = { tape_state, 0 };
    fp_t x = { v[0], v[1] };
    orc_ct* c = enc_fp_depth(k, &t, x, depth);
    if (draws) *draws = t.draws;
    return c;
}
/* ops/encrypt.hpp:281-291. enc_value_depth passes two enc_fp_depth calls as arguments of combine_ciphers; with
 * g++ 13.3 the second (-mask) is evaluated first (pinned by tests/test_oracle_vs_ref.py). */
orc_ct* orc_enc_value_depth(const orc_keys* k, uint64_t tape_state, uint64_t v, int depth_hint, uint64_t* draws) {
    tape_t t = { tape_state, 0 };
    fp_t val = fp_from_u64(v);
    fp_t mask = rand_fp_nonzero(&t);
    orc_ct* b = enc_fp_depth(k, &t, fp_neg(mask), depth_hint);
    orc_ct* a = enc_fp_depth(k, &t, fp_add(val, mask), depth_hint);
    orc_ct* c = concat_ct(a, b);
    orc_ct_free(a); orc_ct_free(b);
    if (draws) *draws = t.draws;
    return c;
}
orc_ct* orc_enc_value(const orc_keys* k, uint64_t tape_state, uint64_t v, uint64_t* draws) { return orc_enc_value_depth(k, tape_state, v, 0, draws); }
/* ops/encrypt.hpp:293-298: combine(enc_fp_depth(mask), enc_fp_depth(-mask)), second argument first like enc_value_depth */
orc_ct* orc_enc_zero_depth(const orc_keys* k, uint64_t tape_state, int depth_hint, uint64_t* draws) {
    tape_t t = { tape_state, 0 };
    fp_t mask = rand_fp_nonzero(&t);
    orc_ct* b = enc_fp_depth(k, &t, fp_neg(mask), depth_hint);
    orc_ct* a = enc_fp_depth(k, &t, mask, depth_hint);
    orc_ct* c = concat_ct(a, b);
    orc_ct_free(a); orc_ct_free(b);
    if (draws) *draws = t.draws;
    return c;
}
/* utils/text.hpp:15-61 -- enc_value(length), then enc_fp_depth(pack_15_bytes(block j), 2 + j), all from one tape. Returns the
 * number of ciphertexts written to out (1 + ceil(len / 15)), or -1 if cap is too small. */
int orc_enc_text(const orc_keys* k, uint64_t tape_state, const uint8_t* msg, uint64_t len, orc_ct** out, int cap, uint64_t* draws) {
    int need = 1 + (int)((len + 14) / 15);
    if (cap < need) return -1;
    tape_t t = { tape_state, 0 };
    {
        fp_t val = fp_from_u64(len);
        fp_t mask = rand_fp_nonzero(&t);
        orc_ct* b = enc_fp_depth(k, &t, fp_neg(mask), 0);
        orc_ct* a = enc_fp_depth(k, &t, fp_add(val, mask), 0);
        out[0] = concat_ct(a, b);
        orc_ct_free(a); orc_ct_free(b);
    }
    int depth = 2;
    for (uint64_t pos = 0; pos < len; pos += 15, depth++) {
        uint64_t take = len - pos < 15 ? len - pos : 15, lo = 0, hi = 0;
        for (uint64_t i = 0; i < take; i++) {
            if (i < 8) lo |= (uint64_t)msg[pos + i] << (8 * i);
            else hi |= (uint64_t)msg[pos + i] << (8 * (i - 8));
        }
        fp_t x = fp_from_words(lo, hi);
        out[1 + pos / 15] = enc_fp_depth(k, &t, x, depth);
    }
    if (draws) *draws = t.draws;
    return need;
}
orc_ct* orc_ct_add(const orc_ct* a, const orc_ct* b) { return concat_ct(a, b); }
/* ops/arithmetic.hpp:33-37 */
orc_ct* orc_ct_scale(const orc_ct* a, const uint64_t* s) {
    orc_ct* c = ct_new(); /* plain copy: ct_scale does not run compact_layers */
    for (uint32_t i = 0; i < a->nL; i++) ct_push_layer(c, a->L[i]);
    fp_t sc = { s[0], s[1] };
    for (uint32_t i = 0; i < a->nE; i++) { edge_t* x = ct_push_edge(c); *x = a->E[i]; x->w = fp_mul(x->w, sc); }
    return c;
}
/* ops/encrypt.hpp:39-71 on a copy (the reference mutates in place) */
orc_ct* orc_compact_edges(const orc_ct* a) {
    orc_ct* c = ct_new();
    for (uint32_t i = 0; i < a->nL; i++) ct_push_layer(c, a->L[i]);
    for (uint32_t i = 0; i < a->nE; i++) *ct_push_edge(c) = a->E[i];
    compact_edges(c);
    return c;
}
/* ops/arithmetic.hpp:39-45 */
orc_ct* orc_ct_sub(const orc_ct* a, const orc_ct* b) {
    fp_t m1 = fp_neg(fp_from_u64(1));
    uint64_t s[2] = { m1.lo, m1.hi };
    orc_ct* nb = orc_ct_scale(b, s);
    orc_ct* c = concat_ct(a, nb);
    orc_ct_free(nb);
    return c;
}

/* ops/arithmetic.hpp:47-106. Emission order = libstdc++ unordered_map iteration order, restated: nodes live in one
 * singly linked list; a key landing in an empty bucket goes to the list head, a key landing in an occupied bucket goes
 * to the front of that bucket's run. So the final order is: buckets by first-occupation time DESCENDING, and inside a
 * bucket keys by insertion time DESCENDING. bucket = (k * 0x9E3779B97F4A7C15 mod 2^64) % nb with
 * nb = bucket_count after reserve(|A.E|*|B.E|) (no rehash can happen afterwards). */
typedef struct { uint64_t key; fp_t wp, wm; int ip, im; uint64_t t_ins, t_bkt; } agg_t;
static int agg_cmp(const void* x, const void* y) {
    const agg_t* a = (const agg_t*)x; const agg_t* b = (const agg_t*)y;
    if (a->t_bkt != b->t_bkt) return a->t_bkt > b->t_bkt ? -1 : 1;
    if (a->t_ins != b->t_ins) return a->t_ins > b->t_ins ? -1 : 1;
    return 0;
}
orc_ct* orc_ct_mul(const orc_keys* k, uint64_t tape_state, const orc_ct* A, const orc_ct* B, uint64_t* draws) {
    tape_t t = { tape_state, 0 };
    orc_ct* C = ct_new();
    for (uint32_t i = 0; i < A->nL; i++) ct_push_layer(C, A->L[i]);
    uint32_t off = A->nL, LA = A->nL, LB = B->nL;
    for (uint32_t i = 0; i < B->nL; i++) { layer_t l = B->L[i]; if (l.rule == 1) { l.pa += off; l.pb += off; } ct_push_layer(C, l); }
    uint32_t base = C->nL;
    for (uint32_t la = 0; la < LA; la++)
        for (uint32_t lb = 0; lb < LB; lb++) {
            layer_t l; memset(&l, 0, sizeof l);
            l.rule = 1; l.pa = la; l.pb = off + lb;
            l.nlo = tape_u64(&t); l.nhi = tape_u64(&t);
            l.ztag = orc_prg_layer_ztag(k->canon_tag, l.nlo, l.nhi);
            ct_push_layer(C, l);
        }
    uint64_t npairs = (uint64_t)A->nE * B->nE;
    if (npairs) {
        uint64_t nb = orc_next_bkt(npairs);
        /* dense index over the possible keys: (la*LB+lb, (idxa+idxb)%B) */
        size_t nkeys = (size_t)LA * LB * ORC_B;
        int64_t* slot = (int64_t*)malloc(nkeys * 8);
        for (size_t i = 0; i < nkeys; i++) slot[i] = -1;
        agg_t* agg = (agg_t*)calloc(nkeys < npairs ? nkeys : (size_t)npairs, sizeof(agg_t));
        size_t na = 0;
        for (uint32_t ia = 0; ia < A->nE; ia++)
            for (uint32_t ib = 0; ib < B->nE; ib++) {
                const edge_t* ea = &A->E[ia]; const edge_t* eb = &B->E[ib];
                uint32_t lp = ea->lid * LB + eb->lid;
                uint32_t sidx = (uint32_t)((ea->idx + eb->idx) % ORC_B);
                size_t di = (size_t)lp * ORC_B + sidx;
                if (slot[di] < 0) {
                    slot[di] = (int64_t)na;
                    agg[na].key = ((uint64_t)lp << 32) | sidx;
                    agg[na].t_ins = na; /* insertion rank */
                    agg[na].wp = agg[na].wm = fp_from_u64(0);
                    na++;
                }
                agg_t* a = &agg[slot[di]];
                fp_t ww = fp_mul(ea->w, eb->w);
                if (ea->ch == eb->ch) { a->ip = 1; a->wp = fp_add(a->wp, ww); }
                else { a->im = 1; a->wm = fp_add(a->wm, ww); }
            }
        /* bucket first-occupation times */
        {
            /* sort-free: keys are few (<= nkeys); use a small open-addressing map bucket -> first time */
            size_t cap = 1; while (cap < na * 2 + 1) cap <<= 1;
            uint64_t* bk = (uint64_t*)malloc(cap * 8); uint64_t* bt = (uint64_t*)malloc(cap * 8);
            for (size_t i = 0; i < cap; i++) bk[i] = UINT64_MAX;
            for (size_t i = 0; i < na; i++) { /* agg[] is already in insertion order */
                uint64_t b = (agg[i].key * 0x9E3779B97F4A7C15ull) % nb;
                size_t h = (size_t)(mix64(b) & (cap - 1));
                while (bk[h] != UINT64_MAX && bk[h] != b) h = (h + 1) & (cap - 1);
                if (bk[h] == UINT64_MAX) { bk[h] = b; bt[h] = agg[i].t_ins; }
                agg[i].t_bkt = bt[h];
            }
            free(bk); free(bt);
        }
        qsort(agg, na, sizeof(agg_t), agg_cmp);
        for (size_t i = 0; i < na; i++) {
            uint32_t lid = base + (uint32_t)(agg[i].key >> 32);
            uint16_t idx = (uint16_t)(agg[i].key & 0xFFFF);
            const layer_t* Lp = &C->L[lid];
            for (int sg = 0; sg < 2; sg++) {
                int have = sg ? agg[i].im : agg[i].ip;
                fp_t w = sg ? agg[i].wm : agg[i].wp;
                if (!have || fp_is_zero(w)) continue;
                edge_t* e = ct_push_edge(C);
                e->lid = lid; e->idx = idx; e->ch = (uint8_t)sg; e->w = w;
                orc_sigma_from_H(k, Lp->ztag, Lp->nlo, Lp->nhi, idx, (uint8_t)sg, tape_u64(&t), e->s);
            }
        }
        free(slot); free(agg);
    }
    if (C->nE > 1200000u) compact_edges(C);
    compact_layers(C);
    if (draws) *draws = t.draws;
    return C;
}

/* ops/commit.hpp:12-87 -- SHA-256("pvac.dom.commit" || H_digest || canon_tag || layers || edges (lid, idx as u64, ch, w 16 B, sigma)) */
void orc_commit_ct(const orc_keys* k, const orc_ct* C, uint8_t out[32]) {
    sha_t s; sha_init(&s);
    sha_update(&s, "pvac.dom.commit", 15);
    sha_update(&s, k->h_digest, 32);
    sha_u64le(&s, k->canon_tag);
    for (uint32_t i = 0; i < C->nL; i++) {
        const layer_t* L = &C->L[i];
        uint8_t r = (uint8_t)L->rule;
        sha_update(&s, &r, 1);
        if (L->rule == 0) { sha_u64le(&s, L->ztag); sha_u64le(&s, L->nlo); sha_u64le(&s, L->nhi); }
        else { sha_u64le(&s, L->pa); sha_u64le(&s, L->pb); }
    }
    for (uint32_t i = 0; i < C->nE; i++) {
        const edge_t* e = &C->E[i];
        sha_u64le(&s, e->lid);
        sha_u64le(&s, e->idx);
        uint8_t ch = e->ch;
        sha_update(&s, &ch, 1);
        sha_u64le(&s, e->w.lo);
        sha_u64le(&s, e->w.hi & MASK63);
        for (int w = 0; w < ORC_M_WORDS; w++) sha_u64le(&s, e->s[w]);
    }
    sha_final(&s, out);
}

/* ------------------------------------------------------------------ UBK, sigma_density, ct_recrypt */
/* crypto/matrix.hpp:95-164 -- Fisher-Yates over 0..m-1 driven by SHA-256("UBK" || LE64(tag) || LE64(ctr)), 4 words per hash */
void orc_ubk_perm(uint64_t canon_tag, int32_t* perm /* 8192 */) {
    for (int i = 0; i < ORC_M_BITS; i++) perm[i] = i;
    uint64_t ctr = 0;
    uint8_t buf[32];
    int idx = 32;
    for (int i = ORC_M_BITS - 1; i > 0; --i) {
        uint64_t M = (uint64_t)i + 1, lim = UINT64_MAX - (UINT64_MAX % M), x;
        for (;;) {
            if (idx >= 32) {
                sha_t s; sha_init(&s);
                sha_update(&s, "UBK", 3);
                sha_u64le(&s, canon_tag);
                sha_u64le(&s, ctr++);
                sha_final(&s, buf);
                idx = 0;
            }
            x = le64(buf + idx);
            idx += 8;
            if (x <= lim) break;
        }
        int j = (int)(x % M);
        int32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
}
/* crypto/matrix.hpp:167-188,306-310 -- every sigma: out bit inv[src] set for every set bit src */
static void ubk_apply(const orc_keys* k, orc_ct* c) {
    static int32_t perm[ORC_M_BITS], inv[ORC_M_BITS];
    static uint64_t tag_of = 0; static int have = 0;
    if (!have || tag_of != k->canon_tag) {
        orc_ubk_perm(k->canon_tag, perm);
        for (int i = 0; i < ORC_M_BITS; i++) inv[perm[i]] = i;
        tag_of = k->canon_tag; have = 1;
    }
    for (uint32_t e = 0; e < c->nE; e++) {
        uint64_t o[ORC_M_WORDS];
        memset(o, 0, sizeof o);
        for (int src = 0; src < ORC_M_BITS; src++)
            if ((c->E[e].s[src >> 6] >> (src & 63)) & 1) { int j = inv[src]; o[j >> 6] |= 1ull << (j & 63); }
        memcpy(c->E[e].s, o, sizeof o);
    }
}
orc_ct* orc_ubk_apply(const orc_keys* k, const orc_ct* a) {
    orc_ct* c = ct_new();
    for (uint32_t i = 0; i < a->nL; i++) ct_push_layer(c, a->L[i]);
    for (uint32_t i = 0; i < a->nE; i++) *ct_push_edge(c) = a->E[i];
    ubk_apply(k, c);
    return c;
}
/* ops/encrypt.hpp:29-37 */
double orc_sigma_density(const orc_ct* c) {
    if (c->nE == 0) return 0.0;
    long double ones = 0, total = 0;
    for (uint32_t e = 0; e < c->nE; e++) {
        uint64_t p = 0;
        for (int w = 0; w < ORC_M_WORDS; w++) p += (uint64_t)__builtin_popcountll(c->E[e].s[w]);
        ones += p;
        total += ORC_M_BITS;
    }
    return (double)(ones / total);
}
/* ops/recrypt.hpp:26-41 */
orc_ct* orc_ct_recrypt(const orc_keys* k, uint64_t tape_state, const orc_ct* in, orc_ct* const* pool, int npool, uint64_t* draws) {
    tape_t t = { tape_state, 0 };
    orc_ct* r = ct_new();
    for (uint32_t i = 0; i < in->nL; i++) ct_push_layer(r, in->L[i]);
    for (uint32_t i = 0; i < in->nE; i++) *ct_push_edge(r) = in->E[i];
    if (npool == 0 || in->nE == 0) { if (draws) *draws = 0; return r; }
    for (int it = 0; it < 8; it++) {
        double d = orc_sigma_density(r);
        if (!(d < 0.495 || d > 0.505)) break;
        uint64_t idx = tape_u64(&t) % (uint64_t)npool;
        orc_ct* s = concat_ct(r, pool[idx]);     /* ct_add: includes guard_budget + compact_layers */
        orc_ct_free(r);
        r = s;
        ubk_apply(k, r);
        if (r->nE > 1200000u) compact_edges(r);
    }
    compact_edges(r);
    compact_layers(r);
    if (draws) *draws = t.draws;
    return r;
}

/* ops/decrypt.hpp:12-89. Returns 0, or -1 where the reference aborts (parent out of range / cycle). */
static int layer_R(const orc_keys* k, const orc_ct* C, uint32_t lid, int* vis, fp_t* cache, fp_t* out) {
    if (lid >= C->nL) return -1;
    if (!fp_is_zero(cache[lid])) { *out = cache[lid]; return 0; }
    if (vis[lid]) return -1;
    vis[lid] = 1;
    const layer_t* L = &C->L[lid];
    fp_t R;
    if (L->rule == 0) R = prf_R(k, L->ztag, L->nlo, L->nhi);
    else {
        fp_t Ra, Rb;
        if (layer_R(k, C, L->pa, vis, cache, &Ra)) return -1;
        if (layer_R(k, C, L->pb, vis, cache, &Rb)) return -1;
        R = fp_mul(Ra, Rb);
    }
    vis[lid] = 0; cache[lid] = R; *out = R;
    return 0;
}
int orc_dec_value(const orc_keys* k, const orc_ct* C, uint64_t* out) {
    uint32_t L = C->nL;
    fp_t* cache = (fp_t*)calloc(L ? L : 1, sizeof(fp_t));
    fp_t* rinv = (fp_t*)calloc(L ? L : 1, sizeof(fp_t));
    int* vis = (int*)calloc(L ? L : 1, sizeof(int));
    int rc = 0;
    for (uint32_t l = 0; l < L && !rc; l++) { fp_t R; rc = layer_R(k, C, l, vis, cache, &R); if (!rc) rinv[l] = fp_inv(R); }
    fp_t acc = fp_from_u64(0);
    for (uint32_t i = 0; i < C->nE && !rc; i++) {
        const edge_t* e = &C->E[i];
        fp_t term = fp_mul(fp_mul(e->w, k->powg[e->idx]), rinv[e->lid]);
        acc = e->ch == 0 ? fp_add(acc, term) : fp_sub(acc, term);
    }
    out[0] = acc.lo; out[1] = acc.hi;
    free(cache); free(rinv); free(vis);
    return rc;
}

/* ------------------------------------------------------------------ struct-of-arrays import / export (same shape as oracle/ref_shim.cpp) */
void orc_ct_counts(const orc_ct* c, uint32_t* nL, uint32_t* nE) { *nL = c->nL; *nE = c->nE; }
void orc_ct_export(const orc_ct* c, uint8_t* rule, uint64_t* ztag, uint64_t* nlo, uint64_t* nhi, uint32_t* pa, uint32_t* pb,
                   uint32_t* lid, uint16_t* idx, uint8_t* ch, uint64_t* w, uint64_t* sigma) {
    for (uint32_t i = 0; i < c->nL; i++) {
        const layer_t* L = &c->L[i];
        rule[i] = L->rule; ztag[i] = L->ztag; nlo[i] = L->nlo; nhi[i] = L->nhi;
        pa[i] = L->rule == 1 ? L->pa : 0; pb[i] = L->rule == 1 ? L->pb : 0;
    }
    for (uint32_t i = 0; i < c->nE; i++) {
        const edge_t* e = &c->E[i];
        lid[i] = e->lid; idx[i] = e->idx; ch[i] = e->ch; w[2 * i] = e->w.lo; w[2 * i + 1] = e->w.hi;
        if (sigma) memcpy(sigma + (size_t)i * ORC_M_WORDS, e->s, ORC_M_WORDS * 8);
    }
}
orc_ct* orc_ct_import(uint32_t nL, uint32_t nE, const uint8_t* rule, const uint64_t* ztag, const uint64_t* nlo, const uint64_t* nhi,
                      const uint32_t* pa, const uint32_t* pb, const uint32_t* lid, const uint16_t* idx, const uint8_t* ch,
                      const uint64_t* w, const uint64_t* sigma) {
    orc_ct* c = ct_new();
    for (uint32_t i = 0; i < nL; i++) { layer_t l; memset(&l, 0, sizeof l); l.rule = rule[i]; l.ztag = ztag[i]; l.nlo = nlo[i]; l.nhi = nhi[i]; l.pa = pa[i]; l.pb = pb[i]; ct_push_layer(c, l); }
    for (uint32_t i = 0; i < nE; i++) {
        edge_t* e = ct_push_edge(c);
        e->lid = lid[i]; e->idx = idx[i]; e->ch = ch[i]; e->w.lo = w[2 * i]; e->w.hi = w[2 * i + 1];
        if (sigma) memcpy(e->s, sigma + (size_t)i * ORC_M_WORDS, ORC_M_WORDS * 8); else memset(e->s, 0, ORC_M_WORDS * 8);
    }
    return c;
}

/* ------------------------------------------------------------------ flat entry points for ctypes */
void orc_fp_mul(const uint64_t* a, const uint64_t* b, uint64_t* o) { fp_t x = { a[0], a[1] }, y = { b[0], b[1] }; fp_t r = fp_mul(x, y); o[0] = r.lo; o[1] = r.hi; }
void orc_fp_add(const uint64_t* a, const uint64_t* b, uint64_t* o) { fp_t x = { a[0], a[1] }, y = { b[0], b[1] }; fp_t r = fp_add(x, y); o[0] = r.lo; o[1] = r.hi; }
void orc_fp_sub(const uint64_t* a, const uint64_t* b, uint64_t* o) { fp_t x = { a[0], a[1] }, y = { b[0], b[1] }; fp_t r = fp_sub(x, y); o[0] = r.lo; o[1] = r.hi; }
void orc_fp_neg(const uint64_t* a, uint64_t* o) { fp_t x = { a[0], a[1] }; fp_t r = fp_neg(x); o[0] = r.lo; o[1] = r.hi; }
void orc_fp_inv(const uint64_t* a, uint64_t* o) { fp_t x = { a[0], a[1] }; fp_t r = fp_inv(x); o[0] = r.lo; o[1] = r.hi; }
void orc_fp_from_words(uint64_t lo, uint64_t hi, uint64_t* o) { fp_t r = fp_from_words(lo, hi); o[0] = r.lo; o[1] = r.hi; }
void orc_hash_to_fp_nonzero(uint64_t lo, uint64_t hi, uint64_t* o) { fp_t r = hash_to_fp_nonzero(lo, hi); o[0] = r.lo; o[1] = r.hi; }
void orc_prf_R_core(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, const char* dom, uint64_t* o) { fp_t r = orc_prf_R_core_fp(k, ztag, nlo, nhi, dom); o[0] = r.lo; o[1] = r.hi; }
void orc_prf_R(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint64_t* o) { fp_t r = prf_R(k, ztag, nlo, nhi); o[0] = r.lo; o[1] = r.hi; }
void orc_prf_R_noise(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint64_t* o) { fp_t r = prf_R_noise(k, ztag, nlo, nhi); o[0] = r.lo; o[1] = r.hi; }
void orc_prf_noise_delta(const orc_keys* k, uint64_t ztag, uint64_t nlo, uint64_t nhi, uint32_t gid, uint8_t kind, uint64_t* o) { fp_t r = prf_noise_delta(k, ztag, nlo, nhi, gid, kind); o[0] = r.lo; o[1] = r.hi; }
uint64_t orc_tape_word(uint64_t* state) { *state += 0x9E3779B97F4A7C15ull; return mix64(*state); }   /* SplitMix64 stepping, for the tests */
