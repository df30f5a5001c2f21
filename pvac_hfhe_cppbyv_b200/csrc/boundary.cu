// The rest of the drop-in boundary (include/pvacb.h): Params, key generation from a seed, the reference's pk / sk file formats,
// the RNG tape of a context, whole-batch blob export / import (one copy per batch, optionally relayed through another GPU's
// host link), test hooks.
#include "engine.h"
#include "../../include/pvacb.h"

#include <sys/random.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace pvacb {
int keygen_host(Tape& t, std::vector<uint64_t>& blob);   // keygen.cpp
int keygen_host(uint64_t tape_state, std::vector<uint64_t>& blob);
}
using namespace pvacb;

static inline Ctx* C(pvacb_ctx* x) { return reinterpret_cast<Ctx*>(x); }
static inline const Ctx* C(const pvacb_ctx* x) { return reinterpret_cast<const Ctx*>(x); }
static inline Batch* Bt(pvacb_batch* x) { return reinterpret_cast<Batch*>(x); }
static inline const Batch* Bt(const pvacb_batch* x) { return reinterpret_cast<const Batch*>(x); }

static int os_random(void* buf, size_t n) {
    uint8_t* p = static_cast<uint8_t*>(buf);
    while (n) {
        ssize_t r = getrandom(p, n, 0);
        if (r <= 0) return PV_E_ARG;
        p += r; n -= (size_t)r;
    }
    return PV_OK;
}

// ------------------------------------------------------------------ Params (core/types.hpp:36-70)
// The kernels are specialised to the shapes of the default set (B, m_bits, n_bits, the column weights, lpn_n, tau); the
// entropy budget of plan_noise, edge_budget, lpn_t and the recrypt fields are run-time values of the context.
static int check_params(Ctx* ctx, const pvacb_params* p) {
    auto bad = [&](const char* what) { ctx->last_error = std::string("Params: ") + what; return (int)PV_E_ARG; };
    if (p->B != kB) return bad("B must be 337 (the carrier group order the kernels are built for)");
    if (p->m_bits != kMBits || p->n_bits != kNBits) return bad("m_bits / n_bits must be 8192 / 16384");
    if (p->h_col_wt != kHColWt || p->x_col_wt != kXColWt || p->err_wt != kErrWt) return bad("h_col_wt / x_col_wt / err_wt must be 192 / 128 / 128");
    if (p->lpn_n != kLpnN) return bad("lpn_n must be 4096");
    if (p->lpn_tau_num != 1 || p->lpn_tau_den != 8) return bad("lpn_tau must be 1/8");
    if (p->lpn_t < 127 || p->lpn_t > kLpnT) return bad("lpn_t must be in 127..16384 (fewer than 127 rows change the PRF output; more than 16384 are not built)");
    if (!(p->noise_entropy_bits >= 0.0) || !(p->depth_slope_bits >= 0.0) || !(p->tuple2_fraction >= 0.0 && p->tuple2_fraction <= 1.0))
        return bad("noise_entropy_bits, depth_slope_bits must be >= 0 and tuple2_fraction in [0, 1]");
    if (p->edge_budget < 4096 || p->edge_budget > 0xFFFFFFFFull) return bad("edge_budget must be in 4096..2^32-1 (enc_value's own guard_budget is not built)");
    return PV_OK;
}
static void apply_params(Ctx* ctx, const pvacb_params* p) {
    ctx->noise_entropy_bits = p->noise_entropy_bits;
    ctx->tuple2_fraction = p->tuple2_fraction;
    ctx->depth_slope_bits = p->depth_slope_bits;
    ctx->edge_budget = (uint32_t)p->edge_budget;
    ctx->lpn_t = p->lpn_t;
    ctx->recrypt_lo = p->recrypt_lo; ctx->recrypt_hi = p->recrypt_hi; ctx->recrypt_rounds = p->recrypt_rounds;
    // every lpn_t >= 127 gives the same PRF values (toep_127 reads rows 0..126 only): fewer rows than the reference's 16384 means the
    // live-row evaluation; with lpn_t = 16384 the mode of the context stays what it is (all rows by default, pvacb_set_prf_mode)
    if (p->lpn_t != kLpnT) ctx->prf_mode = PRF_LIVE;
}

namespace pvacb {
__global__ void blob_validate_kernel(uint64_t n, uint64_t nL, uint64_t nE, const uint32_t* __restrict__ loff, const uint32_t* __restrict__ eoff,
                                     const uint8_t* __restrict__ rule, const uint32_t* __restrict__ pa, const uint32_t* __restrict__ pb,
                                     const uint32_t* __restrict__ lid, const uint16_t* __restrict__ idx, const uint8_t* __restrict__ ch,
                                     const Fp* __restrict__ w, unsigned int* __restrict__ err) {
    const uint64_t i = blockIdx.x;
    if (i == 0 && threadIdx.x == 0 && (loff[0] != 0 || eoff[0] != 0 || loff[n] != nL || eoff[n] != nE)) atomicOr(err, 1u);
    const uint32_t l0 = loff[i], l1 = loff[i + 1], e0 = eoff[i], e1 = eoff[i + 1];
    if (l1 < l0 || e1 < e0 || l1 > nL || e1 > nE) { if (threadIdx.x == 0) atomicOr(err, 1u); return; }
    const uint32_t L = l1 - l0;
    for (uint32_t l = l0 + threadIdx.x; l < l1; l += blockDim.x) {
        if (rule[l] > 1) atomicOr(err, 2u);
        else if (rule[l] == 1 && (pa[l] >= L || pb[l] >= L)) atomicOr(err, 8u);
    }
    for (uint32_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        if (lid[e] >= L || idx[e] >= kB || ch[e] > 1) atomicOr(err, 4u);
        const Fp x = w[e];
        if ((x.hi >> 63) || (x.hi == kMask63 && x.lo == ~0ull)) atomicOr(err, 16u);
    }
}

// device-side validation of a batch that came from outside (what the kernels rely on; the reference would index out of bounds)
int batch_validate(Ctx* ctx, const Batch* b) {
    if (b->n == 0) return PV_OK;
    Scratch scratch(ctx);
    unsigned int* err = nullptr;
    int rc;
    if ((rc = scratch.alloc(err, 4))) return rc;
    PV_CUDA(cudaMemsetAsync(err, 0, 4, ctx->stream));
    blob_validate_kernel<<<(unsigned)b->n, 128, 0, ctx->stream>>>(b->n, b->nL, b->nE, b->loff, b->eoff, b->rule, b->pa, b->pb, b->lid, b->idx, b->ch, b->w, err);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    unsigned int h = 0;
    { SmallRead sr; sr.add(&h, err, 4); if ((rc = read_small_sync(ctx, sr))) return rc; }
    if (h) {
        ctx->last_error = (h & 1) ? "offsets not monotone or inconsistent with the totals" : (h & 2) ? "unknown layer rule" : (h & 8) ? "PROD layer parent out of range"
                        : (h & 4) ? "edge field out of range" : "edge weight not canonical";
        return PV_E_FORMAT;
    }
    return PV_OK;
}

}  // namespace pvacb

extern "C" {

void pvacb_params_default(pvacb_params* p) {
    if (!p) return;
    p->B = kB; p->m_bits = kMBits; p->n_bits = kNBits; p->h_col_wt = kHColWt; p->x_col_wt = kXColWt; p->err_wt = kErrWt;
    p->noise_entropy_bits = 120.0; p->tuple2_fraction = 0.55; p->depth_slope_bits = 16.0;
    p->edge_budget = kEdgeBudget;
    p->lpn_n = kLpnN; p->lpn_t = kLpnT; p->lpn_tau_num = 1; p->lpn_tau_den = 8;
    p->recrypt_lo = 0.48; p->recrypt_hi = 0.52; p->recrypt_rounds = 8;
}

int pvacb_set_params(pvacb_ctx* x, const pvacb_params* prm) {
    Ctx* ctx = C(x);
    if (!ctx || !prm) return PV_E_ARG;
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    apply_params(ctx, prm);
    return PV_OK;
}

int pvacb_get_params(const pvacb_ctx* x, pvacb_params* p) {
    const Ctx* ctx = C(x);
    if (!ctx || !p) return PV_E_ARG;
    pvacb_params_default(p);
    p->noise_entropy_bits = ctx->noise_entropy_bits; p->tuple2_fraction = ctx->tuple2_fraction; p->depth_slope_bits = ctx->depth_slope_bits;
    p->edge_budget = ctx->edge_budget; p->lpn_t = ctx->lpn_t;
    p->recrypt_lo = ctx->recrypt_lo; p->recrypt_hi = ctx->recrypt_hi; p->recrypt_rounds = ctx->recrypt_rounds;
    return PV_OK;
}

// keygen(const Params&, PubKey&, SecKey&), crypto/keygen.hpp:35. Every word is drawn from ChaCha20 keyed by `seed` (32 bytes;
// NULL = 32 fresh bytes from the OS CSPRNG), in the reference's draw order.
int pvacb_keygen_params(pvacb_ctx* x, const pvacb_params* prm, const uint8_t seed[32]) {
    Ctx* ctx = C(x);
    if (!ctx) return PV_E_ARG;
    pvacb_params def;
    pvacb_params_default(&def);
    if (!prm) prm = &def;
    int rc = check_params(ctx, prm);
    if (rc) return rc;
    TapeSpec ts;
    ts.kind = TAPE_CHACHA20;
    if (seed) memcpy(ts.key, seed, 32);
    else if ((rc = os_random(ts.key, 32))) { ctx->last_error = "getrandom failed"; return rc; }
    const uint64_t sid = 0;
    ts.states = &sid;
    Tape t = tape_open(ts, 0);
    std::vector<uint64_t> blob;
    if ((rc = keygen_host(t, blob))) return rc;
    memset(ts.key, 0, sizeof ts.key);
    apply_params(ctx, prm);
    return keys_from_host_blob(ctx, blob.data());
}

// ------------------------------------------------------------------ pk / sk files (tests/bounty2_test.cpp:145-236)
static const uint32_t kMagicSK = 0x66666999u, kMagicPK = 0x06660666u, kFileVer = 1;

struct Writer {
    FILE* f; bool ok = true;
    void put(const void* p, size_t n) { if (ok && fwrite(p, 1, n, f) != n) ok = false; }
    void u32(uint32_t v) { put(&v, 4); }
    void u64(uint64_t v) { put(&v, 8); }
};
struct Reader {
    FILE* f; bool ok = true;
    void get(void* p, size_t n) { if (ok && fread(p, 1, n, f) != n) ok = false; }
    uint32_t u32() { uint32_t v = 0; get(&v, 4); return v; }
    uint64_t u64() { uint64_t v = 0; get(&v, 8); return v; }
};

int pvacb_keys_export_file(pvacb_ctx* x, const char* pk_path, const char* sk_path) {
    Ctx* ctx = C(x);
    if (!ctx) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    const uint64_t* h = ctx->h_hdr.data();
    if (sk_path) {                              // saveSk, :145-152
        if (!ctx->have_sk) { ctx->last_error = "this context holds a public key only"; return PV_E_NOKEYS; }
        FILE* f = fopen(sk_path, "wb");
        if (!f) { ctx->last_error = std::string("cannot write ") + sk_path; return PV_E_ARG; }
        Writer w{f};
        w.u32(kMagicSK); w.u32(kFileVer);
        for (int j = 0; j < 4; j++) w.u64(h[5 + j]);
        w.u64(kLpnWords);
        for (int j = 0; j < kLpnWords; j++) w.u64(h[9 + j]);
        const bool ok = w.ok;
        if (fclose(f) != 0 || !ok) { ctx->last_error = std::string("write failed: ") + sk_path; return PV_E_ARG; }
    }
    if (pk_path) {                              // savePk, :165-192
        std::vector<uint64_t> H((size_t)kNBits * kMWords);
        cudaSetDevice(ctx->device);
        PV_CUDA(cudaMemcpy(H.data(), ctx->d_blob + kBlobHdrWords, H.size() * 8, cudaMemcpyDeviceToHost));
        FILE* f = fopen(pk_path, "wb");
        if (!f) { ctx->last_error = std::string("cannot write ") + pk_path; return PV_E_ARG; }
        Writer w{f};
        w.u32(kMagicPK); w.u32(kFileVer);
        w.u32(kMBits); w.u32(kB); w.u32((uint32_t)ctx->lpn_t); w.u32(kLpnN); w.u32(1); w.u32(8);
        w.u32((uint32_t)ctx->noise_entropy_bits); w.u32((uint32_t)ctx->depth_slope_bits);
        uint64_t t2; memcpy(&t2, &ctx->tuple2_fraction, 8);
        w.u64(t2);
        w.u32(ctx->edge_budget);
        w.u64(h[0]);
        w.put(&h[1], 32);
        w.u64(kNBits);
        for (int c = 0; c < kNBits; c++) { w.u32(kMBits); w.put(&H[(size_t)c * kMWords], kMWords * 8); }
        std::vector<uint32_t> inv(kMBits);
        for (int i = 0; i < kMBits; i++) inv[ctx->h_ubk_perm[i]] = (uint32_t)i;
        w.u64(kMBits);
        for (int i = 0; i < kMBits; i++) w.u32(ctx->h_ubk_perm[i]);
        w.u64(kMBits);
        for (int i = 0; i < kMBits; i++) w.u32(inv[i]);
        w.u64(h[748]); w.u64(h[749]);           // omega_B
        w.u64(kB);
        for (int i = 0; i < kB; i++) { w.u64(h[73 + 2 * i]); w.u64(h[73 + 2 * i + 1]); }
        const bool ok = w.ok;
        if (fclose(f) != 0 || !ok) { ctx->last_error = std::string("write failed: ") + pk_path; return PV_E_ARG; }
    }
    return PV_OK;
}

// loadPk / loadSk, :154-163,194-236. sk_path may be NULL: the context then holds the public key only (ct_add / ct_sub / ct_mul /
// commit_ct / recrypt work; enc_* and dec_* return PVACB_E_NOKEYS). Params stored in the pk file are checked like
// pvacb_keygen_params checks them; the UBK permutation in the file must be the one canon_tag derives.
int pvacb_keys_import_file(pvacb_ctx* x, const char* pk_path, const char* sk_path) {
    Ctx* ctx = C(x);
    if (!ctx || !pk_path) return PV_E_ARG;
    auto fmt = [&](const std::string& what) { ctx->last_error = what; return (int)PV_E_FORMAT; };
    std::vector<uint64_t> blob(kBlobWords, 0);
    pvacb_params prm;
    pvacb_params_default(&prm);
    {
        FILE* f = fopen(pk_path, "rb");
        if (!f) { ctx->last_error = std::string("cannot read ") + pk_path; return PV_E_ARG; }
        Reader r{f};
        struct Close { FILE* f; ~Close() { fclose(f); } } closer{f};
        if (r.u32() != kMagicPK || r.u32() != kFileVer || !r.ok) return fmt(std::string("bad PK: ") + pk_path);
        prm.m_bits = (int32_t)r.u32(); prm.B = (int32_t)r.u32(); prm.lpn_t = (int32_t)r.u32(); prm.lpn_n = (int32_t)r.u32();
        prm.lpn_tau_num = (int32_t)r.u32(); prm.lpn_tau_den = (int32_t)r.u32();
        prm.noise_entropy_bits = (double)r.u32(); prm.depth_slope_bits = (double)r.u32();
        uint64_t t2 = r.u64();
        memcpy(&prm.tuple2_fraction, &t2, 8);
        prm.edge_budget = r.u32();
        if (!r.ok) return fmt("PK file ends inside the header");
        int rc = check_params(ctx, &prm);
        if (rc) return rc;
        blob[0] = r.u64();
        r.get(&blob[1], 32);
        if (r.u64() != (uint64_t)kNBits || !r.ok) return fmt("PK: H does not have n_bits columns");
        for (int c = 0; c < kNBits; c++) {
            if (r.u32() != (uint32_t)kMBits) return fmt("PK: a column of H does not have m_bits bits");
            r.get(&blob[kBlobHdrWords + (size_t)c * kMWords], kMWords * 8);
            if (!r.ok) return fmt("PK file ends inside H");
        }
        std::vector<uint16_t> want(kMBits);
        gen_ubk_perm_host(blob[0], want.data());
        if (r.u64() != (uint64_t)kMBits) return fmt("PK: ubk.perm does not have m_bits entries");
        for (int i = 0; i < kMBits; i++)
            if (r.u32() != want[i]) return fmt("PK: ubk.perm is not the permutation canon_tag derives (crypto/matrix.hpp:95-164)");
        if (r.u64() != (uint64_t)kMBits) return fmt("PK: ubk.inv does not have m_bits entries");
        for (int i = 0; i < kMBits; i++) {
            const uint32_t v = r.u32();
            if (v >= (uint32_t)kMBits || want[v] != i) return fmt("PK: ubk.inv is not the inverse of ubk.perm");
        }
        blob[748] = r.u64(); blob[749] = r.u64();
        if (r.u64() != (uint64_t)kB || !r.ok) return fmt("PK: powg_B does not have B entries");
        for (int i = 0; i < kB; i++) { blob[73 + 2 * i] = r.u64(); blob[73 + 2 * i + 1] = r.u64(); }
        uint8_t extra;
        if (!r.ok) return fmt("PK file ends inside powg_B");
        if (fread(&extra, 1, 1, f) == 1) return fmt("PK: trailing bytes");
        for (int i = 0; i < kB; i++) {
            const uint64_t lo = blob[73 + 2 * i], hi = blob[73 + 2 * i + 1];
            if ((hi >> 63) || (hi == kMask63 && lo == ~0ull)) return fmt("PK: powg_B entry is not a canonical field element");
        }
    }
    bool have_sk = false;
    if (sk_path) {
        FILE* f = fopen(sk_path, "rb");
        if (!f) { ctx->last_error = std::string("cannot read ") + sk_path; return PV_E_ARG; }
        Reader r{f};
        struct Close { FILE* f; ~Close() { fclose(f); } } closer{f};
        if (r.u32() != kMagicSK || r.u32() != kFileVer || !r.ok) return fmt(std::string("bad SK: ") + sk_path);
        for (int j = 0; j < 4; j++) blob[5 + j] = r.u64();
        if (r.u64() != (uint64_t)kLpnWords || !r.ok) return fmt("SK: lpn_s_bits does not have lpn_n / 64 words");
        for (int j = 0; j < kLpnWords; j++) blob[9 + j] = r.u64();
        uint8_t extra;
        if (!r.ok) return fmt("SK file ends inside lpn_s_bits");
        if (fread(&extra, 1, 1, f) == 1) return fmt("SK: trailing bytes");
        have_sk = true;
    }
    apply_params(ctx, &prm);
    int rc = keys_from_host_blob(ctx, blob.data());
    if (rc == PV_OK) ctx->have_sk = have_sk;
    return rc;
}

// ------------------------------------------------------------------ RNG tape of the context (common.cuh)
int pvacb_set_tape(pvacb_ctx* x, int kind, const uint8_t key[32]) {
    Ctx* ctx = C(x);
    if (!ctx) return PV_E_ARG;
    if (kind != TAPE_SPLITMIX && kind != TAPE_CHACHA20 && kind != TAPE_WORDS) { ctx->last_error = "unknown tape kind"; return PV_E_ARG; }
    if (kind == TAPE_CHACHA20) {
        if (key) memcpy(ctx->tape_key, key, 32);
        else if (os_random(ctx->tape_key, 32)) { ctx->last_error = "getrandom failed"; return PV_E_ARG; }
    }
    ctx->tape_kind = kind;
    return PV_OK;
}
int pvacb_get_tape(const pvacb_ctx* x) { return x ? C(x)->tape_kind : -1; }

int pvacb_set_tape_words(pvacb_ctx* x, const uint64_t* words, size_t n_items, size_t words_per_item) {
    Ctx* ctx = C(x);
    if (!ctx || (n_items * words_per_item && !words)) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->d_tape_words) { cudaFree(ctx->d_tape_words); ctx->d_tape_words = nullptr; }
    ctx->tape_words_items = ctx->tape_words_per_item = 0;
    if (n_items * words_per_item) {
        PV_CUDA(cudaMalloc((void**)&ctx->d_tape_words, n_items * words_per_item * 8));
        PV_CUDA(cudaMemcpy(ctx->d_tape_words, words, n_items * words_per_item * 8, cudaMemcpyHostToDevice));
        ctx->tape_words_items = n_items; ctx->tape_words_per_item = words_per_item;
    }
    return PV_OK;
}
int pvacb_set_item_base(pvacb_ctx* x, uint64_t base) {
    if (!x) return PV_E_ARG;
    C(x)->item_base = base;
    return PV_OK;
}
uint64_t pvacb_fresh_seed(pvacb_ctx* x) { return x ? ++C(x)->seed_counter : 0; }

// test hooks. what = 0: ct_mul planning uses the device-wide sort (a = 1) / the global bucket table (b = 1) even for small pairs;
// what = 1: keystream word `a` of every PRF core is OR-ed with `b` (a = ~0: off) -- reaches AesCtr256::bounded's rejection branch;
// what = 2: sigma kernel shape without spare PRG candidates (a = 1), so that the in-kernel stream continuation runs. All bit-exact.
int pvacb_debug_set(pvacb_ctx* x, int what, uint64_t a, uint64_t b) {
    Ctx* ctx = C(x);
    if (!ctx) return PV_E_ARG;
    if (what == 0) { ctx->mul_force_device_sort = a != 0; ctx->mul_force_global_table = b != 0; return PV_OK; }
    if (what == 1) { ctx->prf_patch_word = a; ctx->prf_patch_or = b; return PV_OK; }
    if (what == 2) { ctx->sigma_test_shape = a != 0; return PV_OK; }
    return PV_E_ARG;
}

// ------------------------------------------------------------------ whole-batch blob: ONE copy per batch
// A batch is one device allocation (engine.h: struct Batch). Its image is exported / imported as it is; the 13 field offsets
// inside the image come from pvacb_batch_blob_layout (the same function of (n, layers, edges) on both sides).
static void blob_offsets(uint64_t n, uint64_t nL, uint64_t nE, uint64_t off[14]) {
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { uint64_t at = o; o += ((bytes ? bytes : 1) + 255) & ~(uint64_t)255; return at; };
    off[0] = take((n + 1) * 4); off[1] = take((n + 1) * 4);
    off[2] = take(nL); off[3] = take(nL * 8); off[4] = take(nL * 8); off[5] = take(nL * 8); off[6] = take(nL * 4); off[7] = take(nL * 4);
    off[8] = take(nE * 4); off[9] = take(nE * 2); off[10] = take(nE); off[11] = take(nE * 16); off[12] = take(nE * (uint64_t)kMWords * 8);
    off[13] = o;
}

int pvacb_blob_layout(uint64_t n, uint64_t n_layers, uint64_t n_edges, uint64_t off[14]) {
    if (!off) return PV_E_ARG;
    blob_offsets(n, n_layers, n_edges, off);
    return PV_OK;
}

// (n, layers, edges) the image of this batch was laid out for, and its size: compact_layers shrinks the layer COUNT of a batch in
// place without moving its arrays, so the layout counts can exceed pvacb_batch_totals.
int pvacb_batch_blob_info(const pvacb_batch* pb, uint64_t* n, uint64_t* layout_layers, uint64_t* layout_edges, uint64_t* bytes) {
    const Batch* b = Bt(pb);
    if (!b) return PV_E_ARG;
    if (n) *n = b->n;
    if (layout_layers) *layout_layers = b->nL_alloc;
    if (layout_edges) *layout_edges = b->nE;
    if (bytes) *bytes = b->bytes;
    return PV_OK;
}

static int relay_prepare(Ctx* ctx, size_t bytes, int slot) {
    Ctx::Relay& r = ctx->relay;
    PV_CUDA(cudaSetDevice(r.device));
    if (!r.stream[0]) {
        cudaError_t e = cudaDeviceEnablePeerAccess(ctx->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaSetDevice(ctx->device); ctx->last_error = std::string("relay: peer access: ") + cudaGetErrorString(e); return PV_E_CUDA; }
        cudaGetLastError();
        PV_CUDA(cudaStreamCreateWithFlags(&r.stream[0], cudaStreamNonBlocking));
        PV_CUDA(cudaStreamCreateWithFlags(&r.stream[1], cudaStreamNonBlocking));
        // batches live in the source device's stream-ordered memory pool, which cudaDeviceEnablePeerAccess does NOT open to peers:
        // without this grant the peer copy below is staged through host memory by the driver (twice over the slow link -- r02
        // measured 77 GB/s relayed against 92 GB/s direct on eight GPUs before the grant was added)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) {
            cudaMemAccessDesc acc = {};
            acc.location.type = cudaMemLocationTypeDevice;
            acc.location.id = r.device;
            acc.flags = cudaMemAccessFlagsProtReadWrite;
            cudaError_t pe = cudaMemPoolSetAccess(pool, &acc, 1);
            if (pe != cudaSuccess) { cudaGetLastError(); cudaSetDevice(ctx->device); ctx->last_error = std::string("relay: cudaMemPoolSetAccess: ") + cudaGetErrorString(pe); return PV_E_CUDA; }
        }
    }
    if (r.cap[slot] < bytes) {
        if (r.stage[slot]) { PV_CUDA(cudaStreamSynchronize(r.stream[slot])); PV_CUDA(cudaFree(r.stage[slot])); r.stage[slot] = nullptr; r.cap[slot] = 0; }
        const size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&r.stage[slot], want);
        if (e != cudaSuccess) { cudaSetDevice(ctx->device); ctx->last_error = "relay: staging buffer allocation failed"; return PV_E_OOM; }
        r.cap[slot] = want;
    }
    return PV_OK;
}

// The whole device image of the batch into `host` (pinned, >= the image size) with one copy on the context's second stream,
// ordered after everything queued so far; returns at once (pvacb_export_wait / _wait_one as for pvacb_batch_export_soa_async).
// With a relay device set (pvacb_set_export_relay) the image first crosses NVLink into a staging buffer on that GPU and
// leaves through ITS host link.
int pvacb_batch_export_blob_async(pvacb_ctx* x, const pvacb_batch* pb, void* host, size_t cap) {
    Ctx* ctx = C(x);
    const Batch* b = Bt(pb);
    if (!ctx || !b || !host) return PV_E_ARG;
    if (check_owner(ctx, b)) return PV_E_ARG;
    if (cap < b->bytes) { ctx->last_error = "export buffer smaller than the batch image"; return PV_E_ARG; }
    cudaSetDevice(ctx->device);
    cudaEvent_t ready, done;
    PV_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    PV_CUDA(cudaEventRecord(ready, ctx->stream));
    if (ctx->relay.device >= 0) {
        // one in-order stream on the relay GPU and one staging buffer: peer copy, then the device->host copy, export after export --
        // the configuration profiles/r02_hostlink_relay.txt measured (the NVLink hop costs ~5 % of the time of the PCIe hop)
        Ctx::Relay& r = ctx->relay;
        const int slot = 0;
        int rc = relay_prepare(ctx, b->bytes, slot);
        if (rc) { cudaEventDestroy(ready); return rc; }
        cudaError_t e = cudaStreamWaitEvent(r.stream[slot], ready, 0);
        if (e == cudaSuccess) e = cudaMemcpyPeerAsync(r.stage[slot], r.device, b->base, ctx->device, b->bytes, r.stream[slot]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(host, r.stage[slot], b->bytes, cudaMemcpyDeviceToHost, r.stream[slot]);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(done, r.stream[slot]);
        cudaSetDevice(ctx->device);
        cudaEventDestroy(ready);
        PV_CUDA(e);
    } else {
        PV_CUDA(cudaStreamWaitEvent(ctx->stream2, ready, 0));
        cudaEventDestroy(ready);
        PV_CUDA(cudaMemcpyAsync(host, b->base, b->bytes, cudaMemcpyDeviceToHost, ctx->stream2));
        PV_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        PV_CUDA(cudaEventRecord(done, ctx->stream2));
    }
    ctx->export_events.push_back(done);
    return PV_OK;
}

int pvacb_set_export_relay(pvacb_ctx* x, int relay_device) {
    Ctx* ctx = C(x);
    if (!ctx) return PV_E_ARG;
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (relay_device >= ndev || relay_device == ctx->device) { ctx->last_error = "relay device must be another visible GPU (or -1)"; return PV_E_ARG; }
    if (relay_device >= 0) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, relay_device, ctx->device);
        if (!can) { ctx->last_error = "relay device has no peer access to this GPU"; return PV_E_ARG; }
    }
    Ctx::Relay& r = ctx->relay;
    if (r.device >= 0 && r.device != relay_device) {       // drop the old relay's buffers
        cudaSetDevice(r.device);
        for (int s = 0; s < 2; s++) {
            if (r.stream[s]) { cudaStreamSynchronize(r.stream[s]); cudaStreamDestroy(r.stream[s]); r.stream[s] = nullptr; }
            if (r.stage[s]) { cudaFree(r.stage[s]); r.stage[s] = nullptr; r.cap[s] = 0; }
        }
        cudaSetDevice(ctx->device);
    }
    r.device = relay_device < 0 ? -1 : relay_device;
    return PV_OK;
}

// The inverse: one host->device copy of an image laid out by pvacb_blob_layout(n, n_layers, n_edges), validated on the device.
int pvacb_batch_import_blob(pvacb_ctx* x, size_t n, uint64_t n_layers, uint64_t n_edges, const void* host, size_t bytes, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!ctx || !out || !host) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    Batch* b = nullptr;
    int rc = batch_alloc(ctx, n, n_layers, n_edges, &b);
    if (rc) return rc;
    if (bytes != b->bytes) { batch_free(b); ctx->last_error = "image size does not match pvacb_blob_layout(n, layers, edges)"; return PV_E_ARG; }
    const uint32_t logical = static_cast<const uint32_t*>(host)[n];          // loff[n]: compact_layers may have shrunk the layer count in place
    if (logical > n_layers) { batch_free(b); ctx->last_error = "image holds more layers than its layout"; return PV_E_FORMAT; }
    b->nL = logical;
    cudaError_t e = cudaMemcpyAsync(b->base, host, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) { batch_free(b); ctx->last_error = cudaGetErrorString(e); return PV_E_CUDA; }
    if ((rc = batch_validate(ctx, b))) { batch_free(b); return rc; }
    *out = reinterpret_cast<pvacb_batch*>(b);
    return PV_OK;
}

}  // extern "C"
