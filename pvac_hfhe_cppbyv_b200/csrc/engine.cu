// Host side of the pvacb engine: context, key replication blob, device batches (structure-of-arrays), import/export,
// and the extern "C" boundary declared in include/pvacb.h.
#include "engine.h"
#include "sha256.cuh"
#include "../../include/pvacb.h"

#include <cstring>
#include <cstdio>
#include <algorithm>
#include <unordered_map>

namespace pvacb {

int keygen_host(uint64_t tape_state, std::vector<uint64_t>& blob);   // keygen.cu

// ------------------------------------------------------------------ memory
int dev_alloc(Ctx* ctx, void** p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e != cudaSuccess) {
        ctx->last_error = std::string("cudaMallocAsync(") + std::to_string(bytes) + "): " + cudaGetErrorString(e);
        cudaGetLastError();
        *p = nullptr;
        return e == cudaErrorMemoryAllocation ? PV_E_OOM : PV_E_CUDA;
    }
    return PV_OK;
}
void dev_free(Ctx* ctx, void* p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct SmallReadArgs { const uint32_t* src[SmallRead::kMaxItems]; uint32_t words[SmallRead::kMaxItems]; int n; };
__global__ void small_read_kernel(SmallReadArgs a, volatile uint32_t* mail) {
    int o = 0;
    for (int i = 0; i < a.n; i++)
        for (uint32_t k = 0; k < a.words[i]; k++) mail[o++] = a.src[i][k];
    __threadfence_system();
}
int read_small_sync(Ctx* ctx, const SmallRead& r) {
    SmallReadArgs a;
    a.n = r.n;
    uint32_t total = 0;
    for (int i = 0; i < r.n; i++) { a.src[i] = r.src[i]; a.words[i] = r.words[i]; total += r.words[i]; }
    if (r.n > SmallRead::kMaxItems || total > (uint32_t)SmallRead::kMaxWords) { ctx->last_error = "read_small_sync: too many words"; return PV_E_ARG; }
    small_read_kernel<<<1, 1, 0, ctx->stream>>>(a, ctx->d_mail);
    PV_CUDA(cudaGetLastError());
    ctx->stat_kernel_launches += 1;
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t o = 0;
    for (int i = 0; i < r.n; i++) { memcpy(r.dst[i], ctx->h_mail + o, r.words[i] * 4); o += r.words[i]; }
    return PV_OK;
}

int batch_alloc(Ctx* ctx, uint64_t n, uint64_t nL, uint64_t nE, Batch** out) {
    if (n >= (1ull << 32) - 1 || nL >= (1ull << 32) || nE >= (1ull << 32)) {      // per-ciphertext offsets are 32-bit
        ctx->last_error = "batch too large: layer / edge offsets are 32-bit, split the batch into tiles";
        return PV_E_SHAPE;
    }
    Batch* b = new Batch();
    b->ctx = ctx; b->n = n; b->nL = nL; b->nE = nE; b->nL_alloc = nL;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 1); return o; };
    size_t o_loff = take((n + 1) * 4), o_eoff = take((n + 1) * 4);
    size_t o_rule = take(nL), o_ztag = take(nL * 8), o_nlo = take(nL * 8), o_nhi = take(nL * 8), o_pa = take(nL * 4), o_pb = take(nL * 4);
    size_t o_lid = take(nE * 4), o_idx = take(nE * 2), o_ch = take(nE), o_w = take(nE * 16), o_sig = take(nE * (size_t)kMWords * 8);
    b->bytes = off;
    int rc = dev_alloc(ctx, &b->base, off);
    if (rc) { delete b; return rc; }
    char* p = (char*)b->base;
    b->loff = (uint32_t*)(p + o_loff); b->eoff = (uint32_t*)(p + o_eoff);
    b->rule = (uint8_t*)(p + o_rule); b->ztag = (uint64_t*)(p + o_ztag); b->nlo = (uint64_t*)(p + o_nlo); b->nhi = (uint64_t*)(p + o_nhi);
    b->pa = (uint32_t*)(p + o_pa); b->pb = (uint32_t*)(p + o_pb);
    b->lid = (uint32_t*)(p + o_lid); b->idx = (uint16_t*)(p + o_idx); b->ch = (uint8_t*)(p + o_ch); b->w = (Fp*)(p + o_w);
    b->sigma = (uint64_t*)(p + o_sig);
    *out = b;
    return PV_OK;
}
// a new batch with the same contents. Field by field: a batch whose layers were compacted in place (compact_layers_batch)
// keeps the allocation layout of its original layer count, so its bytes are NOT one flat image of (n, nL, nE).
int batch_clone(Ctx* ctx, const Batch* s, Batch** out) {
    Batch* o = nullptr;
    int rc = batch_alloc(ctx, s->n, s->nL, s->nE, &o);
    if (rc) return rc;
    auto cp = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream) : cudaSuccess;
    };
    cudaError_t e = cudaSuccess;
    auto ck = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ck(cp(o->loff, s->loff, (s->n + 1) * 4)); ck(cp(o->eoff, s->eoff, (s->n + 1) * 4));
    ck(cp(o->rule, s->rule, s->nL)); ck(cp(o->ztag, s->ztag, s->nL * 8)); ck(cp(o->nlo, s->nlo, s->nL * 8)); ck(cp(o->nhi, s->nhi, s->nL * 8));
    ck(cp(o->pa, s->pa, s->nL * 4)); ck(cp(o->pb, s->pb, s->nL * 4));
    ck(cp(o->lid, s->lid, s->nE * 4)); ck(cp(o->idx, s->idx, s->nE * 2)); ck(cp(o->ch, s->ch, s->nE)); ck(cp(o->w, s->w, s->nE * 16));
    ck(cp(o->sigma, s->sigma, s->nE * (size_t)kMWords * 8));
    if (e != cudaSuccess) {
        batch_free(o);
        ctx->last_error = std::string("batch_clone: ") + cudaGetErrorString(e);
        return PV_E_CUDA;
    }
    *out = o;
    return PV_OK;
}
void batch_free(Batch* b) {
    if (!b) return;
    // a batch may be freed from any host thread, whatever device is current there (the group's batches die on the caller's thread):
    // the stream-ordered free has to run with the owning device current, or the memory never goes back to its pool
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != b->ctx->device) cudaSetDevice(b->ctx->device);
    dev_free(b->ctx, b->base);
    if (cur >= 0 && cur != b->ctx->device) cudaSetDevice(cur);
    delete b;
}

// ------------------------------------------------------------------ keys
static int derive_key_view(Ctx* ctx) {
    cudaSetDevice(ctx->device);
    const uint64_t* h = ctx->h_hdr.data();
    KeyView& kv = ctx->kv;
    kv.canon_tag = h[0];
    // block 0 of derive_aes_key's message: prf_k[0..3] || canon_tag || H_digest[0..23]  (crypto/lpn.hpp:174-181)
    uint64_t q[8] = {h[5], h[6], h[7], h[8], h[0], h[1], h[2], h[3]};
    uint32_t w[16];
    sha_block_from_le64(q, w);
    ShaState st;
    sha_init(st);
    sha_compress(st, w);
    for (int i = 0; i < 8; i++) kv.kd_mid[i] = st.h[i];
    kv.digest3 = h[4];
    kv.powg = reinterpret_cast<const Fp*>(ctx->d_blob + 73);
    kv.H = ctx->d_blob + kBlobHdrWords;
    kv.T0 = ctx->d_aes->t0;
    kv.sbox = ctx->d_aes->sbox;
    lpn_masks_from_secret(&h[9], ctx->lpn_m);
    ctx->h_ubk_perm.resize(kMBits);
    gen_ubk_perm_host(kv.canon_tag, ctx->h_ubk_perm.data());
    if (!ctx->d_ubk_perm) PV_CUDA(cudaMalloc((void**)&ctx->d_ubk_perm, kMBits * 2));
    PV_CUDA(cudaMemcpy(ctx->d_ubk_perm, ctx->h_ubk_perm.data(), kMBits * 2, cudaMemcpyHostToDevice));
    ctx->have_keys = true;
    ctx->have_sk = true;
    return PV_OK;
}

// The key blob must not straddle a 4 GiB boundary: the sigma kernel forms the addresses of the columns of H with 32-bit
// arithmetic on the low word (one IMAD per column instead of a 64-bit add on the ALU pipe that bounds it).
static int ensure_blob(Ctx* ctx) {
    if (ctx->d_blob) return PV_OK;
    cudaSetDevice(ctx->device);          // the calling host thread may have another device current (the group's worker threads do)
    void* held[4] = {nullptr, nullptr, nullptr, nullptr};
    int nheld = 0;
    int rc = PV_OK;
    for (;;) {
        uint64_t* p = nullptr;
        cudaError_t e = cudaMalloc((void**)&p, kBlobBytes);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); rc = PV_E_OOM; break; }
        const uint64_t a = reinterpret_cast<uint64_t>(p);
        if ((a >> 32) == ((a + kBlobBytes - 1) >> 32)) { ctx->d_blob = p; break; }
        if (nheld == 4) { cudaFree(p); ctx->last_error = "key blob: no placement inside one 4 GiB window"; rc = PV_E_CUDA; break; }
        held[nheld++] = p;                 // keep the straddling block allocated so that the next try lands elsewhere
    }
    for (int i = 0; i < nheld; i++) cudaFree(held[i]);
    return rc;
}

int keys_from_host_blob(Ctx* ctx, const uint64_t* blob) {
    cudaSetDevice(ctx->device);
    int rc = ensure_blob(ctx);
    if (rc) return rc;
    PV_CUDA(cudaMemcpyAsync(ctx->d_blob, blob, kBlobBytes, cudaMemcpyHostToDevice, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->h_hdr.assign(blob, blob + kBlobHdrWords);
    return derive_key_view(ctx);
}

// libstdc++ bucket counts: every value _Prime_rehash_policy::_M_next_bkt can return, ascending.
static void build_prime_table(std::vector<uint64_t>& t) {
    std::__detail::_Prime_rehash_policy pol;
    uint64_t n = 1;
    for (;;) {
        uint64_t p = (uint64_t)pol._M_next_bkt((std::size_t)n);
        if (!t.empty() && p <= t.back()) break;
        t.push_back(p);
        if (p >= (1ull << 40)) break;
        n = p + 1;
    }
}

}  // namespace pvacb

using namespace pvacb;

struct pvacb_ctx { Ctx c; };
struct pvacb_batch { Batch b; };
static inline Ctx* C(pvacb_ctx* x) { return reinterpret_cast<Ctx*>(x); }
static inline const Ctx* C(const pvacb_ctx* x) { return reinterpret_cast<const Ctx*>(x); }
static inline Batch* Bt(pvacb_batch* x) { return reinterpret_cast<Batch*>(x); }
static inline const Batch* Bt(const pvacb_batch* x) { return reinterpret_cast<const Batch*>(x); }

extern "C" {

int pvacb_ctx_create(int device, pvacb_ctx** out) {
    if (!out) return PV_E_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return PV_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PV_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PV_E_CUDA;
    if (prop.major < 10) {
        fprintf(stderr, "pvacb: device %d is sm_%d%d; this library only contains sm_100a code\n", device, prop.major, prop.minor);
        return PV_E_CUDA;
    }
    Ctx* ctx = new Ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    auto fail = [&](const char* what) {
        fprintf(stderr, "pvacb_ctx_create: %s failed: %s\n", what, cudaGetErrorString(cudaGetLastError()));
        pvacb_ctx_destroy(reinterpret_cast<pvacb_ctx*>(ctx));
        return (int)PV_E_CUDA;
    };
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
    if (cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess) return fail("second stream");
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    AesTables t;
    aes_make_tables(t);
    if (cudaMalloc((void**)&ctx->d_aes, sizeof(AesTables)) != cudaSuccess) return fail("cudaMalloc(AES tables)");
    if (cudaMemcpy(ctx->d_aes, &t, sizeof t, cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy(AES tables)");
    std::vector<uint64_t> primes;
    build_prime_table(primes);
    ctx->n_primes = (int)primes.size();
    if (cudaMalloc((void**)&ctx->d_primes, primes.size() * 8) != cudaSuccess) return fail("cudaMalloc(bucket-count table)");
    if (cudaMemcpy(ctx->d_primes, primes.data(), primes.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) return fail("cudaMemcpy(bucket-count table)");
    if (cudaMalloc((void**)&ctx->d_work, 64) != cudaSuccess) return fail("cudaMalloc(work counter)");
    if (cudaHostAlloc((void**)&ctx->h_mail, SmallRead::kMaxWords * 4, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->d_mail, ctx->h_mail, 0) != cudaSuccess) return fail("mapped mailbox");
    // RNG tape: ChaCha20 under 256 fresh bits from the OS (pvacb_set_tape replaces kind / key)
    if (pvacb_set_tape(reinterpret_cast<pvacb_ctx*>(ctx), PVACB_TAPE_CHACHA20, nullptr) != PV_OK) {
        fprintf(stderr, "pvacb_ctx_create: the OS CSPRNG (getrandom) failed\n");
        pvacb_ctx_destroy(reinterpret_cast<pvacb_ctx*>(ctx));
        return PV_E_CUDA;
    }
    *out = reinterpret_cast<pvacb_ctx*>(ctx);
    return PV_OK;
}

void pvacb_ctx_destroy(pvacb_ctx* x) {
    if (!x) return;
    Ctx* ctx = C(x);
    cudaSetDevice(ctx->device);
    if (ctx->relay.device >= 0) {
        cudaSetDevice(ctx->relay.device);
        for (int s = 0; s < 2; s++) {
            if (ctx->relay.stream[s]) { cudaStreamSynchronize(ctx->relay.stream[s]); cudaStreamDestroy(ctx->relay.stream[s]); }
            if (ctx->relay.stage[s]) cudaFree(ctx->relay.stage[s]);
        }
        cudaSetDevice(ctx->device);
    }
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);      // asynchronous exports still in flight
    for (cudaEvent_t e : ctx->export_events) cudaEventDestroy(e);
    memset(ctx->tape_key, 0, sizeof ctx->tape_key);
    cudaFree(ctx->d_tape_words);
    cudaFree(ctx->d_blob);
    cudaFree(ctx->d_aes);
    cudaFree(ctx->d_primes);
    cudaFree(ctx->d_work);
    cudaFree(ctx->d_ubk_perm);
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

const char* pvacb_last_error(const pvacb_ctx* x) { return x ? C(x)->last_error.c_str() : "null context"; }
int pvacb_set_prf_mode(pvacb_ctx* x, int mode) {
    if (!x || (mode != PRF_FAITHFUL && mode != PRF_LIVE)) return PV_E_ARG;
    C(x)->prf_mode = mode;
    return PV_OK;
}
int pvacb_get_prf_mode(const pvacb_ctx* x) { return C(x)->prf_mode; }
void* pvacb_stream(pvacb_ctx* x) { return (void*)C(x)->stream; }
int pvacb_sync(pvacb_ctx* x) {
    Ctx* ctx = C(x);
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
void pvacb_stats(const pvacb_ctx* x, uint64_t* k, uint64_t* a, uint64_t* s) {
    if (k) *k = C(x)->stat_kernel_launches;
    if (a) *a = C(x)->stat_aes_blocks;
    if (s) *s = C(x)->stat_sigma_edges;
}
void pvacb_stats_reset(pvacb_ctx* x) { C(x)->stat_kernel_launches = C(x)->stat_aes_blocks = C(x)->stat_sigma_edges = 0; }

int pvacb_profile_enable(pvacb_ctx* x, int on) {
    Ctx* ctx = C(x);
    ctx->profile = on ? 1 : 0;
    return PV_OK;
}
// sums the CUDA-event time (ms) of every bracketed kernel launch since the last collect, per tag; clears the spans
int pvacb_profile_collect(pvacb_ctx* x, float* ms_out /*8*/, uint32_t* launches_out /*8*/) {
    Ctx* ctx = C(x);
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < PROF_NTAGS; i++) { ms_out[i] = 0.f; if (launches_out) launches_out[i] = 0; }
    for (auto& s : ctx->prof_spans) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.a, s.b);
        ms_out[s.tag] += ms;
        if (launches_out) launches_out[s.tag]++;
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    ctx->prof_spans.clear();
    return PV_OK;
}
int pvacb_keys_copy_blob_to(pvacb_ctx* x, void* dst_device) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    PV_CUDA(cudaMemcpyAsync(dst_device, ctx->d_blob, kBlobBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}
int pvacb_keys_adopt_blob_from(pvacb_ctx* x, const void* src_device) {
    Ctx* ctx = C(x);
    int rc = ensure_blob(ctx);
    if (rc) return rc;
    PV_CUDA(cudaMemcpyAsync(ctx->d_blob, src_device, kBlobBytes, cudaMemcpyDefault, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    return pvacb_keys_adopt_blob(x);
}

// ---- keys
int pvacb_keygen(pvacb_ctx* x, uint64_t tape_state) {
    Ctx* ctx = C(x);
    std::vector<uint64_t> blob;
    int rc = keygen_host(tape_state, blob);
    if (rc) return rc;
    return keys_from_host_blob(ctx, blob.data());
}

int pvacb_keys_import_raw(pvacb_ctx* x, uint64_t canon_tag, const uint8_t h_digest[32], const uint64_t* H, const uint64_t* powg,
                          const uint64_t prf_k[4], const uint64_t* lpn_s) {
    Ctx* ctx = C(x);
    if (!h_digest || !prf_k || !lpn_s) return PV_E_ARG;
    std::vector<uint64_t> blob(kBlobWords, 0);
    blob[0] = canon_tag;
    memcpy(&blob[1], h_digest, 32);
    memcpy(&blob[5], prf_k, 32);
    memcpy(&blob[9], lpn_s, kLpnWords * 8);
    for (int i = 0; i < kB; i++) {
        blob[73 + 2 * i] = powg ? powg[2 * i] : 1;
        blob[73 + 2 * i + 1] = powg ? powg[2 * i + 1] : 0;
    }
    if (H) memcpy(&blob[kBlobHdrWords], H, (size_t)kNBits * kMWords * 8);
    return keys_from_host_blob(ctx, blob.data());
}

int pvacb_keys_export_raw(pvacb_ctx* x, uint64_t* canon_tag, uint8_t h_digest[32], uint64_t* H, uint64_t* powg, uint64_t prf_k[4],
                          uint64_t* lpn_s) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys) return PV_E_NOKEYS;
    const uint64_t* h = ctx->h_hdr.data();
    if (canon_tag) *canon_tag = h[0];
    if (h_digest) memcpy(h_digest, &h[1], 32);
    if (prf_k) memcpy(prf_k, &h[5], 32);
    if (lpn_s) memcpy(lpn_s, &h[9], kLpnWords * 8);
    if (powg) memcpy(powg, &h[73], kB * 16);
    if (H) {
        PV_CUDA(cudaMemcpyAsync(H, ctx->d_blob + kBlobHdrWords, (size_t)kNBits * kMWords * 8, cudaMemcpyDeviceToHost, ctx->stream));
        PV_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return PV_OK;
}

int pvacb_keys_device_blob(pvacb_ctx* x, void** dptr, size_t* bytes) {
    Ctx* ctx = C(x);
    if (!ctx->have_keys) return PV_E_NOKEYS;
    *dptr = ctx->d_blob;
    if (bytes) *bytes = kBlobBytes;
    return PV_OK;
}
int pvacb_keys_alloc_blob(pvacb_ctx* x, void** dptr) {
    Ctx* ctx = C(x);
    int rc = ensure_blob(ctx);
    if (rc) return rc;
    *dptr = ctx->d_blob;
    return PV_OK;
}
int pvacb_keys_adopt_blob(pvacb_ctx* x);
int pvacb_keys_adopt_blob(pvacb_ctx* x) {
    Ctx* ctx = C(x);
    if (!ctx->d_blob) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    ctx->h_hdr.resize(kBlobHdrWords);
    PV_CUDA(cudaMemcpy(ctx->h_hdr.data(), ctx->d_blob, kBlobHdrWords * 8, cudaMemcpyDeviceToHost));
    return derive_key_view(ctx);
}

// ---- ops
int pvacb_enc_value_ex(pvacb_ctx* x, const uint64_t* values, size_t n, uint64_t seed, const uint64_t* tape_states, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || (n && !values)) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    Batch* b = nullptr;
    int rc = op_enc_value(ctx, values, false, n, seed, tape_states, &b);
    *out = reinterpret_cast<pvacb_batch*>(b);
    return rc;
}
int pvacb_enc_value(pvacb_ctx* x, const uint64_t* values, size_t n, uint64_t seed, pvacb_batch** out) {
    return pvacb_enc_value_ex(x, values, n, seed, nullptr, out);
}
int pvacb_enc_value_depth(pvacb_ctx* x, const uint64_t* values, size_t n, int depth_hint, uint64_t seed, const uint64_t* tape_states, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || (n && !values)) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    Batch* b = nullptr;
    int rc = op_enc_value(ctx, values, false, n, seed, tape_states, &b, depth_hint);
    *out = reinterpret_cast<pvacb_batch*>(b);
    return rc;
}
// enc_fp_depth (ops/encrypt.hpp:162): one share per item, 1 BASE layer, 8 + 2 Z2 + 3 Z3 edges; values are field elements (lo, hi)
int pvacb_enc_fp_depth(pvacb_ctx* x, const uint64_t* fp_values, size_t n, int depth_hint, uint64_t seed, const uint64_t* tape_states, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || (n && !fp_values)) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    for (size_t i = 0; i < n; i++) {
        uint64_t lo = fp_values[2 * i], hi = fp_values[2 * i + 1];
        if ((hi >> 63) || (hi == kMask63 && lo == ~0ull)) { ctx->last_error = "enc_fp_depth: value not a canonical field element"; return PV_E_ARG; }
    }
    cudaSetDevice(ctx->device);
    Batch* b = nullptr;
    int rc = op_enc_value(ctx, fp_values, false, n, seed, tape_states, &b, depth_hint, 1);
    *out = reinterpret_cast<pvacb_batch*>(b);
    return rc;
}
// enc_zero_depth draws the mask and both shares exactly like enc_value_depth(0, depth): fp_add(0, mask) == mask
int pvacb_enc_zero_depth(pvacb_ctx* x, size_t n, int depth_hint, uint64_t seed, const uint64_t* tape_states, pvacb_batch** out) {
    std::vector<uint64_t> zeros(n ? n : 1, 0);
    return pvacb_enc_value_depth(x, zeros.data(), n, depth_hint, seed, tape_states, out);
}
int pvacb_plan_noise(int depth_hint, int* z2, int* z3) {
    if (!z2 || !z3) return PV_E_ARG;
    plan_noise_host(nullptr, depth_hint, *z2, *z3);
    return PV_OK;
}
static int binop(pvacb_ctx* x, const pvacb_batch* a, const pvacb_batch* b, int mode, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!a || !b || !out) return PV_E_ARG;
    if (check_owner(ctx, Bt(a)) || check_owner(ctx, Bt(b))) return PV_E_ARG;
    if (Bt(a)->n != Bt(b)->n) { ctx->last_error = "batches differ in length"; return PV_E_SHAPE; }
    cudaSetDevice(ctx->device);
    Batch* o = nullptr;
    int rc = op_ct_add(ctx, Bt(a), Bt(b), mode, &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}
int pvacb_ct_add(pvacb_ctx* x, const pvacb_batch* a, const pvacb_batch* b, pvacb_batch** out) { return binop(x, a, b, 0, out); }
int pvacb_ct_sub(pvacb_ctx* x, const pvacb_batch* a, const pvacb_batch* b, pvacb_batch** out) { return binop(x, a, b, 1, out); }
int pvacb_ct_neg(pvacb_ctx* x, const pvacb_batch* a, pvacb_batch** out) {
    const Fp m1 = fp_neg(fp_one());
    const uint64_t s[2] = {m1.lo, m1.hi};
    return pvacb_ct_scale(x, a, s, out);
}
int pvacb_ct_div_const(pvacb_ctx* x, const pvacb_batch* a, const uint64_t k[2], pvacb_batch** out) {
    if (!k) return PV_E_ARG;
    const Fp inv = fp_inv(fp_from_words(k[0], k[1]));     // one scalar: the 127-squaring Fermat ladder runs once, on the host
    const uint64_t s[2] = {inv.lo, inv.hi};
    return pvacb_ct_scale(x, a, s, out);
}
int pvacb_ct_scale(pvacb_ctx* x, const pvacb_batch* a, const uint64_t s[2], pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!a || !s || !out) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    Batch* o = nullptr;
    int rc = op_ct_scale(ctx, Bt(a), fp_from_words(s[0], s[1]), &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}
int pvacb_ct_mul_ex(pvacb_ctx* x, const pvacb_batch* a, const pvacb_batch* b, uint64_t seed, const uint64_t* tape_states, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!a || !b || !out) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    if (check_owner(ctx, Bt(a)) || check_owner(ctx, Bt(b))) return PV_E_ARG;
    if (Bt(a)->n != Bt(b)->n) { ctx->last_error = "batches differ in length"; return PV_E_SHAPE; }
    cudaSetDevice(ctx->device);
    Batch* o = nullptr;
    int rc = op_ct_mul(ctx, Bt(a), Bt(b), seed, tape_states, &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}
int pvacb_ct_mul(pvacb_ctx* x, const pvacb_batch* a, const pvacb_batch* b, uint64_t seed, pvacb_batch** out) {
    return pvacb_ct_mul_ex(x, a, b, seed, nullptr, out);
}
int pvacb_dec_value(pvacb_ctx* x, const pvacb_batch* c, uint64_t* out) {
    Ctx* ctx = C(x);
    if (!c || !out) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    if (check_owner(ctx, Bt(c))) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    return op_dec_value(ctx, Bt(c), out);
}

// ---- batches
void pvacb_batch_free(pvacb_batch* b) { batch_free(Bt(b)); }
size_t pvacb_batch_count(const pvacb_batch* b) { return b ? (size_t)Bt(b)->n : 0; }
int pvacb_batch_totals(const pvacb_batch* b, uint64_t* nl, uint64_t* ne) {
    if (!b) return PV_E_ARG;
    if (nl) *nl = Bt(b)->nL;
    if (ne) *ne = Bt(b)->nE;
    return PV_OK;
}
size_t pvacb_batch_device_bytes(const pvacb_batch* b) { return b ? Bt(b)->bytes : 0; }

int pvacb_batch_offsets(pvacb_ctx* x, const pvacb_batch* pb, uint32_t* loff, uint32_t* eoff) {
    Ctx* ctx = C(x);
    const Batch* b = Bt(pb);
    if (loff) PV_CUDA(cudaMemcpyAsync(loff, b->loff, (b->n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (eoff) PV_CUDA(cudaMemcpyAsync(eoff, b->eoff, (b->n + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}

int pvacb_batch_export_soa(pvacb_ctx* x, const pvacb_batch* pb, uint32_t* loff, uint32_t* eoff, uint8_t* rule, uint64_t* ztag,
                           uint64_t* nlo, uint64_t* nhi, uint32_t* pa, uint32_t* pbb, uint32_t* lid, uint16_t* idx, uint8_t* ch,
                           uint64_t* w, uint64_t* sigma) {
    Ctx* ctx = C(x);
    const Batch* b = Bt(pb);
    cudaSetDevice(ctx->device);
    auto cp = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        if (!dst || !bytes) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    };
    PV_CUDA(cp(loff, b->loff, (b->n + 1) * 4));
    PV_CUDA(cp(eoff, b->eoff, (b->n + 1) * 4));
    PV_CUDA(cp(rule, b->rule, b->nL));
    PV_CUDA(cp(ztag, b->ztag, b->nL * 8));
    PV_CUDA(cp(nlo, b->nlo, b->nL * 8));
    PV_CUDA(cp(nhi, b->nhi, b->nL * 8));
    PV_CUDA(cp(pa, b->pa, b->nL * 4));
    PV_CUDA(cp(pbb, b->pb, b->nL * 4));
    PV_CUDA(cp(lid, b->lid, b->nE * 4));
    PV_CUDA(cp(idx, b->idx, b->nE * 2));
    PV_CUDA(cp(ch, b->ch, b->nE));
    PV_CUDA(cp(w, b->w, b->nE * 16));
    PV_CUDA(cp(sigma, b->sigma, b->nE * (size_t)kMWords * 8));
    PV_CUDA(cudaStreamSynchronize(ctx->stream));
    return PV_OK;
}

// Same copies on the context's second stream, ordered after everything already queued on the compute stream; returns at
// once. The caller keeps the batch and the (pinned) destination buffers alive until pvacb_export_wait() returns. This is
// what lets the 1.3 MB-per-product device->host read of step k overlap the compute of step k+1.
int pvacb_batch_export_soa_async(pvacb_ctx* x, const pvacb_batch* pb, uint32_t* loff, uint32_t* eoff, uint8_t* rule, uint64_t* ztag,
                                 uint64_t* nlo, uint64_t* nhi, uint32_t* pa, uint32_t* pbb, uint32_t* lid, uint16_t* idx, uint8_t* ch,
                                 uint64_t* w, uint64_t* sigma) {
    Ctx* ctx = C(x);
    const Batch* b = Bt(pb);
    cudaSetDevice(ctx->device);
    cudaEvent_t ready;
    PV_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    PV_CUDA(cudaEventRecord(ready, ctx->stream));
    PV_CUDA(cudaStreamWaitEvent(ctx->stream2, ready, 0));
    cudaEventDestroy(ready);
    auto cp = [&](void* dst, const void* src, size_t bytes) -> cudaError_t {
        if (!dst || !bytes) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream2);
    };
    PV_CUDA(cp(loff, b->loff, (b->n + 1) * 4));
    PV_CUDA(cp(eoff, b->eoff, (b->n + 1) * 4));
    PV_CUDA(cp(rule, b->rule, b->nL));
    PV_CUDA(cp(ztag, b->ztag, b->nL * 8));
    PV_CUDA(cp(nlo, b->nlo, b->nL * 8));
    PV_CUDA(cp(nhi, b->nhi, b->nL * 8));
    PV_CUDA(cp(pa, b->pa, b->nL * 4));
    PV_CUDA(cp(pbb, b->pb, b->nL * 4));
    PV_CUDA(cp(lid, b->lid, b->nE * 4));
    PV_CUDA(cp(idx, b->idx, b->nE * 2));
    PV_CUDA(cp(ch, b->ch, b->nE));
    PV_CUDA(cp(w, b->w, b->nE * 16));
    PV_CUDA(cp(sigma, b->sigma, b->nE * (size_t)kMWords * 8));
    cudaEvent_t done;
    PV_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    PV_CUDA(cudaEventRecord(done, ctx->stream2));
    ctx->export_events.push_back(done);
    return PV_OK;
}

int pvacb_export_wait(pvacb_ctx* x) {
    Ctx* ctx = C(x);
    cudaSetDevice(ctx->device);
    // every outstanding export, whichever stream carries it (the second stream of this device, or a stream of the relay device)
    cudaError_t err = cudaSuccess;
    for (cudaEvent_t e : ctx->export_events) {
        cudaError_t q = cudaEventSynchronize(e);
        if (err == cudaSuccess) err = q;
        cudaEventDestroy(e);
    }
    ctx->export_events.clear();
    PV_CUDA(err);
    return PV_OK;
}

int pvacb_export_wait_one(pvacb_ctx* x) {
    Ctx* ctx = C(x);
    cudaSetDevice(ctx->device);
    if (ctx->export_events.empty()) return PV_OK;
    cudaEvent_t e = ctx->export_events.front();
    ctx->export_events.pop_front();
    cudaError_t err = cudaEventSynchronize(e);
    cudaEventDestroy(e);
    PV_CUDA(err);
    return PV_OK;
}

int pvacb_batch_import_soa(pvacb_ctx* x, size_t n, const uint32_t* loff, const uint32_t* eoff, const uint8_t* rule, const uint64_t* ztag,
                           const uint64_t* nlo, const uint64_t* nhi, const uint32_t* pa, const uint32_t* pbb, const uint32_t* lid,
                           const uint16_t* idx, const uint8_t* ch, const uint64_t* w, const uint64_t* sigma, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || !loff || !eoff) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    uint64_t nL = loff[n], nE = eoff[n];
    if (loff[0] != 0 || eoff[0] != 0) return PV_E_FORMAT;
    for (size_t i = 0; i < n; i++)
        if (loff[i + 1] < loff[i] || eoff[i + 1] < eoff[i]) { ctx->last_error = "offsets not monotone"; return PV_E_FORMAT; }
    if ((nL && (!rule || !ztag || !nlo || !nhi)) || (nE && (!lid || !idx || !ch || !w))) {
        ctx->last_error = "import_soa: a layer / edge array is NULL although the batch has layers / edges (only pa, pb and sigma may be omitted)";
        return PV_E_ARG;
    }
    Batch* b = nullptr;
    int rc = batch_alloc(ctx, n, nL, nE, &b);
    if (rc) return rc;
    cudaError_t ce = cudaSuccess;
    auto cp = [&](void* dst, const void* src, size_t bytes) {
        if (!bytes || ce != cudaSuccess) return;
        ce = src ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream) : cudaMemsetAsync(dst, 0, bytes, ctx->stream);
    };
    cp(b->loff, loff, (n + 1) * 4); cp(b->eoff, eoff, (n + 1) * 4);
    cp(b->rule, rule, nL); cp(b->ztag, ztag, nL * 8); cp(b->nlo, nlo, nL * 8); cp(b->nhi, nhi, nL * 8); cp(b->pa, pa, nL * 4); cp(b->pb, pbb, nL * 4);
    cp(b->lid, lid, nE * 4); cp(b->idx, idx, nE * 2); cp(b->ch, ch, nE); cp(b->w, w, nE * 16); cp(b->sigma, sigma, nE * (size_t)kMWords * 8);
    if (ce != cudaSuccess) { batch_free(b); ctx->last_error = std::string("import_soa: ") + cudaGetErrorString(ce); return PV_E_CUDA; }
    // validation the kernels rely on (the reference would index out of bounds instead): on the device, one launch + one
    // synchronisation, which also completes the copies above
    if ((rc = batch_validate(ctx, b))) { batch_free(b); return rc; }
    *out = reinterpret_cast<pvacb_batch*>(b);
    return PV_OK;
}

// ---- wire format (tests/bounty2_test.cpp:17-143)
static const uint32_t kMagicCT = 0x66699666u, kWireVer = 1;

struct HostSoA {
    std::vector<uint32_t> loff, eoff, pa, pb, lid;
    std::vector<uint8_t> rule, ch;
    std::vector<uint64_t> ztag, nlo, nhi, w, sigma;
    std::vector<uint16_t> idx;
};

static int to_host(pvacb_ctx* x, const pvacb_batch* pb, HostSoA& h, bool with_sigma) {
    const Batch* b = Bt(pb);
    h.loff.resize(b->n + 1); h.eoff.resize(b->n + 1);
    h.rule.resize(b->nL); h.ztag.resize(b->nL); h.nlo.resize(b->nL); h.nhi.resize(b->nL); h.pa.resize(b->nL); h.pb.resize(b->nL);
    h.lid.resize(b->nE); h.idx.resize(b->nE); h.ch.resize(b->nE); h.w.resize(b->nE * 2);
    if (with_sigma) h.sigma.resize(b->nE * (size_t)kMWords);
    return pvacb_batch_export_soa(x, pb, h.loff.data(), h.eoff.data(), h.rule.data(), h.ztag.data(), h.nlo.data(), h.nhi.data(), h.pa.data(),
                                  h.pb.data(), h.lid.data(), h.idx.data(), h.ch.data(), h.w.data(), with_sigma ? h.sigma.data() : nullptr);
}

int pvacb_batch_wire_size(pvacb_ctx* x, const pvacb_batch* pb, size_t* bytes) {
    const Batch* b = Bt(pb);
    HostSoA h;
    h.rule.resize(b->nL);
    Ctx* ctx = C(x);
    PV_CUDA(cudaMemcpy(h.rule.data(), b->rule, b->nL, cudaMemcpyDeviceToHost));
    size_t s = 16 + b->n * 8;
    for (uint8_t r : h.rule) s += 1 + (r == 1 ? 8 : 24);
    s += b->nE * (size_t)(4 + 2 + 1 + 1 + 16 + 4 + kMWords * 8);
    *bytes = s;
    return PV_OK;
}

int pvacb_batch_export_wire(pvacb_ctx* x, const pvacb_batch* pb, void* buf, size_t cap, size_t* written) {
    Ctx* ctx = C(x);
    const Batch* b = Bt(pb);
    HostSoA h;
    int rc = to_host(x, pb, h, true);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)buf;
    size_t off = 0;
    auto put = [&](const void* src, size_t nbytes) -> bool {
        if (off + nbytes > cap) return false;
        memcpy(p + off, src, nbytes);
        off += nbytes;
        return true;
    };
    uint64_t cnt = b->n;
    bool ok = put(&kMagicCT, 4) && put(&kWireVer, 4) && put(&cnt, 8);
    for (uint64_t i = 0; ok && i < b->n; i++) {
        uint32_t nl = h.loff[i + 1] - h.loff[i], ne = h.eoff[i + 1] - h.eoff[i];
        ok = put(&nl, 4) && put(&ne, 4);
        for (uint32_t k = h.loff[i]; ok && k < h.loff[i + 1]; k++) {
            ok = put(&h.rule[k], 1);
            if (h.rule[k] == 0) ok = ok && put(&h.ztag[k], 8) && put(&h.nlo[k], 8) && put(&h.nhi[k], 8);
            else ok = ok && put(&h.pa[k], 4) && put(&h.pb[k], 4);
        }
        for (uint32_t e = h.eoff[i]; ok && e < h.eoff[i + 1]; e++) {
            uint8_t zero = 0;
            uint32_t nbits = kMBits;
            ok = put(&h.lid[e], 4) && put(&h.idx[e], 2) && put(&h.ch[e], 1) && put(&zero, 1) && put(&h.w[2 * e], 16) && put(&nbits, 4) &&
                 put(&h.sigma[(size_t)e * kMWords], kMWords * 8);
        }
    }
    if (!ok) { ctx->last_error = "wire buffer too small"; return PV_E_ARG; }
    if (written) *written = off;
    return PV_OK;
}

int pvacb_batch_import_wire(pvacb_ctx* x, const void* buf, size_t bytes, pvacb_batch** out) {
    Ctx* ctx = C(x);
    const uint8_t* p = (const uint8_t*)buf;
    size_t off = 0;
    auto get = [&](void* dst, size_t nbytes) -> bool {
        if (off + nbytes > bytes) return false;
        memcpy(dst, p + off, nbytes);
        off += nbytes;
        return true;
    };
    auto truncated = [&]() { ctx->last_error = "ciphertext file ends inside a record"; return (int)PV_E_FORMAT; };
    uint32_t magic = 0, ver = 0;
    uint64_t cnt = 0;
    if (!get(&magic, 4) || !get(&ver, 4) || !get(&cnt, 8) || magic != kMagicCT || ver != kWireVer) { ctx->last_error = "bad CT header"; return PV_E_FORMAT; }
    HostSoA h;
    h.loff.push_back(0); h.eoff.push_back(0);
    for (uint64_t i = 0; i < cnt; i++) {
        uint32_t nl = 0, ne = 0;
        if (!get(&nl, 4) || !get(&ne, 4)) return truncated();
        for (uint32_t k = 0; k < nl; k++) {
            uint8_t r = 0;
            uint64_t a = 0, b = 0, c = 0;
            uint32_t pa = 0, pb = 0;
            if (!get(&r, 1)) return truncated();
            if (r == 1) { if (!get(&pa, 4) || !get(&pb, 4)) return truncated(); }
            else if (!get(&a, 8) || !get(&b, 8) || !get(&c, 8)) return truncated();
            h.rule.push_back(r); h.ztag.push_back(a); h.nlo.push_back(b); h.nhi.push_back(c); h.pa.push_back(pa); h.pb.push_back(pb);
        }
        for (uint32_t e = 0; e < ne; e++) {
            uint32_t lid = 0, nbits = 0;
            uint16_t idx = 0;
            uint8_t ch = 0, pad = 0;
            uint64_t w[2];
            if (!get(&lid, 4) || !get(&idx, 2) || !get(&ch, 1) || !get(&pad, 1) || !get(w, 16) || !get(&nbits, 4)) return truncated();
            if (nbits != kMBits) { ctx->last_error = "sigma length is not m_bits"; return PV_E_FORMAT; }
            if (pad != 0) { ctx->last_error = "reserved edge byte is not zero"; return PV_E_FORMAT; }   // putEdge always writes 0 there
            size_t so = h.sigma.size();
            h.sigma.resize(so + kMWords);
            if (!get(&h.sigma[so], kMWords * 8)) return truncated();
            h.lid.push_back(lid); h.idx.push_back(idx); h.ch.push_back(ch); h.w.push_back(w[0]); h.w.push_back(w[1]);
        }
        h.loff.push_back((uint32_t)h.rule.size());
        h.eoff.push_back((uint32_t)h.lid.size());
    }
    if (off != bytes) { ctx->last_error = "trailing bytes after the last ciphertext"; return PV_E_FORMAT; }
    return pvacb_batch_import_soa(x, cnt, h.loff.data(), h.eoff.data(), h.rule.data(), h.ztag.data(), h.nlo.data(), h.nhi.data(), h.pa.data(),
                                  h.pb.data(), h.lid.data(), h.idx.data(), h.ch.data(), h.w.data(), h.sigma.data(), out);
}

}  // extern "C"
