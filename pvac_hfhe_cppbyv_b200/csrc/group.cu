// Multi-GPU group behind the C ABI: N contexts of one process, one per device, one host thread per device while a
// call runs. Ciphertexts are independent, so a batch is split into contiguous index ranges (item i -> member floor(i G / n)); every
// item keeps the RNG stream of its GLOBAL index (pvacb_set_item_base), so the bytes do not depend on how many GPUs share the
// batch; the keys are replicated once with peer copies over NVLink; there is no collective in the steady state.
#include "engine.h"
#include "../../include/pvacb.h"

#include <chrono>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <algorithm>

using namespace pvacb;

struct pvacb_group {
    std::vector<pvacb_ctx*> ctx;
    std::vector<int> dev;
    std::string last_error;
};
struct pvacb_gbatch {
    pvacb_group* g;
    std::vector<pvacb_batch*> part;      // one per member (possibly holding 0 items)
    std::vector<uint64_t> first;         // global index of the member's first item
    uint64_t n = 0;
};

static inline Ctx* CC(pvacb_ctx* x) { return reinterpret_cast<Ctx*>(x); }

// contiguous ranges: member r owns [ceil(r n / G), ceil((r+1) n / G))
static void partition(uint64_t n, int G, std::vector<uint64_t>& first, std::vector<uint64_t>& count) {
    first.resize(G); count.resize(G);
    for (int r = 0; r < G; r++) {
        const uint64_t a = ((uint64_t)r * n + G - 1) / G, b = ((uint64_t)(r + 1) * n + G - 1) / G;
        first[r] = a; count[r] = b - a;
    }
}

// fn(member) on one thread per member; the first non-zero status wins and its message becomes the group's
static int parallel(pvacb_group* g, const std::function<int(int)>& fn) {
    const int G = (int)g->ctx.size();
    std::vector<int> rc(G, 0);
    if (G == 1) rc[0] = fn(0);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < G; k++) th.emplace_back([&, k]() { rc[k] = fn(k); });
        for (auto& t : th) t.join();
    }
    for (int k = 0; k < G; k++)
        if (rc[k]) { g->last_error = "device " + std::to_string(g->dev[k]) + ": " + pvacb_last_error(g->ctx[k]); return rc[k]; }
    return PV_OK;
}

static pvacb_gbatch* gb_new(pvacb_group* g, uint64_t n) {
    pvacb_gbatch* o = new pvacb_gbatch();
    o->g = g; o->n = n;
    o->part.assign(g->ctx.size(), nullptr);
    std::vector<uint64_t> cnt;
    partition(n, (int)g->ctx.size(), o->first, cnt);
    return o;
}

extern "C" {

int pvacb_group_create(const int* devices, int n, pvacb_group** out) {
    if (!devices || n < 1 || !out) return PV_E_ARG;
    for (int a = 0; a < n; a++)
        for (int b = a + 1; b < n; b++)
            if (devices[a] == devices[b]) return PV_E_ARG;
    pvacb_group* g = new pvacb_group();
    for (int k = 0; k < n; k++) {
        pvacb_ctx* c = nullptr;
        int rc = pvacb_ctx_create(devices[k], &c);
        if (rc) { pvacb_group_destroy(g); return rc; }
        g->ctx.push_back(c);
        g->dev.push_back(devices[k]);
    }
    // one tape for the whole group: member 0's fresh key
    for (int k = 1; k < n; k++) memcpy(CC(g->ctx[k])->tape_key, CC(g->ctx[0])->tape_key, 32);
    *out = g;
    return PV_OK;
}

void pvacb_group_destroy(pvacb_group* g) {
    if (!g) return;
    for (pvacb_ctx* c : g->ctx) pvacb_ctx_destroy(c);
    delete g;
}
int pvacb_group_size(const pvacb_group* g) { return g ? (int)g->ctx.size() : 0; }
pvacb_ctx* pvacb_group_ctx(pvacb_group* g, int k) { return (g && k >= 0 && k < (int)g->ctx.size()) ? g->ctx[k] : nullptr; }
const char* pvacb_group_last_error(const pvacb_group* g) { return g ? g->last_error.c_str() : "null group"; }

// member 0's keys -> every other member: one peer copy of the 16.8 MB blob each (NVLink when peer access exists, staged by the
// driver otherwise), plus the run-time Params
int pvacb_group_replicate_keys(pvacb_group* g) {
    if (!g) return PV_E_ARG;
    Ctx* c0 = CC(g->ctx[0]);
    if (!c0->have_keys) { g->last_error = "member 0 has no keys"; return PV_E_NOKEYS; }
    pvacb_params prm;
    pvacb_get_params(g->ctx[0], &prm);
    return parallel(g, [&](int k) -> int {
        if (k == 0) return PV_OK;
        Ctx* ctx = CC(g->ctx[k]);
        cudaSetDevice(ctx->device);
        void* dst = nullptr;
        int rc = pvacb_keys_alloc_blob(g->ctx[k], &dst);
        if (rc) return rc;
        // no cudaDeviceEnablePeerAccess: the copy is 16.8 MB once and the driver routes it (NVLink where there is a path); nothing in the
        // steady state touches another GPU's memory, so an accidental cross-device pointer faults instead of silently crawling over NVLink
        PV_CUDA(cudaMemcpyPeerAsync(dst, ctx->device, c0->d_blob, c0->device, kBlobBytes, ctx->stream));
        PV_CUDA(cudaStreamSynchronize(ctx->stream));
        if ((rc = pvacb_keys_adopt_blob(g->ctx[k]))) return rc;
        ctx->have_sk = c0->have_sk;
        return pvacb_set_params(g->ctx[k], &prm);
    });
}

int pvacb_group_keygen_params(pvacb_group* g, const pvacb_params* prm, const uint8_t seed[32]) {
    if (!g) return PV_E_ARG;
    int rc = pvacb_keygen_params(g->ctx[0], prm, seed);
    if (rc) { g->last_error = pvacb_last_error(g->ctx[0]); return rc; }
    return pvacb_group_replicate_keys(g);
}
int pvacb_group_keys_import_file(pvacb_group* g, const char* pk_path, const char* sk_path) {
    if (!g) return PV_E_ARG;
    int rc = pvacb_keys_import_file(g->ctx[0], pk_path, sk_path);
    if (rc) { g->last_error = pvacb_last_error(g->ctx[0]); return rc; }
    return pvacb_group_replicate_keys(g);
}

int pvacb_group_set_tape(pvacb_group* g, int kind, const uint8_t key[32]) {
    if (!g) return PV_E_ARG;
    int rc = pvacb_set_tape(g->ctx[0], kind, key);
    if (rc) { g->last_error = pvacb_last_error(g->ctx[0]); return rc; }
    for (size_t k = 1; k < g->ctx.size(); k++) {
        CC(g->ctx[k])->tape_kind = kind;
        memcpy(CC(g->ctx[k])->tape_key, CC(g->ctx[0])->tape_key, 32);
    }
    return PV_OK;
}

// ---- sharded batches
void pvacb_group_batch_free(pvacb_gbatch* b) {
    if (!b) return;
    for (pvacb_batch* p : b->part) pvacb_batch_free(p);
    delete b;
}
size_t pvacb_group_batch_count(const pvacb_gbatch* b) { return b ? (size_t)b->n : 0; }
pvacb_batch* pvacb_group_batch_part(pvacb_gbatch* b, int k) { return (b && k >= 0 && k < (int)b->part.size()) ? b->part[k] : nullptr; }

// Cipher enc_value(pk, sk, v) for values[0..n): member k encrypts its index range under the GLOBAL item streams of batch_seed
int pvacb_group_enc_value(pvacb_group* g, const uint64_t* values, size_t n, uint64_t batch_seed, pvacb_gbatch** out) {
    if (!g || !out || (n && !values)) return PV_E_ARG;
    pvacb_gbatch* o = gb_new(g, n);
    std::vector<uint64_t> first, cnt;
    partition(n, (int)g->ctx.size(), first, cnt);
    int rc = parallel(g, [&](int k) -> int {
        pvacb_set_item_base(g->ctx[k], first[k]);
        int r = pvacb_enc_value_ex(g->ctx[k], values + first[k], cnt[k], batch_seed, nullptr, &o->part[k]);
        pvacb_set_item_base(g->ctx[k], 0);
        return r;
    });
    if (rc) { pvacb_group_batch_free(o); return rc; }
    *out = o;
    return PV_OK;
}

static int group_binop(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, int op, uint64_t seed, pvacb_gbatch** out) {
    if (!g || !a || !b || !out || a->g != g || b->g != g) return PV_E_ARG;
    if (a->n != b->n) { g->last_error = "batches differ in length"; return PV_E_SHAPE; }
    pvacb_gbatch* o = gb_new(g, a->n);
    int rc = parallel(g, [&](int k) -> int {
        if (op == 0) return pvacb_ct_add(g->ctx[k], a->part[k], b->part[k], &o->part[k]);
        if (op == 1) return pvacb_ct_sub(g->ctx[k], a->part[k], b->part[k], &o->part[k]);
        pvacb_set_item_base(g->ctx[k], a->first[k]);
        int r = pvacb_ct_mul_ex(g->ctx[k], a->part[k], b->part[k], seed, nullptr, &o->part[k]);
        pvacb_set_item_base(g->ctx[k], 0);
        return r;
    });
    if (rc) { pvacb_group_batch_free(o); return rc; }
    *out = o;
    return PV_OK;
}
int pvacb_group_ct_add(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, pvacb_gbatch** out) { return group_binop(g, a, b, 0, 0, out); }
int pvacb_group_ct_sub(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, pvacb_gbatch** out) { return group_binop(g, a, b, 1, 0, out); }
int pvacb_group_ct_mul(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, uint64_t batch_seed, pvacb_gbatch** out) {
    return group_binop(g, a, b, 2, batch_seed, out);
}

// Fp dec_value(pk, sk, C): out = n x (lo, hi) in global order (every member writes its own range of the host array)
int pvacb_group_dec_value(pvacb_group* g, const pvacb_gbatch* c, uint64_t* out) {
    if (!g || !c || !out || c->g != g) return PV_E_ARG;
    return parallel(g, [&](int k) -> int {
        if (pvacb_batch_count(c->part[k]) == 0) return PV_OK;
        return pvacb_dec_value(g->ctx[k], c->part[k], out + 2 * c->first[k]);
    });
}
int pvacb_group_commit_ct(pvacb_group* g, const pvacb_gbatch* c, uint8_t* out) {
    if (!g || !c || !out || c->g != g) return PV_E_ARG;
    return parallel(g, [&](int k) -> int {
        if (pvacb_batch_count(c->part[k]) == 0) return PV_OK;
        return pvacb_commit_ct(g->ctx[k], c->part[k], out + 32 * c->first[k]);
    });
}

// every member exports the image of its part into host_bufs[k] (pinned, caps[k] bytes) at the same time; returns when all are done
int pvacb_group_export_blobs(pvacb_group* g, const pvacb_gbatch* c, void* const* host_bufs, const size_t* caps) {
    if (!g || !c || !host_bufs || !caps || c->g != g) return PV_E_ARG;
    return parallel(g, [&](int k) -> int {
        int rc = pvacb_batch_export_blob_async(g->ctx[k], c->part[k], host_bufs[k], caps[k]);
        if (rc) return rc;
        return pvacb_export_wait(g->ctx[k]);
    });
}

// Measures what the members can export to the host AT THE SAME TIME (256 MB images, a few copies each, all members between two
// barriers): first with every GPU on its own host link, then -- if the links turn out unequal -- with the slower half of the
// members relayed through the faster half over NVLink (pvacb_set_export_relay). Keeps whichever moved more bytes per second.
int pvacb_group_tune_export(pvacb_group* g, double* direct_gbs, double* relay_gbs) {
    if (!g) return PV_E_ARG;
    const int G = (int)g->ctx.size();
    const size_t n_items = 6144;                               // x 40 edges x ~1 KiB = 256 MB per member
    std::vector<pvacb_batch*> probe(G, nullptr);
    std::vector<void*> host(G, nullptr);
    std::vector<size_t> bytes(G, 0);
    int rc = parallel(g, [&](int k) -> int {
        int r = pvacb_batch_synthetic(g->ctx[k], n_items, 20, 77 + k, &probe[k]);
        if (r) return r;
        uint64_t by = 0;
        pvacb_batch_blob_info(probe[k], nullptr, nullptr, nullptr, &by);
        bytes[k] = by;
        cudaSetDevice(g->dev[k]);
        if (cudaHostAlloc(&host[k], by, cudaHostAllocDefault) != cudaSuccess) { CC(g->ctx[k])->last_error = "pinned allocation failed"; return PV_E_OOM; }
        return PV_OK;
    });
    auto cleanup = [&]() {
        for (int k = 0; k < G; k++) { if (probe[k]) pvacb_batch_free(probe[k]); if (host[k]) { cudaSetDevice(g->dev[k]); cudaFreeHost(host[k]); } }
    };
    if (rc) { cleanup(); return rc; }
    std::vector<double> rate(G, 0.0);
    auto measure = [&](double& aggregate) -> int {
        std::atomic<int> arrived{0}, failed{0};
        std::vector<double> t0(G), t1(G);
        const int reps = 6;
        int r = parallel(g, [&](int k) -> int {
            int q = 0;
            for (int w = 0; w < 2 && !q; w++) {                                                   // warm-up (staging buffer, peer mappings, first-touch of the route)
                q = pvacb_batch_export_blob_async(g->ctx[k], probe[k], host[k], bytes[k]);
                if (!q) q = pvacb_export_wait(g->ctx[k]);
            }
            if (q) failed.store(1);
            arrived.fetch_add(1);                                                                 // a member whose warm-up failed still arrives: nobody spins forever
            while (arrived.load() < G) std::this_thread::yield();
            if (failed.load()) return q;                                                          // the failing member reports; the others skip the measurement
            t0[k] = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
            for (int i = 0; i < reps; i++)
                if ((q = pvacb_batch_export_blob_async(g->ctx[k], probe[k], host[k], bytes[k]))) return q;
            q = pvacb_export_wait(g->ctx[k]);
            t1[k] = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
            return q;
        });
        if (r) return r;
        double a = 1e300, b = 0, tot = 0;
        for (int k = 0; k < G; k++) { a = std::min(a, t0[k]); b = std::max(b, t1[k]); tot += (double)bytes[k] * reps; rate[k] = (double)bytes[k] * reps / (t1[k] - t0[k]) / 1e9; }
        aggregate = tot / (b - a) / 1e9;
        return PV_OK;
    };
    for (int k = 0; k < G; k++) pvacb_set_export_relay(g->ctx[k], -1);
    double A0 = 0, A1 = 0;
    if ((rc = measure(A0))) { cleanup(); return rc; }
    if (G >= 2) {
        std::vector<int> order(G);
        for (int k = 0; k < G; k++) order[k] = k;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return rate[x] < rate[y]; });
        if (rate[order[0]] < 0.8 * rate[order[G - 1]]) {
            const int half = G / 2;
            bool ok = true;
            for (int i = 0; i < half && ok; i++) ok = pvacb_set_export_relay(g->ctx[order[i]], g->dev[order[G - half + i]]) == PV_OK;
            if (ok && measure(A1) == PV_OK && A1 > 1.05 * A0) { /* keep the relays */ }
            else for (int k = 0; k < G; k++) pvacb_set_export_relay(g->ctx[k], -1);
        }
    }
    cleanup();
    if (direct_gbs) *direct_gbs = A0;
    if (relay_gbs) *relay_gbs = A1;
    return PV_OK;
}

}  // extern "C"
