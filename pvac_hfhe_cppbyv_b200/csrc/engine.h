// Internal host-side declarations of the pvacb engine (not part of the C ABI; see include/pvacb.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <vector>
#include <deque>
#include <string>

#include "common.cuh"
#include "fp127.cuh"
#include "aes256.cuh"
#include "prf_core.cuh"

namespace pvacb {

// ---- device-resident key material. One flat "key blob" (what is replicated across GPUs) + derived tables.
// blob layout (little-endian u64 words):
//   [0]            canon_tag
//   [1..4]         H_digest (32 bytes)
//   [5..8]         prf_k[4]
//   [9..72]        lpn_s[64]
//   [73..746]      powg_B[337] as (lo,hi)
//   [747]          reserved (0)
//   [748..749]     omega_B (lo,hi): computed by keygen and stored in the reference's pk files, read by no operation
//   [750..751]     reserved (0)
//   [752..]        H : 16384 columns x 128 words
constexpr size_t kBlobHdrWords = 752;
constexpr size_t kBlobWords = kBlobHdrWords + (size_t)kNBits * kMWords;
constexpr size_t kBlobBytes = kBlobWords * 8;


struct KeyView {             // passed by value to kernels (pointers into the device blob)
    uint64_t canon_tag;
    uint32_t kd_mid[8];      // SHA-256 state after block 0 of derive_aes_key's message (crypto/lpn.hpp:166-192)
    uint64_t digest3;        // LE64(H_digest[24..31]) = first word of block 1
    const uint64_t* H;       // [16384][128]
    const Fp* powg;          // [337]
    const uint32_t* T0;      // [256] AES T-table
    const uint8_t* sbox;     // [256]
};

enum PrfMode : int { PRF_FAITHFUL = 0, PRF_LIVE = 1 };

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;
    std::deque<cudaEvent_t> export_events;   // one per pvacb_batch_export_soa_async still to be waited for, in issue order
    int sm_count = 148;
    bool have_keys = false;
    bool have_sk = false;            // false after a pk-only import: enc_* / dec_* refuse
    uint64_t* d_blob = nullptr;      // kBlobBytes
    AesTables* d_aes = nullptr;
    uint64_t* d_primes = nullptr;    // libstdc++ bucket-count table
    int n_primes = 0;
    unsigned long long* d_work = nullptr;   // work-queue counter of the persistent sigma kernel
    uint16_t* d_ubk_perm = nullptr;         // UBK: public permutation of the sigma bits (crypto/matrix.hpp:95-164), derived from canon_tag
    std::vector<uint16_t> h_ubk_perm;
    uint32_t* h_mail = nullptr;             // mapped pinned "mailbox": kernels drop small results here (see SmallRead)
    uint32_t* d_mail = nullptr;             // device alias of h_mail
    KeyView kv{};
    LpnMasks lpn_m{};                // the LPN secret as per-stream-word masks (prf_core.cuh)
    std::vector<uint64_t> h_hdr;     // host copy of the blob header
    int prf_mode = PRF_FAITHFUL;
    // the run-time part of Params (core/types.hpp:36-70); everything else is compiled in (common.cuh) and checked by pvacb_keygen_params
    double noise_entropy_bits = 120.0, tuple2_fraction = 0.55, depth_slope_bits = 16.0;
    uint32_t edge_budget = kEdgeBudget;
    int lpn_t = kLpnT;
    double recrypt_lo = 0.48, recrypt_hi = 0.52;
    int recrypt_rounds = 8;
    // RNG tape (common.cuh): ChaCha20 under a key from the OS unless the caller chose otherwise (pvacb_set_tape)
    int tape_kind = TAPE_CHACHA20;
    uint32_t tape_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint64_t item_base = 0;                 // global index of item 0 of the next calls (pvacb_set_item_base)
    uint64_t seed_counter = 0;              // pvacb_fresh_seed
    uint64_t* d_tape_words = nullptr;       // TAPE_WORDS: [tape_words_items][tape_words_per_item]
    uint64_t tape_words_items = 0, tape_words_per_item = 0;
    // test hook: keystream word `prf_patch_word` of every PRF core is OR-ed with `prf_patch_or` (forces AesCtr256::bounded's rejection branch)
    uint64_t prf_patch_word = ~0ull, prf_patch_or = 0;
    std::string last_error;
    // statistics of the last call (for bench.py)
    uint64_t stat_kernel_launches = 0;
    uint64_t stat_aes_blocks = 0;
    uint64_t stat_sigma_edges = 0;
    uint64_t stat_rare_prf_cores = 0;   // launches of the serial PRF slow path (bounded() rejections)
    int profile = 0;                 // when set, the dominant kernels are bracketed by CUDA events on ctx->stream
    struct ProfSpan { int tag; cudaEvent_t a, b; };
    std::vector<ProfSpan> prof_spans;
    struct Relay {                   // exports leave through another GPU's host link (pvacb_set_export_relay)
        int device = -1;
        void* stage[2] = {nullptr, nullptr};
        size_t cap[2] = {0, 0};
        cudaStream_t stream[2] = {nullptr, nullptr};
        unsigned next = 0;
    } relay;
    bool lpn_attr_set = false;
    bool mul_force_device_sort = false, mul_force_global_table = false;   // ct_mul planning: A/B switches for tests (pvacb_debug_set)
    bool sigma_test_shape = false;          // sigma kernel shape without spare PRG candidates (exercises the in-kernel continuation)
    std::vector<const void*> configured_kernels;   // kernels whose per-device function attributes this context has set
};

// profiling tags (pvacb_profile_collect)
enum : int { PROF_PRF_LPN = 0, PROF_MUL_PAIRS = 1, PROF_SIGMA = 2, PROF_CONCAT = 3, PROF_DEC_EDGES = 4, PROF_COMMIT = 5, PROF_COMPACT = 6, PROF_NTAGS = 8 };
struct ProfScope {
    Ctx* ctx; cudaEvent_t a = nullptr, b = nullptr; int tag;
    ProfScope(Ctx* c, int t) : ctx(c), tag(t) {
        if (ctx->profile) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream); }
    }
    ~ProfScope() {
        if (a) { cudaEventRecord(b, ctx->stream); ctx->prof_spans.push_back({tag, a, b}); }
    }
};

// device ciphertext batch, structure-of-arrays
struct Batch {
    Ctx* ctx = nullptr;
    uint64_t n = 0;          // ciphertexts
    uint64_t nL = 0, nE = 0; // total layers / edges
    uint64_t nL_alloc = 0;   // layer count the allocation was laid out for (compact_layers shrinks nL in place)
    void* base = nullptr;    // one allocation
    size_t bytes = 0;
    uint32_t* loff = nullptr;  // [n+1]
    uint32_t* eoff = nullptr;  // [n+1]
    uint8_t* rule = nullptr;   // [nL]
    uint64_t* ztag = nullptr;  // [nL]
    uint64_t* nlo = nullptr;
    uint64_t* nhi = nullptr;
    uint32_t* pa = nullptr;
    uint32_t* pb = nullptr;
    uint32_t* lid = nullptr;   // [nE] layer id relative to the ciphertext
    uint16_t* idx = nullptr;   // [nE]
    uint8_t* ch = nullptr;     // [nE]
    Fp* w = nullptr;           // [nE] 16-byte aligned
    uint64_t* sigma = nullptr; // [nE][128]
};

// status codes (mirrored in include/pvacb.h)
enum : int {
    PV_OK = 0,
    PV_E_ARG = 1,
    PV_E_CUDA = 2,
    PV_E_OOM = 3,
    PV_E_NOKEYS = 4,
    PV_E_EDGE_BUDGET = 5,   // result would exceed Params::edge_budget (reference: guard_budget -> compact_edges)
    PV_E_LAYER_GRAPH = 6,   // parent out of range / cycle (reference: std::abort in layer_R_cached)
    PV_E_RARE_PATH = 7,     // a 2^-61-probability rejection-sampling branch was hit (not supported on device)
    PV_E_DUP_EDGE = 8,      // input holds two edges with equal (layer, idx, sign); run compaction first
    PV_E_FORMAT = 9,
    PV_E_SHAPE = 10,
};

#define PV_CUDA(expr)                                                                       \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            ctx->last_error = std::string(#expr) + ": " + cudaGetErrorString(_e);           \
            return (_e == cudaErrorMemoryAllocation) ? PV_E_OOM : PV_E_CUDA;                \
        }                                                                                   \
    } while (0)

// Small device->host reads (counts, totals, error flags) that the host needs before it can size the next allocation.
// They do NOT use cudaMemcpy: a 4-byte copy shares the device->host copy engine with bulk exports running on the second
// stream (pvacb_batch_export_soa_async) and would sit behind hundreds of MB. One single-thread kernel writes the words
// into mapped pinned memory instead, then the compute stream is synchronised.
struct SmallRead {
    static constexpr int kMaxItems = 12, kMaxWords = 48;
    const uint32_t* src[kMaxItems];
    uint32_t words[kMaxItems];
    void* dst[kMaxItems];
    int n = 0;
    void add(void* host_dst, const void* dev_src, size_t bytes) {
        src[n] = static_cast<const uint32_t*>(dev_src); words[n] = (uint32_t)(bytes / 4); dst[n] = host_dst; n++;
    }
};
int read_small_sync(Ctx* ctx, const SmallRead& r);   // enqueue + cudaStreamSynchronize(ctx->stream) + scatter to the host dsts

// a batch lives in the memory of the context (device) that made it
inline int check_owner(Ctx* ctx, const Batch* b) {
    if (b && b->ctx != ctx) { ctx->last_error = "batch belongs to another context (device)"; return PV_E_ARG; }
    return PV_OK;
}
int batch_alloc(Ctx* ctx, uint64_t n, uint64_t nL, uint64_t nE, Batch** out);
void batch_free(Batch* b);
int batch_clone(Ctx* ctx, const Batch* src, Batch** out);   // field-by-field device copy
int batch_validate(Ctx* ctx, const Batch* b);                // boundary.cu: ranges / canonical weights of a batch that came from outside
int keys_from_host_blob(Ctx* ctx, const uint64_t* blob);     // engine.cu: upload a host key blob and derive the key view

// scratch allocator: stream-ordered
int dev_alloc(Ctx* ctx, void** p, size_t bytes);
void dev_free(Ctx* ctx, void* p);

// temporaries of one host-side op: freed (stream-ordered) when the scope ends, on every return path
struct Scratch {
    Ctx* ctx;
    std::vector<void*> ptrs;
    explicit Scratch(Ctx* c) : ctx(c) {}
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() { for (void* p : ptrs) dev_free(ctx, p); }
    template <class T>
    int alloc(T*& p, size_t bytes) {
        void* v = nullptr;
        int rc = dev_alloc(ctx, &v, bytes);
        if (rc) return rc;
        ptrs.push_back(v);
        p = static_cast<T*>(v);
        return PV_OK;
    }
};

// ---- PRF (prf.cu): out[j] = prf_R(seed_j) (family 0) or prf_R_noise(seed_j) (family 1); inactive jobs give 0
int prf_run(Ctx* ctx, uint64_t njobs, const uint64_t* d_ztag, const uint64_t* d_nlo, const uint64_t* d_nhi,
            const uint8_t* d_flags /*bit0 family, bit1 active*/, Fp* d_out, uint64_t* d_ybits_out /*optional debug: [njobs*3][rows/64]*/);

// ---- sigma (sigma.cu): one sigma_from_H per job, written to row out_row[job] (1 KiB rows) of `out`
struct SigmaJobs {
    uint64_t n = 0;
    const uint64_t* ztag = nullptr;     // layer seed tables, indexed by seed_idx[job] (or by job if seed_idx == nullptr)
    const uint64_t* nlo = nullptr;
    const uint64_t* nhi = nullptr;
    const uint32_t* seed_idx = nullptr;
    const uint16_t* idx = nullptr;      // [n]
    const uint8_t* ch = nullptr;        // [n]
    const uint64_t* salt = nullptr;     // [n]
    const uint32_t* out_row = nullptr;  // [n] or nullptr (row = job)
    uint64_t* out = nullptr;            // rows [0, out_split)
    uint64_t out_split = ~0ull;         // rows >= out_split live in out2 (scratch rows of merged enc_value edges)
    uint64_t* out2 = nullptr;
};
int sigma_run(Ctx* ctx, const SigmaJobs& jobs);
// out[dst] ^= row src for every pair (dst_row, src_row); rows >= split are taken from out2
int sigma_xor_rows(Ctx* ctx, uint64_t npairs, const uint2* d_pairs, uint64_t* out, uint64_t split, const uint64_t* out2);
// exclusive scan of n u32 counts into n+1 offsets (arith.cu)
int scan_u32(Ctx* ctx, uint64_t n, const uint32_t* in, uint32_t* out);

// ---- ops
int op_enc_value(Ctx* ctx, const uint64_t* h_or_d_values, bool on_device, uint64_t n, uint64_t batch_seed, const uint64_t* h_states, Batch** out, int depth_hint = 0, int shares = 2,
                 uint64_t* h_next = nullptr /* optional out: first unused word of each item's tape */, const uint64_t* h_k0 = nullptr /* optional: first word to use */,
                 const uint64_t* h_ids = nullptr /* optional: global item numbers (default item_base + i) */);
int tape_spec(Ctx* ctx, Scratch& scratch, uint64_t n, uint64_t batch_seed, const uint64_t* h_states, const uint64_t* h_k0, const uint64_t* h_ids, TapeSpec& ts);
void plan_noise_host(const Ctx* ctx, int depth_hint, int& z2, int& z3);   // ctx == nullptr: default Params
int op_ct_add(Ctx* ctx, const Batch* A, const Batch* B, int mode /*0 add, 1 sub*/, Batch** out);
int op_ct_scale(Ctx* ctx, const Batch* A, Fp s, Batch** out);
int op_ct_mul(Ctx* ctx, const Batch* A, const Batch* B, uint64_t batch_seed, const uint64_t* h_states, Batch** out);
int op_dec_value(Ctx* ctx, const Batch* C, uint64_t* h_out /*n x 2*/);
int op_commit_ct(Ctx* ctx, const Batch* b, uint8_t* h_out /*n x 32*/);
void gen_ubk_perm_host(uint64_t canon_tag, uint16_t perm[kMBits]);
int batch_select(Ctx* ctx, const Batch* const* srcs, int nsrc, const std::vector<uint32_t>& which, const std::vector<uint32_t>& index, Batch** out);
int batch_concat(Ctx* ctx, const Batch* const* parts, size_t nparts, Batch** out);

// compact_layers (ops/encrypt.hpp:73-104) of every ciphertext of b, in place (layer arrays shrink, edges stay).
// d_err / h_err (optional): one more device word fetched by the same mailbox read (saves the caller a synchronisation)
int compact_layers_batch(Ctx* ctx, Batch* b, const unsigned int* d_err = nullptr, unsigned int* h_err = nullptr);
// guard_budget (ops/encrypt.hpp:106-111): compact_edges on every ciphertext of *pb with more than edge_budget edges; may replace *pb
int guard_budget_batch(Ctx* ctx, Batch** pb, uint32_t budget);

}  // namespace pvacb
