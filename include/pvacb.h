/* pvacb -- batched B200 (sm_100a) engine for the data-parallel path of pvac-hfhe.
 *
 * C ABI of libpvacb.so. The reference (vasihh2009/pvac_hfhe_cppbyv) has no FFI layer: its boundary is the header-only
 * C++ API in include/pvac/ (paths below are relative to that directory). Every entry point here is the batched
 * drop-in for one of those functions: an array of independent ciphertexts is processed per call, ciphertexts stay on
 * the device between calls as opaque pvacb_batch handles (structure-of-arrays), and nothing C++ or CUDA crosses the
 * boundary. There is no CPU fallback: every call fails with PVACB_E_CUDA if no sm_100-class device is usable.
 *
 * Conventions (mirroring the reference): value semantics -- inputs are never modified, every op
 * returns a new batch that the caller frees with pvacb_batch_free; where the reference calls std::abort() the batched
 * call returns a non-zero status instead (pvacb_last_error gives text). A context is single-caller (one host thread).
 *
 * RNG tape. The reference draws 64-bit words from the OS CSPRNG with no seed hook (core/random.hpp:106-110). Here the
 * caller passes a 64-bit batch_seed; item i of a batch consumes the SplitMix64 stream
 *     state0 = mix64(batch_seed + 0xD1342543DE82EF95 * (i + 1)),  word k = mix64(state0 + (k + 1) * 0x9E3779B97F4A7C15)
 * in exactly the order the reference would call csprng_u64() for that item. With the same words the outputs are
 * bit-identical to the reference (compiled with g++ 13 / libstdc++), including edge order.
 */
#ifndef PVACB_H
#define PVACB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pvacb_ctx pvacb_ctx;     /* one device: streams, replicated keys, scratch */
typedef struct pvacb_batch pvacb_batch; /* device-resident array of ciphertexts (pvac::Cipher, core/types.hpp:116-119) */

enum {
    PVACB_OK = 0,
    PVACB_E_ARG = 1,
    PVACB_E_CUDA = 2,
    PVACB_E_OOM = 3,         /* reference: std::bad_alloc */
    PVACB_E_NOKEYS = 4,
    PVACB_E_EDGE_BUDGET = 5, /* reserved (results over Params::edge_budget are compacted like the reference's guard_budget does) */
    PVACB_E_LAYER_GRAPH = 6, /* reference: std::abort() in layer_R_cached (ops/decrypt.hpp:23,36) */
    PVACB_E_RARE_PATH = 7,   /* AesCtr256::bounded rejected a word (p = 2^-61 per LPN row, crypto/lpn.hpp:141-148) */
    PVACB_E_DUP_EDGE = 8,    /* ct_mul input with two edges of equal (layer, idx, sign) */
    PVACB_E_FORMAT = 9,
    PVACB_E_SHAPE = 10
};

/* PRF evaluation mode. Both are bit-exact. FAITHFUL evaluates all lpn_t = 16384 LPN rows like the reference
 * (crypto/lpn.hpp:219); LIVE evaluates rows 0..127, the only ones toep_127 can observe (crypto/toeplitz.hpp:153-162). */
enum { PVACB_PRF_FAITHFUL = 0, PVACB_PRF_LIVE = 1 };

/* sizes of the default Params (core/types.hpp:36-70), the only parameter set this engine is built for */
#define PVACB_B 337
#define PVACB_M_WORDS 128      /* sigma: m_bits/64 */
#define PVACB_N_COLS 16384     /* columns of H */
#define PVACB_LPN_WORDS 64
#define PVACB_KEY_BLOB_BYTES ((size_t)(748 + 16384 * 128) * 8)

/* ---- context ------------------------------------------------------------------------------------------------- */
int pvacb_ctx_create(int device, pvacb_ctx** out);
void pvacb_ctx_destroy(pvacb_ctx* ctx);
const char* pvacb_last_error(const pvacb_ctx* ctx);
int pvacb_set_prf_mode(pvacb_ctx* ctx, int mode);
int pvacb_get_prf_mode(const pvacb_ctx* ctx);
void* pvacb_stream(pvacb_ctx* ctx);       /* the cudaStream_t every kernel of this context is launched on */
int pvacb_sync(pvacb_ctx* ctx);
/* counters since the last reset: kernels launched, AES-256 blocks computed, sigma_from_H evaluations */
void pvacb_stats(const pvacb_ctx* ctx, uint64_t* kernel_launches, uint64_t* aes_blocks, uint64_t* sigma_edges);
void pvacb_stats_reset(pvacb_ctx* ctx);
/* optional per-kernel timing with CUDA events on the context's stream. Tags: 0 prf_lpn, 1 unused, 2 sigma (fused sigma_from_H kernel),
 * 3 concat (ct_add/ct_sub), 4 dec_edges, 5 mul planning. collect() synchronises, sums ms and launch counts per tag. */
int pvacb_profile_enable(pvacb_ctx* ctx, int on);
/* microbenchmark: achieved GB/s of warp-wide 1 KiB gathers from the L2-resident matrix H (ceiling of the sigma kernel) */
int pvacb_l2_gather_probe(pvacb_ctx* ctx, int reps, double* gbps_out);
int pvacb_profile_collect(pvacb_ctx* ctx, float ms_out[8], uint32_t launches_out[8]);

/* ---- keys: replaces keygen(const Params&, PubKey&, SecKey&), crypto/keygen.hpp:35 ---------------------------------- */
/* keygen with default Params; consumes the tape stream with initial state `tape_state` exactly like the reference. */
int pvacb_keygen(pvacb_ctx* ctx, uint64_t tape_state);
/* PubKey/SecKey fields as raw arrays (core/types.hpp:121-134): H[16384][128], powg_B[337][2] (lo,hi), prf_k[4], lpn_s[64] */
int pvacb_keys_import_raw(pvacb_ctx* ctx, uint64_t canon_tag, const uint8_t h_digest[32], const uint64_t* H,
                          const uint64_t* powg_B, const uint64_t prf_k[4], const uint64_t* lpn_s);
int pvacb_keys_export_raw(pvacb_ctx* ctx, uint64_t* canon_tag, uint8_t h_digest[32], uint64_t* H /* may be NULL */,
                          uint64_t* powg_B, uint64_t prf_k[4], uint64_t* lpn_s);
/* the flat device-resident key blob (PVACB_KEY_BLOB_BYTES): broadcast it to the other GPUs of the box (NCCL broadcast or
 * cudaMemcpyPeer over NVLink) and adopt it there. */
int pvacb_keys_device_blob(pvacb_ctx* ctx, void** dptr, size_t* bytes);
int pvacb_keys_alloc_blob(pvacb_ctx* ctx, void** dptr);            /* empty blob on this device to receive a broadcast */
int pvacb_keys_adopt_blob(pvacb_ctx* ctx);                         /* after the blob has been filled */
/* copy the blob to / adopt it from an arbitrary device buffer (e.g. a tensor that an NCCL broadcast filled) */
int pvacb_keys_copy_blob_to(pvacb_ctx* ctx, void* dst_device);
int pvacb_keys_adopt_blob_from(pvacb_ctx* ctx, const void* src_device);

/* ---- the hot path ------------------------------------------------------------------------------------------------ */
/* Cipher enc_value(pk, sk, uint64_t)                      ops/encrypt.hpp:289 */
int pvacb_enc_value(pvacb_ctx* ctx, const uint64_t* values, size_t n, uint64_t batch_seed, pvacb_batch** out);
/* same, with the initial tape state of every item given explicitly (n host words, e.g. drawn from the OS CSPRNG);
 * tape_states == NULL falls back to the batch_seed derivation above. */
int pvacb_enc_value_ex(pvacb_ctx* ctx, const uint64_t* values, size_t n, uint64_t batch_seed, const uint64_t* tape_states,
                       pvacb_batch** out);
/* Cipher enc_value_depth(pk, sk, v, depth_hint)            ops/encrypt.hpp:281  (plan_noise(depth_hint) noise groups; depth_hint 0..23)
 * Cipher enc_zero_depth(pk, sk, depth_hint)                ops/encrypt.hpp:293  (identical draws to enc_value_depth(0, depth_hint)) */
int pvacb_enc_value_depth(pvacb_ctx* ctx, const uint64_t* values, size_t n, int depth_hint, uint64_t batch_seed, const uint64_t* tape_states,
                          pvacb_batch** out);
int pvacb_enc_zero_depth(pvacb_ctx* ctx, size_t n, int depth_hint, uint64_t batch_seed, const uint64_t* tape_states, pvacb_batch** out);
/* Cipher enc_fp_depth(pk, sk, Fp v, depth_hint)             ops/encrypt.hpp:162  (one share: 1 BASE layer; fp_values = n x (lo, hi), canonical) */
int pvacb_enc_fp_depth(pvacb_ctx* ctx, const uint64_t* fp_values, size_t n, int depth_hint, uint64_t batch_seed, const uint64_t* tape_states,
                       pvacb_batch** out);
/* std::pair<int,int> plan_noise(pk, depth_hint)            ops/encrypt.hpp:16 */
int pvacb_plan_noise(int depth_hint, int* z2, int* z3);
/* Cipher ct_add / ct_sub(pk, A, B)                        ops/arithmetic.hpp:12,43 */
int pvacb_ct_add(pvacb_ctx* ctx, const pvacb_batch* a, const pvacb_batch* b, pvacb_batch** out);
int pvacb_ct_sub(pvacb_ctx* ctx, const pvacb_batch* a, const pvacb_batch* b, pvacb_batch** out);
/* Cipher ct_scale(pk, A, Fp s)                            ops/arithmetic.hpp:33  (s = lo,hi; one scalar for the batch) */
int pvacb_ct_scale(pvacb_ctx* ctx, const pvacb_batch* a, const uint64_t s[2], pvacb_batch** out);
/* Cipher ct_neg(pk, A), ct_div_const(pk, A, Fp k)          ops/arithmetic.hpp:39,108 */
int pvacb_ct_neg(pvacb_ctx* ctx, const pvacb_batch* a, pvacb_batch** out);
int pvacb_ct_div_const(pvacb_ctx* ctx, const pvacb_batch* a, const uint64_t k[2], pvacb_batch** out);
/* void compact_edges(pk, Cipher&)                          ops/encrypt.hpp:39  (returns a new batch; also what ct_add / ct_mul
 * apply to any result with more than Params::edge_budget edges, like the reference's guard_budget, ops/encrypt.hpp:106) */
int pvacb_compact_edges(pvacb_ctx* ctx, const pvacb_batch* a, pvacb_batch** out);
/* Cipher ct_mul(pk, A, B)                                 ops/arithmetic.hpp:47  (draws nonces and salts from the tape) */
int pvacb_ct_mul(pvacb_ctx* ctx, const pvacb_batch* a, const pvacb_batch* b, uint64_t batch_seed, pvacb_batch** out);
int pvacb_ct_mul_ex(pvacb_ctx* ctx, const pvacb_batch* a, const pvacb_batch* b, uint64_t batch_seed, const uint64_t* tape_states,
                    pvacb_batch** out);
/* Fp dec_value(pk, sk, C)                                 ops/decrypt.hpp:62   out: n x (lo,hi) host words */
int pvacb_dec_value(pvacb_ctx* ctx, const pvacb_batch* c, uint64_t* out);

/* std::array<uint8_t,32> commit_ct(pk, C)                  ops/commit.hpp:12   out: n x 32 bytes (SHA-256 over layers, edges and sigma) */
int pvacb_commit_ct(pvacb_ctx* ctx, const pvacb_batch* c, uint8_t* out);

/* std::vector<Cipher> enc_text(pk, sk, msg) / std::string dec_text(pk, sk, cts)     utils/text.hpp:39,63
 * for n messages at once: bytes = all messages back to back, msg_off[n+1] their byte ranges (<= 330 bytes each). Message m
 * draws from ONE tape stream (tape_states[m] or the batch_seed derivation) exactly like the reference: enc_value(length), then
 * enc_fp_depth(15-byte block j, depth_hint 2 + j). The result batch is WAVE-MAJOR: the n length ciphertexts in message
 * order, then block 0 of every message that has one (message order), then block 1, ... pvacb_dec_text expects that order. */
int pvacb_enc_text(pvacb_ctx* ctx, const uint8_t* bytes, const uint64_t* msg_off, size_t n, uint64_t batch_seed, const uint64_t* tape_states,
                   pvacb_batch** out);
int pvacb_dec_text(pvacb_ctx* ctx, const pvacb_batch* c, size_t n_msgs, uint8_t* out_bytes, size_t cap, uint64_t* out_off /* n_msgs + 1 */);

/* Cipher ct_recrypt(pk, ek, C)                             ops/recrypt.hpp:26  (zero_pool = EvalKey::zero_pool as a batch, e.g. from
 * pvacb_enc_zero_depth; item i draws its pool indices from its tape stream). double sigma_density(pk, C)  ops/encrypt.hpp:29.
 * void ubk_apply(pk, C)  crypto/matrix.hpp:306 (returns a new batch). pvacb_ubk_perm: Ubk::perm (8192 entries). */
int pvacb_ct_recrypt(pvacb_ctx* ctx, const pvacb_batch* c, const pvacb_batch* zero_pool, uint64_t batch_seed, const uint64_t* tape_states, pvacb_batch** out);
int pvacb_sigma_density(pvacb_ctx* ctx, const pvacb_batch* c, double* out /* n */);
int pvacb_ubk_apply(pvacb_ctx* ctx, const pvacb_batch* c, pvacb_batch** out);
int pvacb_ubk_perm(pvacb_ctx* ctx, uint16_t* perm_out /* 8192 */);

/* ---- batches --------------------------------------------------------------------------------------------------- */
void pvacb_batch_free(pvacb_batch* b);
size_t pvacb_batch_count(const pvacb_batch* b);
int pvacb_batch_totals(const pvacb_batch* b, uint64_t* n_layers, uint64_t* n_edges);
size_t pvacb_batch_device_bytes(const pvacb_batch* b);
/* per-ciphertext layer / edge offsets, n+1 entries each (host) */
int pvacb_batch_offsets(pvacb_ctx* ctx, const pvacb_batch* b, uint32_t* layer_off, uint32_t* edge_off);
/* the items of parts[0], parts[1], ... as one new batch (device copy; the inverse of pvacb_batch_slice) */
int pvacb_batch_concat(pvacb_ctx* ctx, const pvacb_batch* const* parts, size_t nparts, pvacb_batch** out);
/* out item i = item index[i] of srcs[which[i]] (up to 4 source batches): gather / scatter / interleave of ciphertexts */
int pvacb_batch_select(pvacb_ctx* ctx, const pvacb_batch* const* srcs, int nsrc, const uint32_t* which, const uint32_t* index, size_t n, pvacb_batch** out);
/* slice [first, first+count) of a batch as a new batch (device copy) */
int pvacb_batch_slice(pvacb_ctx* ctx, const pvacb_batch* b, size_t first, size_t count, pvacb_batch** out);

/* order-independent checksums of a whole batch, computed on the device (for parity checks at sizes that are not exported):
 * out[0] XOR of all sigma words; out[1], out[2] sums of w.lo, w.hi; out[3..5] sums of layer_id, idx, ch (mod 2^64);
 * out[6] sum over layers of (ztag ^ nonce_lo ^ nonce_hi) for BASE, pa + (pb << 32) for PROD; out[7] number of edges */
int pvacb_batch_checksum(pvacb_ctx* ctx, const pvacb_batch* b, uint64_t out[8]);

/* struct-of-arrays host export / import (for parity checks and interop). Any output pointer may be NULL.
 * layers: rule u8 (0 BASE, 1 PROD), ztag, nonce_lo, nonce_hi u64, pa, pb u32 (0 for BASE);
 * edges: layer_id u32 (relative to its ciphertext), idx u16, ch u8, w u64[2] (lo,hi), sigma u64[128]. */
int pvacb_batch_export_soa(pvacb_ctx* ctx, const pvacb_batch* b, uint32_t* layer_off, uint32_t* edge_off, uint8_t* rule,
                           uint64_t* ztag, uint64_t* nonce_lo, uint64_t* nonce_hi, uint32_t* pa, uint32_t* pb,
                           uint32_t* layer_id, uint16_t* idx, uint8_t* ch, uint64_t* w, uint64_t* sigma);
/* asynchronous form: the copies run on the context's second stream after everything queued so far; the batch and the
 * (pinned) destination buffers must stay alive until pvacb_export_wait returns. Lets the device->host read of one batch
 * overlap the computation of the next. */
int pvacb_batch_export_soa_async(pvacb_ctx* ctx, const pvacb_batch* b, uint32_t* layer_off, uint32_t* edge_off, uint8_t* rule,
                                 uint64_t* ztag, uint64_t* nonce_lo, uint64_t* nonce_hi, uint32_t* pa, uint32_t* pb,
                                 uint32_t* layer_id, uint16_t* idx, uint8_t* ch, uint64_t* w, uint64_t* sigma);
int pvacb_export_wait(pvacb_ctx* ctx);            /* all asynchronous exports issued so far are complete */
/* only the OLDEST outstanding asynchronous export is complete: with two sets of host buffers the next export can already be
 * queued behind the one being waited for, so the copy engine never idles between them */
int pvacb_export_wait_one(pvacb_ctx* ctx);
int pvacb_batch_import_soa(pvacb_ctx* ctx, size_t n, const uint32_t* layer_off, const uint32_t* edge_off, const uint8_t* rule,
                           const uint64_t* ztag, const uint64_t* nonce_lo, const uint64_t* nonce_hi, const uint32_t* pa,
                           const uint32_t* pb, const uint32_t* layer_id, const uint16_t* idx, const uint8_t* ch,
                           const uint64_t* w, const uint64_t* sigma, pvacb_batch** out);

/* the reference's on-disk ciphertext format (tests/bounty2_test.cpp:17-143: magic 0x66699666, ver 1, u64 count, then per
 * cipher u32 nL, u32 nE, layers, edges). PROD-layer seeds are not part of that format. */
int pvacb_batch_wire_size(pvacb_ctx* ctx, const pvacb_batch* b, size_t* bytes);
int pvacb_batch_export_wire(pvacb_ctx* ctx, const pvacb_batch* b, void* buf, size_t cap, size_t* written);
int pvacb_batch_import_wire(pvacb_ctx* ctx, const void* buf, size_t bytes, pvacb_batch** out);

/* synthetic fresh-shaped ciphertexts for benchmarks: 2 BASE layers, edges_per_layer edges each,
 * uniform idx/ch/w/sigma from batch_seed. Not decryptable; exercises the bandwidth-bound ops. */
int pvacb_batch_synthetic(pvacb_ctx* ctx, size_t n, int edges_per_layer, uint64_t batch_seed, pvacb_batch** out);

/* ---- building blocks exposed for parity tests (device kernels, not CPU code) ------------------------------------- */
/* prf_R (family 0) / prf_R_noise (family 1) of n seeds, crypto/lpn.hpp:263-275. ybits (optional) receives every
 * core's LPN sample: n*3 cores x (16384/64 or 128/64) words. */
int pvacb_prf(pvacb_ctx* ctx, size_t n, const uint64_t* ztag, const uint64_t* nonce_lo, const uint64_t* nonce_hi,
              int family, uint64_t* out /* n x 2 */, uint64_t* ybits);
/* sigma_from_H of n edges, crypto/matrix.hpp:267-303; out n x 128 words */
int pvacb_sigma_from_H(pvacb_ctx* ctx, size_t n, const uint64_t* ztag, const uint64_t* nonce_lo, const uint64_t* nonce_hi,
                       const uint16_t* idx, const uint8_t* ch, const uint64_t* salt, uint64_t* out);
/* Fp arithmetic on arrays (core/field.hpp): op 0 add, 1 sub, 2 mul, 3 neg(a), 4 inv(a) */
int pvacb_fp_op(pvacb_ctx* ctx, int op, size_t n, const uint64_t* a, const uint64_t* b, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* PVACB_H */
