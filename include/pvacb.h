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
 * RNG tape. The reference draws 64-bit words from the OS CSPRNG with no seed hook (core/random.hpp:106-110). Here every item
 * of a call consumes its own word stream in exactly the order the reference would call csprng_u64() for that item; with the
 * same words the outputs are bit-identical to the reference (compiled with g++ 13 / libstdc++), including edge order.
 * The stream of item i (global index = item base + i) under the three tape kinds of a context (pvacb_set_tape):
 *   PVACB_TAPE_CHACHA20 (default; key = 256 bits from getrandom() at pvacb_ctx_create, or given by the caller):
 *       word k = little-endian 64-bit word (k mod 8) of ChaCha20 block (k div 8), block input words 12..15 =
 *       (block index, lane, stream id lo, stream id hi); from a batch_seed: stream id = batch_seed, lane = global index + 1;
 *       from an explicit per-item tape_states[i]: stream id = tape_states[i], lane = 0. Only keystream ever reaches a ciphertext.
 *       NEVER use one (key, batch_seed) or one (key, tape state) twice: pvacb_fresh_seed hands out seeds that do not repeat.
 *   PVACB_TAPE_SPLITMIX (parity tests and the committed golden vectors ONLY -- not secret-keeping: the generator is an unkeyed
 *       bijection of a 64-bit state and tape words are published as nonces):
 *       state0 = tape_states[i] or mix64(batch_seed + 0xD1342543DE82EF95 * (global index + 1)), word k = mix64(state0 + (k + 1) * 0x9E3779B97F4A7C15)
 *   PVACB_TAPE_WORDS: the caller supplies every word (pvacb_set_tape_words), e.g. straight from its own CSPRNG.
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
    PVACB_E_RESERVED7 = 7,   /* (round 1: rare rejection branches; both now run on the device) */
    PVACB_E_RESERVED8 = 8,   /* (round 1: ct_mul operands with repeated (layer, idx, sign); now summed like the reference does) */
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
#define PVACB_KEY_BLOB_BYTES ((size_t)(752 + 16384 * 128) * 8)

/* RNG tape kinds (see the header comment) */
enum { PVACB_TAPE_SPLITMIX = 0, PVACB_TAPE_CHACHA20 = 1, PVACB_TAPE_WORDS = 2 };

/* struct Params, core/types.hpp:36-70, field for field. The kernels are built for the shapes of the default set: B, m_bits, n_bits,
 * h_col_wt, x_col_wt, err_wt, lpn_n and lpn_tau must keep their defaults (anything else is refused with PVACB_E_ARG and a
 * message naming the field); noise_entropy_bits, tuple2_fraction, depth_slope_bits, edge_budget (4096..2^32-1), lpn_t (127..16384:
 * every value in that range gives identical ciphertexts; less than 16384 switches the context to the live-row evaluation, 16384 leaves
 * its PRF mode alone) and the recrypt fields (which the reference never reads either) are run-time values. */
typedef struct pvacb_params {
    int32_t B, m_bits, n_bits, h_col_wt, x_col_wt, err_wt;
    double noise_entropy_bits, tuple2_fraction, depth_slope_bits;
    uint64_t edge_budget;
    int32_t lpn_n, lpn_t, lpn_tau_num, lpn_tau_den;
    double recrypt_lo, recrypt_hi;
    int32_t recrypt_rounds;
} pvacb_params;
void pvacb_params_default(pvacb_params* p);

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
/* RNG tape of the context. key: 32 bytes for PVACB_TAPE_CHACHA20 (NULL = fresh from the OS CSPRNG), ignored otherwise. */
int pvacb_set_tape(pvacb_ctx* ctx, int kind, const uint8_t key[32]);
int pvacb_get_tape(const pvacb_ctx* ctx);
/* PVACB_TAPE_WORDS: words[(global item index) * words_per_item + k] is word k of that item. A call that needs more words than
 * supplied fails with PVACB_E_ARG (an enc_value item draws about 214, a ct_mul item 2 per new layer + 1 per output edge). */
int pvacb_set_tape_words(pvacb_ctx* ctx, const uint64_t* words, size_t n_items, size_t words_per_item);
/* global index of item 0 of the following calls (default 0): a shard of a larger batch keeps the streams of the whole batch, so
 * an N-GPU run returns the same bytes as a 1-GPU run */
int pvacb_set_item_base(pvacb_ctx* ctx, uint64_t base);
/* a batch_seed this context has not handed out before (a counter): for callers that have no use for reproducible seeds */
uint64_t pvacb_fresh_seed(pvacb_ctx* ctx);
/* test hooks, all bit-exact: what = 0: ct_mul planning takes the device-wide sort (a = 1) / the global bucket table (b = 1) even
 * for small pairs; what = 1: keystream word a of every PRF core is OR-ed with b (a = ~0: off), which reaches the rejection branch
 * of AesCtr256::bounded (crypto/lpn.hpp:141-148); what = 2: sigma kernel shape without spare PRG candidates (a = 1) */
int pvacb_debug_set(pvacb_ctx* ctx, int what, uint64_t a, uint64_t b);
/* optional per-kernel timing with CUDA events on the context's stream. Tags: 0 prf_lpn, 1 mul_pairs (weight products of ct_mul),
 * 2 sigma (fused sigma_from_H kernel), 3 concat (ct_add/ct_sub), 4 dec_edges, 5 commit, 6 compact_edges sort + merge. collect() synchronises,
 * sums ms and launch counts per tag. */
int pvacb_profile_enable(pvacb_ctx* ctx, int on);
/* microbenchmark: achieved GB/s of warp-wide 1 KiB gathers from the L2-resident matrix H (ceiling of the sigma kernel) */
int pvacb_l2_gather_probe(pvacb_ctx* ctx, int reps, double* gbps_out);
int pvacb_profile_collect(pvacb_ctx* ctx, float ms_out[8], uint32_t launches_out[8]);

/* ---- keys: replaces keygen(const Params&, PubKey&, SecKey&), crypto/keygen.hpp:35 ---------------------------------- */
/* keygen(prm, pk, sk): prm == NULL means the default Params; every word comes from ChaCha20 keyed by seed[32] (NULL = 32 fresh
 * bytes from the OS CSPRNG) in the reference's draw order. The run-time Params fields become those of the context. */
int pvacb_keygen_params(pvacb_ctx* ctx, const pvacb_params* prm, const uint8_t seed[32]);
int pvacb_set_params(pvacb_ctx* ctx, const pvacb_params* prm);     /* for keys that were imported as raw arrays */
int pvacb_get_params(const pvacb_ctx* ctx, pvacb_params* prm);
/* PARITY / TEST ONLY: keygen with default Params from the SplitMix64 stream of a 64-bit state (the committed golden keys). */
int pvacb_keygen(pvacb_ctx* ctx, uint64_t tape_state);
/* the reference's key files (savePk / loadPk / saveSk / loadSk, tests/bounty2_test.cpp:145-236; magics 0x06660666 / 0x66666999),
 * byte for byte. Either path may be NULL on export; on import sk_path == NULL gives a public-key-only context (ct_add / ct_sub /
 * ct_mul / commit_ct / ct_recrypt work, enc_* and dec_* return PVACB_E_NOKEYS). The Params stored in the pk file are checked like
 * pvacb_keygen_params checks them, and ubk.perm / ubk.inv must be the permutation canon_tag derives. */
int pvacb_keys_export_file(pvacb_ctx* ctx, const char* pk_path, const char* sk_path);
int pvacb_keys_import_file(pvacb_ctx* ctx, const char* pk_path, const char* sk_path);
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
/* Cipher enc_value_depth(pk, sk, v, depth_hint)            ops/encrypt.hpp:281  (plan_noise(depth_hint) noise groups; any depth_hint >= 0)
 * Cipher enc_zero_depth(pk, sk, depth_hint)                ops/encrypt.hpp:293  (identical draws to enc_value_depth(0, depth_hint)) */
int pvacb_enc_value_depth(pvacb_ctx* ctx, const uint64_t* values, size_t n, int depth_hint, uint64_t batch_seed, const uint64_t* tape_states,
                          pvacb_batch** out);
int pvacb_enc_zero_depth(pvacb_ctx* ctx, size_t n, int depth_hint, uint64_t batch_seed, const uint64_t* tape_states, pvacb_batch** out);
/* Cipher enc_fp_depth(pk, sk, Fp v, depth_hint)             ops/encrypt.hpp:162  (one share: 1 BASE layer; fp_values = n x (lo, hi), canonical) */
int pvacb_enc_fp_depth(pvacb_ctx* ctx, const uint64_t* fp_values, size_t n, int depth_hint, uint64_t batch_seed, const uint64_t* tape_states,
                       pvacb_batch** out);
/* std::pair<int,int> plan_noise(pk, depth_hint)            ops/encrypt.hpp:16  (default Params) */
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
 * for n messages at once: bytes = all messages back to back, msg_off[n+1] their byte ranges (any length: block j is encrypted with depth_hint 2 + j like the reference does). Message m
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

/* the whole device image of a batch with ONE copy (a batch is one allocation; the 13 arrays above sit at the offsets
 * pvacb_blob_layout gives for the counts of pvacb_batch_blob_info: off[0..12] = layer_off, edge_off, rule, ztag, nonce_lo,
 * nonce_hi, pa, pb, layer_id, idx, ch, w, sigma; off[13] = image size). Asynchronous like pvacb_batch_export_soa_async
 * (pvacb_export_wait / _wait_one); `host` should be pinned. */
int pvacb_blob_layout(uint64_t n, uint64_t n_layers, uint64_t n_edges, uint64_t off[14]);
int pvacb_batch_blob_info(const pvacb_batch* b, uint64_t* n, uint64_t* layout_layers, uint64_t* layout_edges, uint64_t* bytes);
int pvacb_batch_export_blob_async(pvacb_ctx* ctx, const pvacb_batch* b, void* host, size_t cap);
int pvacb_batch_import_blob(pvacb_ctx* ctx, size_t n, uint64_t layout_layers, uint64_t layout_edges, const void* host, size_t bytes, pvacb_batch** out);
/* Route the blob exports of this context through ANOTHER GPU's host link: the image crosses NVLink into a staging buffer on
 * relay_device and leaves from there (-1 = direct, the default). For boxes whose GPUs do not all reach host memory equally fast
 * (profiles/r02_hostlink_probe.txt: four of eight GPUs share 70 GB/s while the other four reach 176 GB/s together). */
int pvacb_set_export_relay(pvacb_ctx* ctx, int relay_device);

/* the reference's on-disk ciphertext format (tests/bounty2_test.cpp:17-143: magic 0x66699666, ver 1, u64 count, then per
 * cipher u32 nL, u32 nE, layers, edges). PROD-layer seeds are not part of that format. */
int pvacb_batch_wire_size(pvacb_ctx* ctx, const pvacb_batch* b, size_t* bytes);
int pvacb_batch_export_wire(pvacb_ctx* ctx, const pvacb_batch* b, void* buf, size_t cap, size_t* written);
int pvacb_batch_import_wire(pvacb_ctx* ctx, const void* buf, size_t bytes, pvacb_batch** out);

/* synthetic fresh-shaped ciphertexts for benchmarks: 2 BASE layers, edges_per_layer edges each,
 * uniform idx/ch/w/sigma from batch_seed. Not decryptable; exercises the bandwidth-bound ops. */
int pvacb_batch_synthetic(pvacb_ctx* ctx, size_t n, int edges_per_layer, uint64_t batch_seed, pvacb_batch** out);

/* ---- several GPUs of one box: N contexts of this process, one per device, one host thread per device while a
 * call runs. A batch is split into contiguous index ranges (item i -> member floor(i N / n)); every item keeps the RNG stream of
 * its GLOBAL index, so the result bytes do not depend on N; keys are replicated once by peer copies over NVLink; there is no
 * collective in the steady state. The reference has no counterpart (it is single-threaded); semantics per item are those of the
 * single-device calls above. */
typedef struct pvacb_group pvacb_group;
typedef struct pvacb_gbatch pvacb_gbatch;   /* a batch sharded over the members */
int pvacb_group_create(const int* devices, int n, pvacb_group** out);     /* one ChaCha20 tape key (from the OS) shared by all members */
void pvacb_group_destroy(pvacb_group* g);
int pvacb_group_size(const pvacb_group* g);
pvacb_ctx* pvacb_group_ctx(pvacb_group* g, int member);                   /* e.g. for pvacb_set_prf_mode / pvacb_stats per device */
const char* pvacb_group_last_error(const pvacb_group* g);
int pvacb_group_keygen_params(pvacb_group* g, const pvacb_params* prm, const uint8_t seed[32]);   /* keygen on member 0, then replicate */
int pvacb_group_keys_import_file(pvacb_group* g, const char* pk_path, const char* sk_path);
int pvacb_group_replicate_keys(pvacb_group* g);                           /* member 0's keys (however they got there) -> all members */
int pvacb_group_set_tape(pvacb_group* g, int kind, const uint8_t key[32]);
int pvacb_group_enc_value(pvacb_group* g, const uint64_t* values, size_t n, uint64_t batch_seed, pvacb_gbatch** out);
int pvacb_group_ct_add(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, pvacb_gbatch** out);
int pvacb_group_ct_sub(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, pvacb_gbatch** out);
int pvacb_group_ct_mul(pvacb_group* g, const pvacb_gbatch* a, const pvacb_gbatch* b, uint64_t batch_seed, pvacb_gbatch** out);
int pvacb_group_dec_value(pvacb_group* g, const pvacb_gbatch* c, uint64_t* out /* n x (lo,hi), global order */);
int pvacb_group_commit_ct(pvacb_group* g, const pvacb_gbatch* c, uint8_t* out /* n x 32, global order */);
void pvacb_group_batch_free(pvacb_gbatch* b);
size_t pvacb_group_batch_count(const pvacb_gbatch* b);
pvacb_batch* pvacb_group_batch_part(pvacb_gbatch* b, int member);         /* the member's slice (owned by the gbatch) */
/* every member exports the image of its slice into host_bufs[member] (pinned) at the same time; returns when all are complete */
int pvacb_group_export_blobs(pvacb_group* g, const pvacb_gbatch* c, void* const* host_bufs, const size_t* caps);
/* measures the simultaneous export bandwidth of the members (GB/s), direct and -- if their host links are unequal -- with the
 * slower half relayed through the faster half over NVLink; keeps the better routing for later exports */
int pvacb_group_tune_export(pvacb_group* g, double* direct_gbs, double* relay_gbs);

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
