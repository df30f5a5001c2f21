// enc_text / dec_text (utils/text.hpp:15-87) for a batch of messages, and concatenation of batches along the item axis.
//
// Reference, per message: out[0] = enc_value(length); then one enc_fp_depth(pack_15_bytes(block j), depth_hint = 2 + j) per
// 15-byte block, all drawing from ONE CSPRNG stream in that order. Batched restatement: "waves" -- wave 0 encrypts the
// lengths of all messages (enc_value), wave j+1 encrypts block j of every message that has one (one enc_fp_depth launch
// with depth_hint 2 + j). A message's tape continues from wave to wave: the planning kernel reports how many words each
// ciphertext consumed (data-dependent: rejection loops) and the next wave starts that many words further (every tape kind is
// addressed by word index: a wave opens the item's stream at word k0 = words consumed so far).
// The resulting batch is WAVE-MAJOR: all length ciphertexts in message order, then every block-0 ciphertext, then block 1 ...
#include "engine.h"
#include "../../include/pvacb.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace pvacb {

__global__ void concat_fix_offsets_kernel(uint64_t cnt, const uint32_t* __restrict__ src_l, const uint32_t* __restrict__ src_e, uint32_t base_l,
                                          uint32_t base_e, uint32_t* __restrict__ dst_l, uint32_t* __restrict__ dst_e, int last) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > cnt || (i == cnt && !last)) return;
    dst_l[i] = src_l[i] + base_l;
    dst_e[i] = src_e[i] + base_e;
}

// items of parts[0], then parts[1], ... as one new batch
int batch_concat(Ctx* ctx, const Batch* const* parts, size_t nparts, Batch** out) {
    uint64_t n = 0, nL = 0, nE = 0;
    for (size_t k = 0; k < nparts; k++) { n += parts[k]->n; nL += parts[k]->nL; nE += parts[k]->nE; }
    Batch* o = nullptr;
    int rc = batch_alloc(ctx, n, nL, nE, &o);
    if (rc) return rc;
    auto cp = [&](void* d, const void* s, size_t bytes) { return bytes ? cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, ctx->stream) : cudaSuccess; };
    uint64_t i0 = 0, l0 = 0, e0 = 0;
    cudaError_t err = cudaSuccess;
    auto ck = [&](cudaError_t x) { if (err == cudaSuccess) err = x; };
    if (n == 0) ck(cudaMemsetAsync(o->loff, 0, 4, ctx->stream)), ck(cudaMemsetAsync(o->eoff, 0, 4, ctx->stream));
    for (size_t k = 0; k < nparts; k++) {
        const Batch* s = parts[k];
        const int last = (k + 1 == nparts) ? 1 : 0;
        if (s->n || last)
            concat_fix_offsets_kernel<<<(unsigned)((s->n + 1 + 255) / 256), 256, 0, ctx->stream>>>(s->n, s->loff, s->eoff, (uint32_t)l0, (uint32_t)e0, o->loff + i0,
                                                                                                    o->eoff + i0, last);
        ck(cp(o->rule + l0, s->rule, s->nL)); ck(cp(o->ztag + l0, s->ztag, s->nL * 8)); ck(cp(o->nlo + l0, s->nlo, s->nL * 8));
        ck(cp(o->nhi + l0, s->nhi, s->nL * 8)); ck(cp(o->pa + l0, s->pa, s->nL * 4)); ck(cp(o->pb + l0, s->pb, s->nL * 4));
        ck(cp(o->lid + e0, s->lid, s->nE * 4)); ck(cp(o->idx + e0, s->idx, s->nE * 2)); ck(cp(o->ch + e0, s->ch, s->nE));
        ck(cp(o->w + e0, s->w, s->nE * 16)); ck(cp(o->sigma + e0 * kMWords, s->sigma, s->nE * (size_t)kMWords * 8));
        i0 += s->n; l0 += s->nL; e0 += s->nE;
    }
    ck(cudaGetLastError());
    if (err != cudaSuccess) {
        batch_free(o);
        ctx->last_error = std::string("batch_concat: ") + cudaGetErrorString(err);
        return PV_E_CUDA;
    }
    ctx->stat_kernel_launches += nparts;
    *out = o;
    return PV_OK;
}

// utils/text.hpp:15-27 : up to 15 bytes little-endian into (lo, hi); always canonical (< 2^120)
static void pack15(const uint8_t* p, size_t len, uint64_t& lo, uint64_t& hi) {
    lo = hi = 0;
    for (size_t i = 0; i < len && i < 15; i++) {
        if (i < 8) lo |= (uint64_t)p[i] << (8 * i);
        else hi |= (uint64_t)p[i] << (8 * (i - 8));
    }
}

}  // namespace pvacb

using namespace pvacb;
static inline Ctx* C(pvacb_ctx* x) { return reinterpret_cast<Ctx*>(x); }
static inline const Batch* Bt(const pvacb_batch* x) { return reinterpret_cast<const Batch*>(x); }

extern "C" {

int pvacb_batch_concat(pvacb_ctx* x, const pvacb_batch* const* parts, size_t nparts, pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || (nparts && !parts)) return PV_E_ARG;
    cudaSetDevice(ctx->device);
    std::vector<const Batch*> v(nparts);
    for (size_t k = 0; k < nparts; k++) { if (!parts[k]) return PV_E_ARG; v[k] = Bt(parts[k]); }
    Batch* o = nullptr;
    int rc = batch_concat(ctx, v.data(), nparts, &o);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}

int pvacb_enc_text(pvacb_ctx* x, const uint8_t* bytes, const uint64_t* msg_off, size_t n, uint64_t batch_seed, const uint64_t* tape_states,
                   pvacb_batch** out) {
    Ctx* ctx = C(x);
    if (!out || !msg_off || (msg_off[n] && !bytes)) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    cudaSetDevice(ctx->device);
    size_t max_blocks = 0;
    for (size_t m = 0; m < n; m++) {
        if (msg_off[m + 1] < msg_off[m]) return PV_E_ARG;
        max_blocks = std::max(max_blocks, (size_t)((msg_off[m + 1] - msg_off[m] + 14) / 15));
    }
    std::vector<Batch*> waves;
    auto fail = [&](int rc) { for (Batch* b : waves) batch_free(b); return rc; };
    // wave 0: enc_value(length)
    // every message walks ONE tape stream (its item's): a wave starts each message at the word the previous wave stopped at
    std::vector<uint64_t> lens(n), next(n, 0);
    for (size_t m = 0; m < n; m++) lens[m] = msg_off[m + 1] - msg_off[m];
    Batch* b = nullptr;
    int rc = op_enc_value(ctx, lens.data(), false, n, batch_seed, tape_states, &b, 0, 2, next.data());
    if (rc) return fail(rc);
    waves.push_back(b);
    // waves 1..: block j of every message that has one, enc_fp_depth(pack15, 2 + j)
    std::vector<uint64_t> vals, st, k0, ids, dr;
    std::vector<size_t> who;
    const uint64_t base = ctx->item_base;
    for (size_t j = 0; j < max_blocks; j++) {
        vals.clear(); st.clear(); who.clear(); k0.clear(); ids.clear();
        for (size_t m = 0; m < n; m++) {
            if (lens[m] <= 15 * j) continue;
            uint64_t lo, hi;
            pack15(bytes + msg_off[m] + 15 * j, (size_t)std::min<uint64_t>(15, lens[m] - 15 * j), lo, hi);
            vals.push_back(lo); vals.push_back(hi);
            if (tape_states) st.push_back(tape_states[m]);
            k0.push_back(next[m]);
            ids.push_back(base + m);
            who.push_back(m);
        }
        dr.assign(who.size(), 0);
        b = nullptr;
        rc = op_enc_value(ctx, vals.data(), false, who.size(), batch_seed, tape_states ? st.data() : nullptr, &b, 2 + (int)j, 1, dr.data(), k0.data(), ids.data());
        if (rc) return fail(rc);
        waves.push_back(b);
        for (size_t q = 0; q < who.size(); q++) next[who[q]] = dr[q];
    }
    std::vector<const Batch*> parts(waves.begin(), waves.end());
    Batch* o = nullptr;
    rc = batch_concat(ctx, parts.data(), parts.size(), &o);
    if (rc == PV_OK) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { batch_free(o); o = nullptr; rc = PV_E_CUDA; ctx->last_error = cudaGetErrorString(e); }
    }
    for (Batch* w : waves) batch_free(w);
    *out = reinterpret_cast<pvacb_batch*>(o);
    return rc;
}

int pvacb_dec_text(pvacb_ctx* x, const pvacb_batch* pb, size_t n_msgs, uint8_t* out_bytes, size_t cap, uint64_t* out_off) {
    Ctx* ctx = C(x);
    const Batch* c = Bt(pb);
    if (!out_off || (cap && !out_bytes)) return PV_E_ARG;
    if (!ctx->have_keys) return PV_E_NOKEYS;
    if (c->n < n_msgs) { ctx->last_error = "dec_text: fewer ciphertexts than messages"; return PV_E_FORMAT; }
    cudaSetDevice(ctx->device);
    std::vector<uint64_t> d(2 * (size_t)c->n);
    int rc = op_dec_value(ctx, c, d.data());
    if (rc) return rc;
    // lengths first; block j of message m sits in wave j + 1 at the rank of m among the messages with more than j blocks
    std::vector<uint64_t> lens(n_msgs), nblk(n_msgs);
    uint64_t total_blocks = 0, max_blocks = 0;
    for (size_t m = 0; m < n_msgs; m++) {
        if (d[2 * m + 1] != 0) { ctx->last_error = "dec_text: a length does not fit 64 bits"; return PV_E_FORMAT; }
        lens[m] = d[2 * m];
        nblk[m] = (lens[m] + 14) / 15;
        total_blocks += nblk[m];
        max_blocks = std::max(max_blocks, nblk[m]);
    }
    if (n_msgs + total_blocks != c->n) { ctx->last_error = "dec_text: ciphertext count does not match the decrypted lengths"; return PV_E_FORMAT; }
    uint64_t off = 0;
    for (size_t m = 0; m < n_msgs; m++) { out_off[m] = off; off += lens[m]; }
    out_off[n_msgs] = off;
    if (off > cap) { ctx->last_error = "dec_text: output buffer too small"; return PV_E_ARG; }
    uint64_t pos = n_msgs;
    for (uint64_t j = 0; j < max_blocks; j++)
        for (size_t m = 0; m < n_msgs; m++) {
            if (nblk[m] <= j) continue;
            const uint64_t lo = d[2 * pos], hi = d[2 * pos + 1];
            pos++;
            const uint64_t take = std::min<uint64_t>(15, lens[m] - 15 * j);
            uint8_t* o = out_bytes + out_off[m] + 15 * j;
            for (uint64_t i = 0; i < take; i++) o[i] = (uint8_t)(i < 8 ? lo >> (8 * i) : hi >> (8 * (i - 8)));   // utils/text.hpp:29-37
        }
    return PV_OK;
}

}  // extern "C"
