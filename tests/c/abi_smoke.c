/* The C ABI from plain C (C99): proves include/pvacb.h needs no C++ and shows the calling sequence a cgo / JNI / ctypes
 * binding would make. Compiled (syntax + link) by the CPU test-suite, run on the GPU by tests/test_gpu_parity.py.
 *   gcc -std=c99 -Iinclude tests/c/abi_smoke.c -Lpvac_hfhe_cppbyv_b200 -lpvacb -Wl,-rpath,$PWD/pvac_hfhe_cppbyv_b200 -o abi_smoke */
#include <stdio.h>
#include <stdlib.h>

#include "pvacb.h"

#define CK(call)                                                                                  \
    do {                                                                                          \
        int rc_ = (call);                                                                         \
        if (rc_ != PVACB_OK) { fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, ctx ? pvacb_last_error(ctx) : ""); return 1; } \
    } while (0)

int main(void) {
    pvacb_ctx* ctx = NULL;
    pvacb_batch *a = NULL, *b = NULL, *s = NULL, *d = NULL, *p = NULL;
    const uint64_t va[3] = {42, 7, 0xFFFFFFFFFFFFFFFFull}, vb[3] = {17, 6, 2};
    uint64_t out[6];
    uint8_t digest[3 * 32];
    uint64_t nl = 0, ne = 0;
    int i;

    CK(pvacb_ctx_create(0, &ctx));
    CK(pvacb_set_prf_mode(ctx, PVACB_PRF_LIVE));
    CK(pvacb_keygen(ctx, 1));
    CK(pvacb_enc_value(ctx, va, 3, 1000, &a));
    CK(pvacb_enc_value(ctx, vb, 3, 2000, &b));
    CK(pvacb_ct_add(ctx, a, b, &s));
    CK(pvacb_ct_sub(ctx, a, b, &d));
    CK(pvacb_ct_mul(ctx, a, b, 3000, &p));
    CK(pvacb_dec_value(ctx, s, out));
    for (i = 0; i < 2; i++)
        if (out[2 * i] != va[i] + vb[i] || out[2 * i + 1] != 0) { fprintf(stderr, "add mismatch at %d\n", i); return 1; }
    CK(pvacb_dec_value(ctx, d, out));
    if (out[0] != 25 || out[2] != 1) { fprintf(stderr, "sub mismatch\n"); return 1; }
    CK(pvacb_dec_value(ctx, p, out));
    if (out[0] != 42 * 17 || out[2] != 42 || out[1] != 0) { fprintf(stderr, "mul mismatch\n"); return 1; }
    /* (2^64-1) * 2 = 2^65 - 2 : lo = 2^64 - 2, hi = 1 */
    if (out[4] != 0xFFFFFFFFFFFFFFFEull || out[5] != 1) { fprintf(stderr, "mul mismatch (wide)\n"); return 1; }
    CK(pvacb_batch_totals(p, &nl, &ne));
    CK(pvacb_commit_ct(ctx, p, digest));
    printf("abi_smoke ok: %zu products, %llu layers, %llu edges, commit[0] = %02x%02x%02x%02x...\n", pvacb_batch_count(p), (unsigned long long)nl,
           (unsigned long long)ne, digest[0], digest[1], digest[2], digest[3]);
    pvacb_batch_free(a); pvacb_batch_free(b); pvacb_batch_free(s); pvacb_batch_free(d); pvacb_batch_free(p);
    pvacb_ctx_destroy(ctx);
    return 0;
}
