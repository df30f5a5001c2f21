// Fp = GF(2^127 - 1) on two 64-bit limbs. Restates core/field.hpp:17-273 of the reference: every result is the
// canonical residue in [0, p), so any correct reduction strategy is bit-identical to the reference's.
// Device code: 64x64->128 products via mul.lo/mul.hi (lowered by ptxas to IMAD.WIDE.U32 chains with carry),
// Mersenne folding by shift-and-add.
#pragma once
#include "common.cuh"

namespace pvacb {

struct Fp {
    uint64_t lo, hi;
};

PV_HD Fp fp_make(uint64_t lo, uint64_t hi) { Fp r; r.lo = lo; r.hi = hi; return r; }
PV_HD Fp fp_zero() { return fp_make(0, 0); }
PV_HD Fp fp_one() { return fp_make(1, 0); }
PV_HD bool fp_is_zero(const Fp& a) { return (a.lo | a.hi) == 0; }
PV_HD bool fp_eq(const Fp& a, const Fp& b) { return a.lo == b.lo && a.hi == b.hi; }

// core/field.hpp:26-48 : fold bit 127 into bit 0, then one conditional subtraction of p (p -> 0, 2^127 -> 1)
PV_HD Fp fp_from_words(uint64_t lo, uint64_t hi) {
    uint64_t top = hi >> 63;
    hi &= kMask63;
    lo += top;
    hi += (lo < top) ? 1ull : 0ull;
    if (hi >> 63) return fp_make(1, 0);                        // value was exactly 2^127
    if (hi == kMask63 && lo == ~0ull) return fp_make(0, 0);    // value was exactly p
    return fp_make(lo, hi);
}

// core/field.hpp:50-56 (canonical inputs: the 128-bit sum cannot overflow)
PV_HD Fp fp_add(const Fp& a, const Fp& b) {
    uint64_t lo = a.lo + b.lo;
    uint64_t hi = a.hi + b.hi + ((lo < a.lo) ? 1ull : 0ull);
    return fp_from_words(lo, hi);
}
// core/field.hpp:58-67
PV_HD Fp fp_neg(const Fp& a) {
    uint64_t lo = ~0ull - a.lo;  // no borrow: p.lo is all ones
    uint64_t hi = kMask63 - a.hi;
    return fp_from_words(lo, hi);
}
// core/field.hpp:69-71
PV_HD Fp fp_sub(const Fp& a, const Fp& b) { return fp_add(a, fp_neg(b)); }

// value of a 320-bit accumulator mod p: 2^128 = 2, 2^256 = 4
PV_HD Fp fp_wide_reduce(const uint64_t acc[5]) {
    const Fp x = fp_from_words(acc[0], acc[1]), y = fp_from_words(acc[2], acc[3]), z = fp_from_words(acc[4], 0);
    const Fp y2 = fp_add(y, y), z2 = fp_add(z, z);
    return fp_add(fp_add(x, y2), fp_add(z2, z2));
}

PV_HD void mul64wide(uint64_t a, uint64_t b, uint64_t& lo, uint64_t& hi) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    lo = (uint64_t)p;
    hi = (uint64_t)(p >> 64);
#endif
}

// core/field.hpp:113-213 : 2x2 schoolbook product (256 bits), Mersenne folds, canonicalise. Inputs must be canonical (< 2^127),
// which every value produced by this engine is. Written on unsigned __int128 (nvcc lowers it to IMAD.WIDE.U32 carry chains: 87
// instructions against 110 for the round-1 form with explicit carry compares). With a = a0 + 2^64 a1, b likewise and 2^127 = 1:
//   a b = p00 + 2^64 mid + 2^128 p11,  mid = a0 b1 + a1 b0 < 2^128 (the high limbs are < 2^63)
//       = (p00 mod 2^127) + (p00 >> 127) + 2^64 (mid_lo mod 2^63) + (mid_lo >> 63) + 2 mid_hi + 2 p11   (mod p)
typedef unsigned __int128 pv_u128;
PV_HD Fp fp_mul(const Fp& a, const Fp& b) {
    const pv_u128 M127 = (((pv_u128)kMask63) << 64) | ~0ull;
    const pv_u128 p00 = (pv_u128)a.lo * b.lo, p11 = (pv_u128)a.hi * b.hi;
    const pv_u128 mid = (pv_u128)a.lo * b.hi + (pv_u128)a.hi * b.lo;
    const uint64_t ml = (uint64_t)mid, mh = (uint64_t)(mid >> 64);
    pv_u128 s1 = (p00 & M127) + ((pv_u128)(ml & kMask63) << 64);                    // < 2^128
    pv_u128 s2 = ((p11 + mh) << 1) + (uint64_t)(p00 >> 127) + (ml >> 63);           // < 2^127 + 2^65 + 2
    s1 = (s1 & M127) + (uint64_t)(s1 >> 127);                                       // <= 2^127 - 1
    s2 = (s2 & M127) + (uint64_t)(s2 >> 127);                                       // <= 2^127
    const pv_u128 r = s1 + s2;                                                      // < 2^128: fold bit 127 and canonicalise
    return fp_from_words((uint64_t)r, (uint64_t)(r >> 64));
}

// acc (320 bits, little-endian 64-bit limbs) += a * b as a plain 256-bit integer product: no reduction at all. dec_value's edge sum
// multiplies every edge by the (reduced) inverse of its layer key this way and reduces once per thread.
PV_HD void fp_mac_wide(uint64_t acc[5], const Fp& a, const Fp& b) {
    const pv_u128 p00 = (pv_u128)a.lo * b.lo, p01 = (pv_u128)a.lo * b.hi, p10 = (pv_u128)a.hi * b.lo, p11 = (pv_u128)a.hi * b.hi;
    pv_u128 t = (pv_u128)acc[0] + (uint64_t)p00;
    acc[0] = (uint64_t)t;
    t = (t >> 64) + acc[1] + (uint64_t)(p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    acc[1] = (uint64_t)t;
    t = (t >> 64) + acc[2] + (uint64_t)(p01 >> 64) + (uint64_t)(p10 >> 64) + (uint64_t)p11;
    acc[2] = (uint64_t)t;
    t = (t >> 64) + acc[3] + (uint64_t)(p11 >> 64);
    acc[3] = (uint64_t)t;
    acc[4] += (uint64_t)(t >> 64);
}
PV_HD Fp fp_sqr_n(Fp a, int n) {
    for (int i = 0; i < n; i++) a = fp_mul(a, a);
    return a;
}

// core/field.hpp:229-273 computes a^(p-2) with a 5-bit window; the value is unique, so a shorter addition chain is
// used here: p-2 = 2^127-3 = 4*(2^125-1)+1.
PV_HD Fp fp_inv(const Fp& a) {
    Fp x1 = a;
    Fp x2 = fp_mul(fp_sqr_n(x1, 1), x1);
    Fp x4 = fp_mul(fp_sqr_n(x2, 2), x2);
    Fp x5 = fp_mul(fp_sqr_n(x4, 1), x1);
    Fp x10 = fp_mul(fp_sqr_n(x5, 5), x5);
    Fp x20 = fp_mul(fp_sqr_n(x10, 10), x10);
    Fp x25 = fp_mul(fp_sqr_n(x20, 5), x5);
    Fp x50 = fp_mul(fp_sqr_n(x25, 25), x25);
    Fp x100 = fp_mul(fp_sqr_n(x50, 50), x50);
    Fp x125 = fp_mul(fp_sqr_n(x100, 25), x25);
    return fp_mul(fp_sqr_n(x125, 2), x1);
}

// crypto/lpn.hpp:25-37
PV_HD Fp hash_to_fp_nonzero(uint64_t lo, uint64_t hi) {
    Fp r = fp_from_words(lo, hi & kMask63);
    if (fp_is_zero(r)) return fp_one();
    return r;
}

// core/types.hpp:145-155 : (lo word, then hi word) per try, retry on zero
PV_HD Fp rand_fp_nonzero(Tape& t) {
    for (;;) {
        uint64_t lo = t.next();
        uint64_t hi = t.next() & kMask63;
        Fp x = fp_from_words(lo, hi);
        if (!fp_is_zero(x)) return x;
    }
}

}  // namespace pvacb
