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

// core/field.hpp:113-213 : 2x2 schoolbook product (256 bits), two Mersenne folds, canonicalise.
// Inputs must be canonical (< 2^127), which every value produced by this engine is.
PV_HD Fp fp_mul(const Fp& a, const Fp& b) {
    uint64_t p00l, p00h, p01l, p01h, p10l, p10h, p11l, p11h;
    mul64wide(a.lo, b.lo, p00l, p00h);
    mul64wide(a.lo, b.hi, p01l, p01h);
    mul64wide(a.hi, b.lo, p10l, p10h);
    mul64wide(a.hi, b.hi, p11l, p11h);
    uint64_t z0 = p00l;
    // z1 = p00h + p01l + p10l (carry c1 in 0..2)
    uint64_t z1 = p00h + p01l;
    uint64_t c1 = (z1 < p00h) ? 1ull : 0ull;
    uint64_t t = z1 + p10l;
    c1 += (t < z1) ? 1ull : 0ull;
    z1 = t;
    // z2 = p01h + p10h + p11l + c1
    uint64_t z2 = p01h + p10h;
    uint64_t c2 = (z2 < p01h) ? 1ull : 0ull;
    t = z2 + p11l;
    c2 += (t < z2) ? 1ull : 0ull;
    z2 = t;
    t = z2 + c1;
    c2 += (t < z2) ? 1ull : 0ull;
    z2 = t;
    uint64_t z3 = p11h + c2;
    // value = L + 2^127 * H with L = low 127 bits, H = bits 127..253 (< 2^127 for canonical inputs)
    uint64_t l0 = z0, l1 = z1 & kMask63;
    uint64_t h0 = (z1 >> 63) | (z2 << 1);
    uint64_t h1 = (z2 >> 63) | (z3 << 1);
    uint64_t s0 = l0 + h0;
    uint64_t s1 = l1 + h1 + ((s0 < l0) ? 1ull : 0ull);   // < 2^64: l1,h1 < 2^63
    return fp_from_words(s0, s1);                          // L + H < 2^128 ; fold bit 127 and canonicalise
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
