// Config 1 of BASELINE.json through the batched C++ API (include/pvacb.hpp): the arithmetic scenarios of the reference's
// examples/basic_usage.cpp (all of its sections except x^16, which needs > 62 GB in the reference itself), each run on a
// batch of LANES independent lanes at once instead of one ciphertext. Lane 0 carries the reference's own operand values.
//
//   g++ -std=c++17 -O2 -I include tests/cpp/basic_usage_batched.cpp -L pvac_hfhe_cppbyv_b200 -lpvacb -Wl,-rpath,... -o basic_usage_batched
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "pvacb.hpp"

using pvacb::Ciphers;
using pvacb::Engine;
using u64 = uint64_t;
using u128 = unsigned __int128;

static int g_pass = 0, g_fail = 0, g_test = 0;
static const u128 P = (((u128)1) << 127) - 1;
static u64 g_seed = 1000;

static void section(const char* name) { std::printf("\n - %d. %s -\n", ++g_test, name); }
static void check(bool ok, const std::string& msg) {
    std::printf("   [%s] %s\n", ok ? "ok" : "FAIL", msg.c_str());
    ok ? ++g_pass : ++g_fail;
}
static u128 val(const pvacb::Fp& f) { return ((u128)f.hi << 64) | f.lo; }
static bool all_eq(const std::vector<pvacb::Fp>& d, const std::vector<u128>& want) {
    if (d.size() != want.size()) return false;
    for (size_t i = 0; i < d.size(); i++)
        if (val(d[i]) != want[i] % P) return false;
    return true;
}
template <class F>
static std::vector<u128> lanes(size_t n, F f) { std::vector<u128> v(n); for (size_t i = 0; i < n; i++) v[i] = f(i); return v; }

int main() {
    constexpr size_t LANES = 3;
    Engine eng(0, PVACB_PRF_LIVE);
    section("keygen");
    eng.keygen(1);
    check(true, "keys on the device");

    const std::vector<u64> a = {42, 1000003, 0xFFFFFFFFull}, b = {17, 999, 0xFFFFFFFFFFFFFFFFull}, c = {5, 31337, 2};
    auto enc = [&](const std::vector<u64>& v) { return eng.enc_value(v, g_seed++); };
    auto mul = [&](const Ciphers& x, const Ciphers& y) { return eng.ct_mul(x, y, g_seed++); };
    auto dec = [&](const Ciphers& x) { return eng.dec_value(x); };
    auto A = [&](size_t i) { return (u128)a[i]; };
    auto B = [&](size_t i) { return (u128)b[i]; };
    auto Cc = [&](size_t i) { return (u128)c[i]; };

    section("enc / dec");
    Ciphers ca = enc(a), cb = enc(b), cc = enc(c);
    check(all_eq(dec(ca), lanes(LANES, A)), "dec(enc(a)) = a");
    check(all_eq(dec(cb), lanes(LANES, B)), "dec(enc(b)) = b");

    section("zero / one");
    Ciphers c0 = enc(std::vector<u64>(LANES, 0)), c1 = enc(std::vector<u64>(LANES, 1));
    check(all_eq(dec(c0), std::vector<u128>(LANES, 0)), "dec(0) = 0");
    check(all_eq(dec(c1), std::vector<u128>(LANES, 1)), "dec(1) = 1");

    section("x + 0 = x");
    check(all_eq(dec(eng.ct_add(ca, c0)), lanes(LANES, A)), "a + 0 = a");
    section("x * 1 = x");
    check(all_eq(dec(mul(ca, c1)), lanes(LANES, A)), "a * 1 = a");
    section("x * 0 = 0");
    check(all_eq(dec(mul(ca, c0)), std::vector<u128>(LANES, 0)), "a * 0 = 0");
    section("x - x = 0");
    check(all_eq(dec(eng.ct_sub(ca, ca)), std::vector<u128>(LANES, 0)), "a - a = 0");

    section("commut");
    check(dec(eng.ct_add(ca, cb)) == dec(eng.ct_add(cb, ca)), "a + b = b + a");
    check(dec(mul(ca, cb)) == dec(mul(cb, ca)), "a * b = b * a");

    section("assoc");
    check(dec(eng.ct_add(eng.ct_add(ca, cb), cc)) == dec(eng.ct_add(ca, eng.ct_add(cb, cc))), "(a + b) + c = a + (b + c)");
    check(dec(mul(mul(ca, cb), cc)) == dec(mul(ca, mul(cb, cc))), "(a * b) * c = a * (b * c)");

    section("distrib");
    {
        auto left = dec(mul(ca, eng.ct_add(cb, cc)));
        auto right = dec(eng.ct_add(mul(ca, cb), mul(ca, cc)));
        check(left == right && all_eq(left, lanes(LANES, [&](size_t i) { return A(i) * ((B(i) + Cc(i)) % P) % P; })), "a * (b + c) = a*b + a*c");
    }

    section("(a + b)^2 = a^2 + 2ab + b^2");
    Ciphers c_apb = eng.ct_add(ca, cb);
    Ciphers c_a_sq = mul(ca, ca), c_b_sq = mul(cb, cb);
    {
        Ciphers ab = mul(ca, cb);
        Ciphers rhs = eng.ct_add(eng.ct_add(c_a_sq, eng.ct_add(ab, ab)), c_b_sq);
        check(dec(mul(c_apb, c_apb)) == dec(rhs), "lhs = rhs on every lane");
    }

    section("(a - b)(a + b) = a^2 - b^2");
    check(dec(mul(eng.ct_sub(ca, cb), c_apb)) == dec(eng.ct_sub(c_a_sq, c_b_sq)), "lhs = rhs on every lane");

    section("poly f(x) = x^3 + 2x^2 + 3x + 4");
    {
        const std::vector<u64> x = {5, 11, 1u << 20};
        Ciphers cx = enc(x), c2 = enc(std::vector<u64>(LANES, 2)), c3 = enc(std::vector<u64>(LANES, 3)), c4 = enc(std::vector<u64>(LANES, 4));
        Ciphers cx2 = mul(cx, cx), cx3 = mul(cx2, cx);
        Ciphers poly = eng.ct_add(eng.ct_add(eng.ct_add(cx3, mul(c2, cx2)), mul(c3, cx)), c4);
        check(all_eq(dec(poly), lanes(LANES, [&](size_t i) { u128 v = x[i]; return (v * v % P * v + 2 * v * v + 3 * v + 4) % P; })), "f(x) on every lane (lane 0: f(5) = 194)");
    }

    section("depth x^8");
    {
        Ciphers x1 = enc({2, 3});
        Ciphers x2 = mul(x1, x1), x4 = mul(x2, x2), x8 = mul(x4, x4);
        check(all_eq(dec(x8), {256, 6561}), "2^8 = 256, 3^8 = 6561");
        std::printf("   edges (both lanes): x^1 = %llu, x^2 = %llu, x^4 = %llu, x^8 = %llu\n", (unsigned long long)x1.totals().second,
                    (unsigned long long)x2.totals().second, (unsigned long long)x4.totals().second, (unsigned long long)x8.totals().second);
    }

    section("rand 10 pairs");
    {
        std::mt19937_64 rng(12345);
        std::vector<u64> r1(10), r2(10);
        for (int i = 0; i < 10; i++) { r1[i] = rng() % 1000; r2[i] = rng() % 1000; }
        Ciphers x = enc(r1), y = enc(r2);          // the ten pairs are ONE batch here
        auto s = dec(eng.ct_add(x, y)), p = dec(mul(x, y));
        for (int i = 0; i < 10; i++) check(val(s[i]) == r1[i] + r2[i] && val(p[i]) == (u128)r1[i] * r2[i], "pair " + std::to_string(i));
    }

    section("fib(10)");
    {
        Ciphers fp = enc(std::vector<u64>(LANES, 0)), fc = enc(std::vector<u64>(LANES, 1));
        for (int i = 2; i <= 10; i++) { Ciphers fn = eng.ct_add(fp, fc); fp = std::move(fc); fc = std::move(fn); }
        check(all_eq(dec(fc), std::vector<u128>(LANES, 55)), "fib(10) = 55");
    }

    section("6!");
    {
        Ciphers fact = enc({1});
        for (u64 i = 2; i <= 6; i++) fact = mul(fact, enc({i}));
        check(all_eq(dec(fact), {720}), "6! = 720");
        std::printf("   edges = %llu, layers = %llu\n", (unsigned long long)fact.totals().second, (unsigned long long)fact.totals().first);
    }

    section("sum of sq 1..5");
    {
        Ciphers sum = enc(std::vector<u64>(LANES, 0));
        for (u64 i = 1; i <= 5; i++) { Ciphers ci = enc(std::vector<u64>(LANES, i)); sum = eng.ct_add(sum, mul(ci, ci)); }
        check(all_eq(dec(sum), std::vector<u128>(LANES, 55)), "1 + 4 + 9 + 16 + 25 = 55");
    }

    section("nested ((a + b) * c - a) * b");
    {
        Ciphers va = enc({3, 30}), vb = enc({5, 50}), vc = enc({7, 70});
        Ciphers nest = mul(eng.ct_sub(mul(eng.ct_add(va, vb), vc), va), vb);
        check(all_eq(dec(nest), {((3 + 5) * 7 - 3) * 5, ((30 + 50) * 70 - 30) * 50}), "((3 + 5) * 7 - 3) * 5 = 265");
    }

    section("diff ct same val");
    {
        Ciphers x = enc({100}), y = enc({100});
        check(dec(x) == dec(y), "both = 100");
        check(eng.to_wire(x) != eng.to_wire(y), "diff rnd");
    }

    section("commit uniq");
    {
        auto cm = eng.commit_ct(enc({100, 100}));
        check(cm[0] != cm[1], "diff ct -> diff commit");
    }

    // the four text sections of the reference as ONE batch of messages
    section("text ascii / special / utf8 / empty");
    {
        const std::vector<std::string> msgs = {"Hello, pvac-hfhe! This is a test of the text encryption.", "!@#$%^&*()_+-=[]{}|;':\",./<>?`~ \t\n", "\xd0\x9f\xd1\x80\xd0\xb8\xd0\xb2\xd0\xb5\xd1\x82 \xe4\xbd\xa0\xe5\xa5\xbd \xf0\x9f\x94\x90", ""};
        auto back = eng.dec_text(eng.enc_text(msgs, g_seed++), msgs.size());
        check(back[0] == msgs[0], "ascii roundtrip");
        check(back[1] == msgs[1], "special roundtrip");
        check(back[2] == msgs[2], "utf8 roundtrip");
        check(back[3] == msgs[3], "empty roundtrip");
    }

    section("perf 100 adds");
    {
        Ciphers sum = enc({0});
        for (u64 i = 0; i < 100; i++) sum = eng.ct_add(sum, enc({i}));
        check(all_eq(dec(sum), {4950}), "sum(0..99) = 4950");
    }

    section("perf 10 muls");
    {
        Ciphers prod = enc({1});
        Ciphers two = enc({2});
        for (int i = 0; i < 10; i++) prod = mul(prod, two);
        check(all_eq(dec(prod), {1024}), "2^10 = 1024");
        std::printf("   edges = %llu, layers = %llu\n", (unsigned long long)prod.totals().second, (unsigned long long)prod.totals().first);
    }

    section("large val");
    check(all_eq(dec(enc({123456789, 0xFFFFFFFFFFFFFFFFull})), {123456789, 0xFFFFFFFFFFFFFFFFull}), "enc / dec 123456789 and 2^64-1");

    section("wire round trip");
    {
        auto buf = eng.to_wire(ca);
        Ciphers back = eng.from_wire(buf);
        check(eng.to_wire(back) == buf && dec(back) == dec(ca), "export -> import -> export is the identity");
    }

    section("recrypt (tests/test_main.cpp:282-306)");
    {
        Ciphers pool = eng.enc_zero_depth(32, 3, g_seed++);            // make_evalkey(pk, sk, 32, 3): the zero pool
        check(all_eq(dec(pool), std::vector<u128>(32, 0)), "the pool encrypts zeros");
        Ciphers x = enc({7, 1000003});
        Ciphers x3 = mul(mul(x, x), x);
        Ciphers u = eng.ct_recrypt(x3, pool, g_seed++);
        check(dec(u) == dec(x3), "recrypt keeps the plaintext");
        auto d = eng.sigma_density(u);
        check(d[0] > 0.45 && d[0] < 0.55 && d[1] > 0.45 && d[1] < 0.55, "sigma density stays near 1/2");
        Ciphers chain = enc({2});
        int rec = 0;
        for (int i = 1; i < 10; i++) {
            chain = mul(chain, enc({2}));
            if (i % 3 == 0) { chain = eng.ct_recrypt(chain, pool, g_seed++); rec++; }
        }
        check(all_eq(dec(chain), {1024}) && rec == 3, "2^10 with three recrypts");
    }

    section("slices and concatenation");
    {
        Ciphers all = enc({1, 2, 3, 4, 5});
        Ciphers lo = eng.slice(all, 0, 2), hi = eng.slice(all, 2, 3);
        Ciphers back = eng.concat({&lo, &hi});
        check(eng.to_wire(back) == eng.to_wire(all), "slice + concat is the identity");
        const std::vector<pvacb::Fp> fv = {pvacb::Fp{5, 0}, pvacb::Fp{~0ull, 0x7FFFFFFFFFFFFFFEull}};
        Ciphers z = eng.enc_fp_depth(fv, 2, g_seed++);
        check(dec(z) == fv, "enc_fp_depth round trip");
    }

    std::printf("\npassed %d/%d\n", g_pass, g_pass + g_fail);
    return g_fail ? 1 : 0;
}
