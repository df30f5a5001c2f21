// TEST INFRASTRUCTURE ONLY (part of the oracle; see pvac_oracle.c).
// libstdc++ dependency of ct_mul's emission order (ops/arithmetic.hpp:75-76): bucket count of an empty
// std::unordered_map after reserve(n). libstdc++ (GCC 13.3.0) computes it as
// _Prime_rehash_policy::_M_next_bkt(max(ceil(n / 1.0), 1)); asking the policy object directly avoids
// allocating n buckets (the reference itself dies with bad_alloc for n ~ 3e10).
#include <cstdint>
#include <unordered_map>

extern "C" uint64_t orc_next_bkt(uint64_t n) {
    std::__detail::_Prime_rehash_policy pol;  // max_load_factor 1.0
    std::size_t want = pol._M_bkt_for_elements(n);
    if (want < 1) want = 1;
    return (uint64_t)pol._M_next_bkt(want);
}

// cross-check used by tests at small n: the real container
extern "C" uint64_t orc_unordered_buckets_real(uint64_t n) {
    std::unordered_map<uint64_t, int> m;
    m.reserve(n);
    return m.bucket_count();
}
