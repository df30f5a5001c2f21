// TEST INFRASTRUCTURE. Force-included (g++ -include) in front of an UNMODIFIED program of the reference: the reference's headers are
// compiled under their own names first, then the entry points the program calls are renamed to the GPU binding (pvac_gpu.hpp), so
// every keygen / enc_value / ct_add / ct_sub / ct_mul / dec_value / commit_ct / enc_text / dec_text of the program runs on the B200
// through libpvacb.so. The program's own `#include <pvac/pvac.hpp>` is a no-op afterwards (#pragma once).
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <iomanip>
#include <iostream>
#include <random>
#include <string>
#include <vector>
#include "pvac_gpu.hpp"
#define keygen gpu_keygen
#define enc_value gpu_enc_value
#define ct_add gpu_ct_add
#define ct_sub gpu_ct_sub
#define ct_mul gpu_ct_mul
#define dec_value gpu_dec_value
#define commit_ct gpu_commit_ct
#define enc_text gpu_enc_text
#define dec_text gpu_dec_text
