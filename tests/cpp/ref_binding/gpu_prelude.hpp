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
#define ct_scale gpu_ct_scale
#define ct_neg gpu_ct_neg
#define ct_div_const gpu_ct_div_const
#define compact_edges gpu_compact_edges
#define ubk_apply gpu_ubk_apply
#define sigma_density gpu_sigma_density
#define enc_value_depth gpu_enc_value_depth
#define enc_fp_depth gpu_enc_fp_depth
#define enc_zero_depth gpu_enc_zero_depth
#define make_evalkey gpu_make_evalkey
#define ct_recrypt gpu_ct_recrypt
