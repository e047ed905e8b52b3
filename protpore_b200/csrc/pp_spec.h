// pp_spec.h -- model-specialised kernel modules (see pp_codegen.cpp / pp_spec.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pp_lowering.h"

namespace pp {

struct SpecModule {
    typedef int (*info_fn)(int*, int);
    typedef int (*set_params_fn)(const double*, int, void*);
    typedef int (*viterbi_fn)(int, const void*, const int64_t*, const int32_t*, const int32_t*, const int64_t*, double*,
                              int32_t*, int32_t*, uint32_t*, int, void*);
    typedef int (*forward_fn)(int, const void*, const int64_t*, const int32_t*, const int32_t*, const int64_t*, double*,
                              int, void*);
    void* handle = nullptr;
    info_fn info = nullptr;
    set_params_fn set_params = nullptr;
    viterbi_fn viterbi = nullptr;
    forward_fn forward = nullptr;
    std::string hash;
    bool loaded() const { return handle != nullptr; }
};

std::string generate_spec_header(const Lowered& L);
std::vector<double> spec_param_table(const Lowered& L);
uint64_t fnv1a64(const std::string& s, uint64_t h);
std::string describe_schedule(const Lowered& L);

// Finds or builds <cache_dir>/pp_spec_<hash>.so for this schedule.  Returns "" or an error message.
std::string spec_build(const Lowered& L, const std::string& src_dir, const std::string& cache_dir, std::string* so_path,
                       std::string* hash_out);
std::string spec_load(const std::string& so_path, const Lowered& L, SpecModule* mod);
void spec_unload(SpecModule* mod);

}  // namespace pp
