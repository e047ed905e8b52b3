// pp_spec.cpp -- builds, caches and loads the model-specialised kernel module (pp_spec_kernels.cu
// compiled against the schedule header pp_codegen.cpp writes).  Plain host C++: nvcc is run as a
// child process, the resulting shared object is bound with dlopen/dlsym.
#include "pp_spec.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "pp_common.h"

namespace pp {

static bool file_exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}

static std::string read_file(const std::string& p) {
    std::ifstream f(p, std::ios::binary);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

std::string spec_build(const Lowered& L, const std::string& src_dir, const std::string& cache_dir, std::string* so_path,
                       std::string* hash_out) {
    const std::string header = generate_spec_header(L);
    const std::string kernels = src_dir + "/pp_spec_kernels.cu";
    if (!file_exists(kernels)) return "kernel source not found: " + kernels;
    uint64_t h = fnv1a64(header, 0xCBF29CE484222325ULL);
    h = fnv1a64(read_file(kernels), h);
    h = fnv1a64(read_file(src_dir + "/pp_device_common.cuh"), h);
    h = fnv1a64(read_file(src_dir + "/pp_lowering.h"), h);
    const char* extra = std::getenv("PP_SPEC_NVCC_FLAGS");
    h = fnv1a64(std::string("sm_100a;v1;") + (extra ? extra : ""), h);
    char hs[32];
    std::snprintf(hs, sizeof hs, "%016llx", static_cast<unsigned long long>(h));
    if (hash_out) *hash_out = hs;
    ::mkdir(cache_dir.c_str(), 0755);
    const std::string base = cache_dir + "/pp_spec_" + hs;
    *so_path = base + ".so";
    if (file_exists(*so_path)) return "";
    {
        std::ofstream f(base + ".h");
        f << header;
        if (!f) return "cannot write " + base + ".h";
    }
    const char* nvcc_env = std::getenv("PP_NVCC");
    std::string nvcc = nvcc_env ? nvcc_env : "/usr/local/cuda/bin/nvcc";
    if (!file_exists(nvcc)) nvcc = "nvcc";
    char tmp[64];
    std::snprintf(tmp, sizeof tmp, ".tmp%d", static_cast<int>(::getpid()));
    const std::string out_tmp = *so_path + tmp;
    std::string cmd = nvcc + " -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared" +
                      " --expt-relaxed-constexpr -diag-suppress 177,550 -DPP_SPEC_HEADER='\"" + base + ".h\"' -I" + src_dir + " " +
                      (extra ? std::string(extra) + " " : std::string()) + kernels + " -o " + out_tmp + " > " + base +
                      ".log 2>&1";
    const int rc = std::system(cmd.c_str());
    if (rc != 0) return "nvcc failed (" + std::to_string(rc) + "), see " + base + ".log";
    if (std::rename(out_tmp.c_str(), so_path->c_str()) != 0) return "cannot move " + out_tmp;
    return "";
}

std::string spec_load(const std::string& so_path, const Lowered& L, SpecModule* mod) {
    void* h = ::dlopen(so_path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) return std::string("dlopen failed: ") + ::dlerror();
    SpecModule m;
    m.handle = h;
    m.info = reinterpret_cast<SpecModule::info_fn>(::dlsym(h, "pp_spec_info"));
    m.set_params = reinterpret_cast<SpecModule::set_params_fn>(::dlsym(h, "pp_spec_set_params"));
    m.viterbi = reinterpret_cast<SpecModule::viterbi_fn>(::dlsym(h, "pp_spec_viterbi"));
    m.forward = reinterpret_cast<SpecModule::forward_fn>(::dlsym(h, "pp_spec_forward"));
    if (!m.info || !m.set_params || !m.viterbi || !m.forward) { ::dlclose(h); return "specialised module lacks an entry point"; }
    int v[8] = {0};
    m.info(v, 8);
    if (v[0] != 1 || v[1] != L.m || v[2] != L.ne || v[3] != L.n_slots || v[4] != L.warps ||
        v[5] != static_cast<int>(2 * L.low_src.size() + 8 * static_cast<size_t>(L.ne)) || v[6] != L.n_ptr_words ||
        v[7] != L.ptr_bits) {
        ::dlclose(h);
        return "specialised module does not match this model";
    }
    *mod = m;
    return "";
}

void spec_unload(SpecModule* mod) {
    // the module stays mapped: CUDA keeps references to its kernels until process exit
    *mod = SpecModule();
}

}  // namespace pp
