// pp_common.h -- small host utilities shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/protpore_b200.h"

namespace pp {

void set_error(const char* fmt, ...);

#define PP_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            pp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PP_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// Guard mode (environment PP_GUARD=1; compute-sanitizer is not available on the GPU pool this was developed on): every
// device buffer gets kGuardBytes of a known pattern in front of and behind the bytes that were asked for, and
// pp_debug_check_guards() -- called by the GPU test-suite after every test -- compares them.  A kernel that writes below a
// buffer, or past the size its launcher computed, is caught at the next check.
constexpr size_t kGuardBytes = 4096;
constexpr int kGuardPattern = 0xA5;

struct GuardRegistry {
    std::mutex mu;
    std::map<void*, size_t> live;                      // base pointer -> payload bytes (guards excluded)
    static GuardRegistry& get() { static GuardRegistry g; return g; }
    static bool enabled() {
        static const bool on = [] { const char* e = std::getenv("PP_GUARD"); return e && std::atoi(e) != 0; }();
        return on;
    }
};

// Device buffer that only ever grows; owned by a model handle.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void drop() {
        if (!p) return;
        if (GuardRegistry::enabled()) {
            void* base = static_cast<char*>(p) - kGuardBytes;
            GuardRegistry& G = GuardRegistry::get();
            std::lock_guard<std::mutex> lk(G.mu);
            G.live.erase(base);
            cudaFree(base);
        } else {
            cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        drop();
        if (!GuardRegistry::enabled()) {
            cudaError_t e = cudaMalloc(&p, bytes);
            if (e == cudaSuccess) cap = bytes;
            return e;
        }
        const size_t payload = (bytes + 255) / 256 * 256;          // keeps the alignment cudaMalloc would have given
        void* base = nullptr;
        cudaError_t e = cudaMalloc(&base, payload + 2 * kGuardBytes);
        if (e != cudaSuccess) return e;
        e = cudaMemset(base, kGuardPattern, kGuardBytes);
        // the back guard starts right behind the bytes asked for, not behind the rounded payload
        if (e == cudaSuccess) e = cudaMemset(static_cast<char*>(base) + kGuardBytes + bytes, kGuardPattern, payload - bytes + kGuardBytes);
        if (e != cudaSuccess) { cudaFree(base); return e; }
        GuardRegistry& G = GuardRegistry::get();
        std::lock_guard<std::mutex> lk(G.mu);
        G.live[base] = bytes;
        p = static_cast<char*>(base) + kGuardBytes;
        cap = bytes;
        return cudaSuccess;
    }
    void release() { drop(); }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

template <typename T>
cudaError_t upload(DevBuf* b, const std::vector<T>& v, cudaStream_t s) {
    cudaError_t e = b->reserve(v.size() * sizeof(T) + 16);
    if (e != cudaSuccess) return e;
    if (v.empty()) return cudaSuccess;
    return cudaMemcpyAsync(b->p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

}  // namespace pp
