// Shared helpers for the evreal_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/evreal_b200.h"

namespace evk {

void set_error(const char* fmt, ...);

#define EVK_CHECK_CUDA(expr)                                                         \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            evk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                      \
            return EVK_ERR_CUDA;                                                     \
        }                                                                            \
    } while (0)

#define EVK_REQUIRE(cond, code, ...)      \
    do {                                  \
        if (!(cond)) {                    \
            evk::set_error(__VA_ARGS__);  \
            return (code);                \
        }                                 \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;   // B200

// host: output channels [n0, n1) of a packer in parallel (weights of a whole network are ~10 M values: the packers run once per
// (model, shape) inside the first forward, i.e. inside the measured wall clock of evaluate())
template <typename F>
static inline void parallel_channels(int cout, F&& body, size_t work = (size_t)1 << 30) {
    // (threads are created per call: one per ~32 k weights, so that small layers stay on the calling thread)
    const int nt = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 16, cout / 16, (int)(work >> 15)}));
    if (nt <= 1) { body(0, cout); return; }
    std::vector<std::thread> th;
    const int per = ((cout + nt - 1) / nt + 15) / 16 * 16;
    for (int n0 = 0; n0 < cout; n0 += per) {
        const int n1 = std::min(cout, n0 + per);
        th.emplace_back([&body, n0, n1] { body(n0, n1); });
    }
    for (auto& t : th) t.join();
}

// Bump allocator over a few large device chunks (zero-initialised, 1 kB aligned): a network's device program is ~300 buffers, and
// one cudaMalloc / cudaFree each made building and -- above all -- destroying a handle cost 0.3-0.7 s of driver calls (cudaFree
// synchronises the device), inside the wall clock of evaluate().  Chunks grow from 32 MB to 512 MB.
struct DeviceArena {
    std::vector<void*> chunks;
    char* cur = nullptr;
    size_t left = 0, next = (size_t)32 << 20;
    void* alloc(size_t bytes) {
        bytes = std::max<size_t>((bytes + 1023) & ~(size_t)1023, 1024);
        if (bytes > left) {
            const size_t csz = std::max(bytes, next);
            void* p = nullptr;
            if (cudaMalloc(&p, csz) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            cudaMemset(p, 0, csz);
            chunks.push_back(p);
            cur = static_cast<char*>(p); left = csz;
            next = std::min(next * 2, (size_t)512 << 20);
        }
        void* r = cur;
        cur += bytes; left -= bytes;
        return r;
    }
    void release() {
        for (void* p : chunks) cudaFree(p);
        chunks.clear(); cur = nullptr; left = 0;
    }
};

// ---- activations used by conv epilogues
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2, ACT_TANH = 3 };

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_SIGMOID: return sigmoidf_(v);
        case ACT_TANH: return tanhf(v);
        default: return v;
    }
}

}  // namespace evk
