// Microbenchmark: L2 reduction throughput for the voxelizer's scatter: per "event" either two scalar red.add.f32 to
// different planes (current kernel), two scalar reds into one 32-byte pixel record, one red.v2.f32, or one red.v4.f32.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bench red_bench.cu && ./red_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* grid, int pixels, long long n_events) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_events; i += stride) {
        const int pix = rng(s) % pixels;
        const int b = rng(s) & 3;                 // lower bin 0..3 (5 bins)
        const float w = 0.25f;
        if (MODE == 0) {                          // planar [5][pixels]: two scalar reds, different planes
            atomicAdd(grid + (size_t)b * pixels + pix, w);
            atomicAdd(grid + (size_t)(b + 1) * pixels + pix, 1.f - w);
        } else if (MODE == 1) {                   // interleaved [pixels][8]: two scalar reds, same sector
            atomicAdd(grid + (size_t)pix * 8 + b, w);
            atomicAdd(grid + (size_t)pix * 8 + b + 1, 1.f - w);
        } else if (MODE == 2) {                   // interleaved, one red.v2 (slots 2*(b>>1).. : aligned pair; illustrative)
            float* p = grid + (size_t)pix * 8 + (b & ~1);
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(w), "f"(1.f - w) : "memory");
        } else if (MODE == 3) {                   // interleaved, one red.v4 on the aligned 16-byte half holding the pair
            float* p = grid + (size_t)pix * 8 + (b == 3 ? 4 : 0);
            float v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (b == 0) { v0 = w; v1 = 1.f - w; } else if (b == 1) { v1 = w; v2 = 1.f - w; }
            else if (b == 2) { v2 = w; v3 = 1.f - w; } else { v0 = w; v1 = 1.f - w; }
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
        } else if (MODE == 4) {                   // one scalar red per event (upper bound for 1 request/event)
            atomicAdd(grid + (size_t)pix * 8 + b, w);
        }
    }
}

template <int MODE>
void run(const char* name, float* grid, int pixels, long long n) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(grid, pixels, n);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<148 * 8, 256>>>(grid, pixels, n);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-70s %8.1f us per 4M events  -> %6.1f Gevents/s  (%s)\n", name, ms / 5 * 1e3, n / (ms / 5 * 1e-3) / 1e9, cudaGetErrorString(e));
}

int main() {
    const int pixels = 640 * 480;
    float* grid;
    cudaMalloc(&grid, sizeof(float) * pixels * 8);
    cudaMemset(grid, 0, sizeof(float) * pixels * 8);
    const long long n = 4000000;
    run<0>("planar, 2 scalar red.f32 (current voxelizer)", grid, pixels, n);
    run<1>("interleaved [pix][8], 2 scalar red.f32 in one sector", grid, pixels, n);
    run<2>("interleaved, 1 red.v2.f32", grid, pixels, n);
    run<3>("interleaved, 1 red.v4.f32", grid, pixels, n);
    run<4>("interleaved, 1 scalar red.f32", grid, pixels, n);
    return 0;
}
