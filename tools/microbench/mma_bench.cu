// Microbenchmark: tcgen05.mma (kind::f16, bf16 -> fp32, cta_group::1, M=128) throughput for the instruction sequences
// the split-bf16 convolution uses.  One thread per CTA issues ROUNDS x (KS K-slices x sequence), fully unrolled with
// constant descriptors, commits to an mbarrier and waits; clock64 around issue and around completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_bench mma_bench.cu && ./mma_bench
#include <cstdio>
#include <vector>
#include "../../evreal_b200/csrc/tc.cuh"

using namespace evk;
namespace evk { void set_error(const char*, ...) {} }

// sequence per K slice: MMA(N1 -> D0, A at 0), then if N2: MMA(N2 -> D0 + D2, A at 32 KB), then if N3: MMA(N3 -> D0 + D3, B at +16 KB)
template <int N1, int N2, int D2, int N3, int D3, int KS, int MODE = 0>
__global__ void __launch_bounds__(128, 1) bench(int rounds, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar2[i]), 1); mbar_fence_init(); }
    if (warp == 1) tc_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t i1 = umma_idesc_bf16(128, N1), i2 = umma_idesc_bf16(128, N2 ? N2 : 16), i3 = umma_idesc_bf16(128, N3 ? N3 : 16);
        const uint64_t a1 = umma_desc_kmajor(base, 128), a2 = umma_desc_kmajor(base + 32768, 128);
        const uint64_t b1 = umma_desc_kmajor(base + 65536, 128), b3 = umma_desc_kmajor(base + 65536 + 32768, 128);
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < rounds; ++it) {
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                tc_mma_bf16(tmem, a1 + 2 * k, b1 + 2 * k, i1, 1u);
                if (N2) tc_mma_bf16(tmem + D2, a2 + 2 * k, b1 + 2 * k, i2, 1u);
                if (N3) tc_mma_bf16(tmem + D3, a1 + 2 * k, b3 + 2 * k, i3, 1u);
            }
            if (MODE & 1) tc_commit(smem_u32(&bar2[it & 3]));                       // one commit per K block
            if (MODE & 2) tc_fence_after();
            if (MODE & 4) (void)mbar_try_wait(smem_u32(&bar2[(it + 1) & 3]), 0);   // a barrier poll per K block
            if (MODE & 8) { tc_commit(smem_u32(&bar2[it & 3])); tc_commit(smem_u32(&bar2[(it + 2) & 3])); }
        }
        const long long t1 = clock64();
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        out[blockIdx.x * 2 + 0] = t1 - t0;
        out[blockIdx.x * 2 + 1] = t2 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tc_dealloc(tmem, 512);
}

template <int N1, int N2, int D2, int N3, int D3, int KS, int MODE = 0>
void run(const char* name, long long* d) {
    auto k = bench<N1, N2, D2, N3, D3, KS, MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {148}) {
        const int rounds = 400;
        k<<<grid, 128, 180 * 1024>>>(rounds, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
        std::vector<long long> h(grid * 2);
        cudaMemcpy(h.data(), d, grid * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
        double issue = 0, total = 0;
        for (int i = 0; i < grid; ++i) { issue += h[2 * i]; total += h[2 * i + 1]; }
        const int len = 1 + (N2 != 0) + (N3 != 0);
        const double seqs = (double)rounds * KS;
        printf("%-52s grid %3d: issue %.1f, complete %.1f cyc per K slice (%d MMAs; tcgen05 floor %.0f)\n", name, grid,
               issue / grid / seqs, total / grid / seqs, len, (N1 + N2 + N3) / 2.0);
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 2 * sizeof(long long));
    run<256, 0, 0, 0, 0, 4>("N=256", d);
    run<128, 0, 0, 0, 0, 4>("N=128", d);
    run<64, 0, 0, 0, 0, 4>("N=64", d);
    run<32, 0, 0, 0, 0, 4>("N=32", d);
    run<16, 0, 0, 0, 0, 4>("N=16", d);
    run<96, 0, 0, 0, 0, 4>("N=96", d);
    run<192, 0, 0, 0, 0, 4>("N=192", d);
    run<256, 128, 0, 0, 0, 4>("[bn=128] N=256 D0 / N=128 D0", d);
    run<256, 128, 256, 0, 0, 4>("[bn=128 sep] N=256 D0 / N=128 D256", d);
    run<128, 128, 0, 128, 0, 4>("[3 pass same D] N=128 x3", d);
    run<128, 128, 128, 128, 256, 4>("[3 pass 3 D] N=128 x3", d);
    run<128, 64, 0, 0, 0, 4>("[bn=64] N=128 D0 / N=64 D0", d);
    run<128, 64, 128, 0, 0, 4>("[bn=64 sep] N=128 D0 / N=64 D128", d);
    run<64, 32, 0, 0, 0, 4>("[bn=32] N=64 D0 / N=32 D0", d);
    run<64, 32, 64, 0, 0, 4>("[bn=32 sep] N=64 D0 / N=32 D64", d);
    run<256, 128, 0, 0, 0, 1>("[bn=128, KS=1] N=256 D0 / N=128 D0", d);
    run<256, 0, 0, 0, 0, 8>("N=256 KS=8", d);
    run<256, 128, 0, 0, 0, 4, 1>("[bn=128] + commit per K block (4 slices)", d);
    run<256, 128, 0, 0, 0, 4, 2>("[bn=128] + fence::after per K block", d);
    run<256, 128, 0, 0, 0, 4, 4>("[bn=128] + try_wait per K block", d);
    run<256, 128, 0, 0, 0, 4, 7>("[bn=128] + commit + fence + try_wait", d);
    run<256, 128, 0, 0, 0, 4, 8>("[bn=128] + 2 commits per K block", d);
    run<256, 128, 0, 0, 0, 8, 1>("[bn=128] + commit per 8 slices", d);
    run<256, 128, 0, 0, 0, 2, 1>("[bn=128] + commit per 2 slices", d);
    run<64, 32, 0, 0, 0, 4, 1>("[bn=32] + commit per K block (4 slices)", d);
    run<64, 32, 0, 0, 0, 8, 1>("[bn=32] + commit per 8 slices", d);
    run<128, 64, 0, 0, 0, 4, 1>("[bn=64] + commit per K block (4 slices)", d);
    run<256, 0, 0, 0, 0, 4, 1>("N=256 + commit per 4", d);
    return 0;
}
