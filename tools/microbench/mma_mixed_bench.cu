// Microbenchmark 3: tensor-pipe time of the operand decompositions of conv_tc.cu with the operands resident in shared memory
// (no TMA, no epilogue, no barrier waits): cycles per 32-byte K step of a 128-row tile for
//   pattern 0  bf16x3      A_hi*[B_hi;B_lo] (N = 2bn, kind::f16)  +  A_lo*B_hi (N = bn, kind::f16)
//   pattern 1  mixed       A16*B16 (N = bn, kind::f16, K = 16)    +  A8*B8 (N = bn, kind::f8f6f4, K = 32)
//   pattern 2  f16 only    one kind::f16 MMA of N = bn per step
//   pattern 3  f8 only     one kind::f8f6f4 MMA of N = bn per step
//   pattern 4  mixed, grouped: the four kind::f16 MMAs of a K block first, then its four kind::f8f6f4 MMAs
// with one issuer, or two issuers on their own accumulators (free-running) -- is the pipe time additive over kinds, and what
// does alternating the kind cost?
#include <cstdio>
#include <vector>
#include "../../evreal_b200/csrc/tc.cuh"
using namespace evk;
namespace evk { void set_error(const char*, ...) {} }

struct Args { int bn, pattern, issuers, blocks, b_stages, a_atoms, waits, readers, data; };   // data: 0 zeros, 1 random finite operands     // waits: 0 none, 1 full-barrier wait + fence per block, 2 + look-ahead poll, 3 wait only (no fence), 4 fence only

__global__ void __launch_bounds__(256, 1) mixed_bench(Args a, long long* out) {
    constexpr uint32_t ROW_BYTES = 128, ATOM = 1024;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
        // random operands: every 16-bit half is a finite fp16 / bf16 (exponent field masked to stay small) and every byte a finite fp8
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem_raw)[i] = a.data ? (h & 0xB7B7B7B7u) : 0u;
    }
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); }
    if (warp == 2) tc_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = slot;
    if (threadIdx.x == 0) mbar_arrive(smem_u32(&bars[6]));
    __syncthreads();
    const uint32_t a_plane = (uint32_t)a.a_atoms * ATOM, a_stage = 2 * a_plane;
    const uint32_t b_plane = (uint32_t)a.bn * ROW_BYTES, b_stage = 2 * b_plane;
    const uint32_t smem_b = base + 2 * a_stage;
    if (warp < a.issuers) {
        const int role = warp;
        const uint32_t bn = (uint32_t)a.bn;
        const uint32_t id_bf2 = umma_idesc_bf16(128, 2 * bn), id_bf1 = umma_idesc_bf16(128, bn);
        const uint32_t id_f16 = umma_idesc_f16(128, bn), id_f8 = umma_idesc_f8_e5m2_e4m3(128, bn);
        const uint32_t desc_hi = (uint32_t)(umma_desc_kmajor(0, ROW_BYTES) >> 32);
        auto mk = [&](uint32_t lo) -> uint64_t { return ((uint64_t)desc_hi << 32) | lo; };
        auto lo_of = [](uint32_t addr) -> uint32_t { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
        const uint32_t a_plane16 = a_plane >> 4, b_plane16 = b_plane >> 4;
        const uint32_t acc_cols = a.pattern == 0 ? 2 * bn : bn;
        const uint32_t d = tmem_base + (uint32_t)role * acc_cols;
        const uint32_t bar = smem_u32(&bars[role]);
        uint32_t sB = 0;
        const long long t0 = clock64();
        for (int blk = 0; blk < a.blocks; ++blk) {
            const uint32_t al = lo_of(base + (uint32_t)(blk & 1) * a_stage + (uint32_t)(blk % 3) * ATOM);
            const uint32_t bl = lo_of(smem_b + sB * b_stage);
            if (++sB == (uint32_t)a.b_stages) sB = 0;
            // (barrier 6 completed its phase 0 before the loop: waits on parity 0 pass immediately, like a weight stage that is ready)
            if (a.waits == 1 || a.waits == 2 || a.waits == 3) mbar_wait(smem_u32(&bars[6]), 0);
            if (a.waits == 1 || a.waits == 2 || a.waits == 4) tc_fence_after();
            if (a.waits == 2) (void)mbar_try_wait(smem_u32(&bars[6]), 0);
            if (elect_one()) {
                if (a.pattern == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc_mma_bf16(d, mk(al + 2 * k), mk(bl + 2 * k), id_bf2, 1u);
                        tc_mma_bf16(d, mk(al + a_plane16 + 2 * k), mk(bl + 2 * k), id_bf1, 1u);
                    }
                } else if (a.pattern == 1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        tc_mma_bf16(d, mk(al + 2 * k), mk(bl + 2 * k), id_f16, 1u);
                        tc_mma_f8(d, mk(al + a_plane16 + 2 * k), mk(bl + b_plane16 + 2 * k), id_f8, 1u);
                    }
                } else if (a.pattern == 2) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_bf16(d, mk(al + 2 * k), mk(bl + 2 * k), id_f16, 1u);
                } else if (a.pattern == 3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f8(d, mk(al + a_plane16 + 2 * k), mk(bl + b_plane16 + 2 * k), id_f8, 1u);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_bf16(d, mk(al + 2 * k), mk(bl + 2 * k), id_f16, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f8(d, mk(al + a_plane16 + 2 * k), mk(bl + b_plane16 + 2 * k), id_f8, 1u);
                }
                tc_commit(smem_u32(&bars[4 + role]));
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(bar);
        __syncwarp();
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        if (lane == 0 && role == 0) out[blockIdx.x] = t2 - t0;
    }
    if (a.readers && warp >= 4) {
        // "epilogue" load: warps 4-7 read 32-column windows of tensor memory (columns 256.., not the accumulators) and do some math,
        // for roughly as long as the issuers run (readers = windows per block of MMAs)
        uint32_t v[32];
        float acc = 0.f;
        const uint32_t t_row = tmem_base + 256u + ((uint32_t)((warp & 3) * 32) << 16);
        for (int i = 0; i < a.blocks * a.readers; ++i) {
            tc_ld_32x32(t_row + (uint32_t)((i & 3) * 32), v);
            tc_wait_ld();
#pragma unroll
            for (int k = 0; k < 32; ++k) acc += __expf(__uint_as_float(v[k]));
        }
        if (acc == 123.456f) out[147] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, 512);
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    cudaFuncSetAttribute(mixed_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const char* names[5] = {"bf16x3 (N=2bn + N=bn)", "mixed f16+f8 alternating", "f16 only", "f8 only", "mixed, four f16 then four f8"};
    for (int data : {0, 1})
    for (int readers : {0, 1, 4})
    for (int waits = 0; waits <= 4; ++waits)
    for (int bn : {128, 256})
        for (int issuers = 1; issuers <= 2; ++issuers)
            for (int pattern = 0; pattern < 5; ++pattern) {
                if (bn == 256 && (pattern == 0 || issuers == 2)) continue;        // tensor memory: 512 columns
                if (waits > 0 && !(bn == 128 && (pattern == 0 || pattern == 1))) continue;
                if (data > 0 && (readers > 0 || waits > 0 || issuers != 2)) continue;
                if (readers > 0 && !(bn == 128 && issuers == 2 && (pattern == 0 || pattern == 1) && waits <= 1)) continue;
                Args a = {bn, pattern, issuers, 2000, bn == 256 ? 2 : 4, 18, waits, readers, data};
                mixed_bench<<<148, 256, 210 * 1024>>>(a, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s: %s\n", names[pattern], cudaGetErrorString(e)); return 1; }
                std::vector<long long> h(148);
                cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
                double tot = 0;
                for (int i = 0; i < 148; ++i) tot += h[i];
                const double steps = (double)a.blocks * 4 * issuers;
                printf("data=%d readers=%d waits=%d bn=%3d issuers=%d %-32s %.1f cyc per K step per CTA\n", data, readers, waits, bn, issuers, names[pattern], tot / 148 / steps);
            }
    return 0;
}
