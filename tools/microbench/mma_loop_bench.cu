// Microbenchmark 2: the MMA-issuer loop of conv_tc.cu in isolation (barriers always ready, no TMA, no epilogue):
// how many cycles per 16-deep K slice does the LOOP FORM cost on top of the tensor-pipe floor?
#include <cstdio>
#include <vector>
#include "../../evreal_b200/csrc/tc.cuh"
using namespace evk;
namespace evk { void set_error(const char*, ...) {} }

namespace evk {
// (experiment only: the product kernel keeps separate asm statements)
// One K block of the split-bf16 scheme as ONE asm statement: KS 16-deep slices x { A_hi*[B_hi;B_lo] (N = 2bn),
// A_lo*B_hi (N = bn) }, then the commit that frees the weight slot -- with the readiness polls of the NEXT block's
// barriers embedded after the first slice, so that their round trip (~100+ cycles even when the phase is already
// complete; measured with tools/microbench/mma_loop_bench.cu) overlaps the issue of the remaining MMAs instead of
// sitting between two K blocks as a tensor-pipe bubble.  Descriptor low words advance by 2 (= 32 bytes >> 4) per
// slice.  mask == 0: CTA-local commit, else cluster multicast.  Returns bit 0 = next weight barrier complete,
// bit 1 = next activation barrier complete.
#define EVK_MMA_HEAD                                                        \
    "{\n\t"                                                                 \
    ".reg .pred pacc, pone, pb, pa, pm;\n\t"                                \
    ".reg .b64 da, dl, db;\n\t"                                             \
    ".reg .b32 rb, ra;\n\t"                                                 \
    "setp.ne.b32 pacc, %8, 0;\n\t"                                          \
    "setp.eq.b32 pone, %8, %8;\n\t"                                         \
    "setp.ne.b16 pm, %10, 0;\n\t"                                           \
    "mov.b64 da, {%2, %1};\n\t"                                             \
    "mov.b64 dl, {%3, %1};\n\t"                                             \
    "mov.b64 db, {%4, %1};\n\t"                                             \
    "tcgen05.mma.cta_group::1.kind::f16 [%5], da, db, %6, pacc;\n\t"        \
    "tcgen05.mma.cta_group::1.kind::f16 [%5], dl, db, %7, pone;\n\t"        \
    "mbarrier.try_wait.parity.shared::cta.b64 pb, [%11], %12;\n\t"          \
    "mbarrier.try_wait.parity.shared::cta.b64 pa, [%13], %14;\n\t"
#define EVK_MMA_SLICE                                                       \
    "add.s64 da, da, 2;\n\t"                                                \
    "add.s64 dl, dl, 2;\n\t"                                                \
    "add.s64 db, db, 2;\n\t"                                                \
    "tcgen05.mma.cta_group::1.kind::f16 [%5], da, db, %6, pone;\n\t"        \
    "tcgen05.mma.cta_group::1.kind::f16 [%5], dl, db, %7, pone;\n\t"
#define EVK_MMA_TAIL                                                                                                   \
    "@pm tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], %10;\n\t"      \
    "@!pm tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"                             \
    "selp.u32 rb, 1, 0, pb;\n\t"                                                                                       \
    "selp.u32 ra, 2, 0, pa;\n\t"                                                                                       \
    "or.b32 %0, rb, ra;\n\t"                                                                                           \
    "}"
#define EVK_MMA_OPERANDS                                                                                                 \
    : "=r"(ready)                                                                                                        \
    : "r"(desc_hi), "r"(ah_lo), "r"(al_lo), "r"(bh_lo), "r"(d_tmem), "r"(idesc2), "r"(idesc1), "r"(accumulate),          \
      "r"(commit_bar), "h"(mask), "r"(next_b_bar), "r"(next_b_parity), "r"(next_a_bar), "r"(next_a_parity)               \
    : "memory"
template <int KS>
__device__ __forceinline__ uint32_t tc_mma_block(uint32_t d_tmem, uint32_t desc_hi, uint32_t ah_lo, uint32_t al_lo, uint32_t bh_lo,
                                                 uint32_t idesc2, uint32_t idesc1, uint32_t accumulate, uint32_t commit_bar,
                                                 uint16_t mask, uint32_t next_b_bar, uint32_t next_b_parity,
                                                 uint32_t next_a_bar, uint32_t next_a_parity) {
    static_assert(KS == 2 || KS == 4, "K block = 2 or 4 slices of 16");
    uint32_t ready;
    if constexpr (KS == 4)
        asm volatile(EVK_MMA_HEAD EVK_MMA_SLICE EVK_MMA_SLICE EVK_MMA_SLICE EVK_MMA_TAIL EVK_MMA_OPERANDS);
    else
        asm volatile(EVK_MMA_HEAD EVK_MMA_SLICE EVK_MMA_TAIL EVK_MMA_OPERANDS);
    return ready;
}
#undef EVK_MMA_HEAD
#undef EVK_MMA_SLICE
#undef EVK_MMA_TAIL
#undef EVK_MMA_OPERANDS
}  // namespace evk

struct Args { int bn, a_stages, b_stages, ar, ntaps, groups, blocks_per_tile, tiles, ni; };

template <int VARIANT>
__global__ void __launch_bounds__(256, 1) loop_bench(Args a, long long* out) {
    constexpr uint32_t ROW_BYTES = 128, ATOM = 1024;
    constexpr int BK = 64;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[40];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
    if (threadIdx.x == 0) { for (int i = 0; i < 40; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); }
    if (warp == 1) tc_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = slot;
    const uint32_t a_plane = (uint32_t)a.ar * ATOM, a_stage = 2 * a_plane;
    const uint32_t b_plane = (uint32_t)a.bn * ROW_BYTES, b_stage = 2 * b_plane;
    const uint32_t smem_b = base + (uint32_t)a.a_stages * a_stage;
    const uint32_t bar_fa = smem_u32(&bars[0]), bar_ea = smem_u32(&bars[8]), bar_fb = smem_u32(&bars[16]), bar_eb = smem_u32(&bars[24]);
    const uint32_t bar_done = smem_u32(&bars[39]);
    if (warp == 0 || (VARIANT >= 6 && warp >= 2 && warp < 1 + (VARIANT == 8 ? 2 : a.ni))) {
        const int role = warp == 0 ? 0 : warp - 1;
        const uint32_t idesc2 = umma_idesc_bf16(128, (uint32_t)(2 * a.bn));
        const uint32_t idesc1 = umma_idesc_bf16(128, (uint32_t)a.bn);
        const uint32_t desc_hi = (uint32_t)(umma_desc_kmajor(0, ROW_BYTES) >> 32);
        auto mk = [&](uint32_t lo) -> uint64_t { return ((uint64_t)desc_hi << 32) | lo; };
        auto lo_of = [](uint32_t addr) -> uint32_t { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
        const uint32_t atom16 = ATOM >> 4, a_plane16 = a_plane >> 4;
        uint32_t sA = 0, sB = 0;
        const long long t0 = clock64();
        if (VARIANT == 0) {
            // ---- the loop as in conv_tc.cu (barrier waits pass immediately: parity 1 of a fresh barrier)
            bool b_ready = false;
            for (int t = 0; t < a.tiles; ++t) {
                uint32_t acc = 0;
                for (int grp = 0; grp < a.groups; ++grp) {
                    uint32_t ah_lo = lo_of(base + sA * a_stage);
                    mbar_wait(bar_fa + 8u * sA, 1);
                    for (int j = 0; j < a.ntaps; ++j) {
                        const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                        const uint32_t bar_free = bar_eb + 8u * sB;
                        if (!b_ready) mbar_wait(bar_fb + 8u * sB, 1);
                        tc_fence_after();
                        if (++sB == (uint32_t)a.b_stages) sB = 0;
                        b_ready = mbar_try_wait(bar_fb + 8u * sB, 1);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                tc_mma_bf16(tmem_base, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, k == 0 ? acc : 1u);
                                tc_mma_bf16(tmem_base, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                            }
                            tc_commit(bar_free);
                        }
                        __syncwarp();
                        acc = 1u;
                        ah_lo += atom16;
                    }
                    if (elect_one()) tc_commit(bar_ea + 8u * sA);
                    __syncwarp();
                    if (++sA == (uint32_t)a.a_stages) sA = 0;
                }
            }
        } else if (VARIANT == 1) {
            // ---- single elected thread runs the whole loop; the other lanes idle at the final __syncwarp
            if (elect_one()) {
                for (int t = 0; t < a.tiles; ++t) {
                    uint32_t acc = 0;
                    for (int grp = 0; grp < a.groups; ++grp) {
                        uint32_t ah_lo = lo_of(base + sA * a_stage);
                        mbar_wait(bar_fa + 8u * sA, 1);
                        for (int j = 0; j < a.ntaps; ++j) {
                            const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                            const uint32_t bar_free = bar_eb + 8u * sB;
                            mbar_wait(bar_fb + 8u * sB, 1);
                            tc_fence_after();
                            if (++sB == (uint32_t)a.b_stages) sB = 0;
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                tc_mma_bf16(tmem_base, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, k == 0 ? acc : 1u);
                                tc_mma_bf16(tmem_base, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                            }
                            tc_commit(bar_free);
                            acc = 1u;
                            ah_lo += atom16;
                        }
                        tc_commit(bar_ea + 8u * sA);
                        if (++sA == (uint32_t)a.a_stages) sA = 0;
                    }
                }
            }
            __syncwarp();
        } else if (VARIANT == 2) {
            // ---- as VARIANT 0 but no barrier waits at all (pure loop + elect + commit)
            for (int t = 0; t < a.tiles; ++t) {
                uint32_t acc = 0;
                for (int grp = 0; grp < a.groups; ++grp) {
                    uint32_t ah_lo = lo_of(base + sA * a_stage);
                    for (int j = 0; j < a.ntaps; ++j) {
                        const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                        const uint32_t bar_free = bar_eb + 8u * sB;
                        if (++sB == (uint32_t)a.b_stages) sB = 0;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                tc_mma_bf16(tmem_base, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, k == 0 ? acc : 1u);
                                tc_mma_bf16(tmem_base, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                            }
                            tc_commit(bar_free);
                        }
                        __syncwarp();
                        acc = 1u;
                        ah_lo += atom16;
                    }
                    if (++sA == (uint32_t)a.a_stages) sA = 0;
                }
            }
        } else if (VARIANT == 4) {
            // ---- fused K-block asm with the next block's barrier polls embedded (tc_mma_block)
            uint32_t ready = 0;
            for (int t = 0; t < a.tiles; ++t) {
                uint32_t acc = 0;
                for (int grp = 0; grp < a.groups; ++grp) {
                    uint32_t ah_lo = lo_of(base + sA * a_stage);
                    if (!(ready & 2u)) mbar_wait(bar_fa + 8u * sA, 1);
                    uint32_t sA1 = sA + 1; if (sA1 == (uint32_t)a.a_stages) sA1 = 0;
                    for (int j = 0; j < a.ntaps; ++j) {
                        const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                        const uint32_t bar_free = bar_eb + 8u * sB;
                        if (!(ready & 1u)) mbar_wait(bar_fb + 8u * sB, 1);
                        tc_fence_after();
                        if (++sB == (uint32_t)a.b_stages) sB = 0;
                        if (elect_one())
                            ready = tc_mma_block<4>(tmem_base, desc_hi, ah_lo, ah_lo + a_plane16, bh_lo, idesc2, idesc1, acc, bar_free, 0,
                                                    bar_fb + 8u * sB, 1, bar_fa + 8u * sA1, 1);
                        ready = __shfl_sync(0xffffffffu, ready, 0);
                        acc = 1u;
                        ah_lo += atom16;
                    }
                    if (elect_one()) tc_commit(bar_ea + 8u * sA);
                    __syncwarp();
                    sA = sA1;
                }
            }
        } else if (VARIANT == 5) {
            // ---- as 4 but one thread runs everything (no shfl)
            if (elect_one()) {
                uint32_t ready = 0;
                for (int t = 0; t < a.tiles; ++t) {
                    uint32_t acc = 0;
                    for (int grp = 0; grp < a.groups; ++grp) {
                        uint32_t ah_lo = lo_of(base + sA * a_stage);
                        if (!(ready & 2u)) mbar_wait(bar_fa + 8u * sA, 1);
                        uint32_t sA1 = sA + 1; if (sA1 == (uint32_t)a.a_stages) sA1 = 0;
                        for (int j = 0; j < a.ntaps; ++j) {
                            const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                            const uint32_t bar_free = bar_eb + 8u * sB;
                            if (!(ready & 1u)) mbar_wait(bar_fb + 8u * sB, 1);
                            tc_fence_after();
                            if (++sB == (uint32_t)a.b_stages) sB = 0;
                            ready = tc_mma_block<4>(tmem_base, desc_hi, ah_lo, ah_lo + a_plane16, bh_lo, idesc2, idesc1, acc, bar_free, 0,
                                                    bar_fb + 8u * sB, 1, bar_fa + 8u * sA1, 1);
                            acc = 1u;
                            ah_lo += atom16;
                        }
                        tc_commit(bar_ea + 8u * sA);
                        sA = sA1;
                    }
                }
            }
            __syncwarp();
        } else if (VARIANT == 6 || VARIANT == 7) {
            // ---- TWO issuer warps alternate K blocks of the same tile / accumulator: while one is between blocks
            // (commit, barrier poll, descriptors, elect, R2UR) the other's MMAs keep the tensor pipe busy.
            // VARIANT 7 adds the (immediately passing) barrier waits of the real kernel.
            bool b_ready = false;
            for (int t = 0; t < a.tiles; ++t) {
                uint32_t acc = 0, blk = 0;
                if (role != 0) asm volatile("bar.sync 1, %0;" ::"r"(32 * a.ni) : "memory");     // block 0 (accumulate = 0) is issued first
                for (int grp = 0; grp < a.groups; ++grp) {
                    uint32_t ah_lo = lo_of(base + sA * a_stage);
                    if (VARIANT == 7) mbar_wait(bar_fa + 8u * sA, 1);
                    for (int j = 0; j < a.ntaps; ++j, ++blk) {
                        const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                        const uint32_t bar_free = bar_eb + 8u * sB;
                        const bool mine = (blk % (uint32_t)a.ni) == (uint32_t)role;
                        if (VARIANT == 7 && mine) { mbar_wait(bar_fb + 8u * sB, 1); tc_fence_after(); }
                        if (++sB == (uint32_t)a.b_stages) sB = 0;
                        if (mine) {
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    tc_mma_bf16(tmem_base, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, (k == 0 && blk == 0) ? 0u : 1u);
                                    tc_mma_bf16(tmem_base, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                                }
                                tc_commit(bar_free);
                            }
                            __syncwarp();
                            if (blk == 0) asm volatile("bar.arrive 1, %0;" ::"r"(32 * a.ni) : "memory");
                        }
                        ah_lo += atom16;
                    }
                    if (elect_one()) tc_commit(bar_ea + 8u * sA);
                    __syncwarp();
                    if (++sA == (uint32_t)a.a_stages) sA = 0;
                }
            }
            (void)b_ready;
        } else if (VARIANT == 8) {
            // ---- two issuers, ONE accumulator, strict block order enforced by a named-barrier handshake (deterministic
            // summation order): issuer r waits on barrier 1+r before block b > 0 and signals barrier 1+(1-r) after
            // issuing block b when a block b+1 follows
            const int KB = a.groups * a.ntaps;
            for (int t = 0; t < a.tiles; ++t) {
                uint32_t blk = 0;
                for (int grp = 0; grp < a.groups; ++grp) {
                    uint32_t ah_lo = lo_of(base + sA * a_stage);
                    mbar_wait(bar_fa + 8u * sA, 1);
                    for (int j = 0; j < a.ntaps; ++j, ++blk) {
                        const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                        const uint32_t bar_free = bar_eb + 8u * sB;
                        const bool mine = (blk & 1u) == (uint32_t)role;
                        if (mine) {
                            mbar_wait(bar_fb + 8u * sB, 1);
                            tc_fence_after();
                            if (blk > 0) { if (role) asm volatile("bar.sync 2, 64;" ::: "memory"); else asm volatile("bar.sync 1, 64;" ::: "memory"); }
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    tc_mma_bf16(tmem_base, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, (k == 0 && blk == 0) ? 0u : 1u);
                                    tc_mma_bf16(tmem_base, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                                }
                                tc_commit(bar_free);
                            }
                            __syncwarp();
                            if ((int)blk + 1 < KB) { if (role) asm volatile("bar.arrive 1, 64;" ::: "memory"); else asm volatile("bar.arrive 2, 64;" ::: "memory"); }
                        }
                        if (++sB == (uint32_t)a.b_stages) sB = 0;
                        ah_lo += atom16;
                    }
                    if (elect_one()) tc_commit(bar_ea + 8u * sA);
                    __syncwarp();
                    if (++sA == (uint32_t)a.a_stages) sA = 0;
                }
            }
        } else if (VARIANT == 3) {
            // ---- same operands every block (no ring), elect + commit: isolates the effect of rotating smem addresses
            for (int t = 0; t < a.tiles; ++t)
                for (int grp = 0; grp < a.groups; ++grp)
                    for (int j = 0; j < a.ntaps; ++j) {
                        const uint32_t ah_lo = lo_of(base), bh_lo = lo_of(smem_b);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                tc_mma_bf16(tmem_base, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, 1u);
                                tc_mma_bf16(tmem_base, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                            }
                            tc_commit(bar_eb);
                        }
                        __syncwarp();
                    }
        }
        const long long t1 = clock64();
        const uint32_t my_done = bar_done - 8u * role;   // bars[39 - role]
        if (elect_one()) tc_commit(my_done);
        __syncwarp();
        mbar_wait(my_done, 0);
        const long long t2 = clock64();
        if (lane == 0 && role == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tc_dealloc(tmem_base, 512);
}

template <int V>
void run(const char* name, Args a, long long* d) {
    cudaFuncSetAttribute(loop_bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    loop_bench<V><<<148, 256, 210 * 1024>>>(a, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<long long> h(296);
    cudaMemcpy(h.data(), d, 296 * sizeof(long long), cudaMemcpyDeviceToHost);
    double tot = 0;
    for (int i = 0; i < 148; ++i) tot += h[2 * i + 1];
    const double slices = (double)a.tiles * a.groups * a.ntaps * 4;
    const double floor_c = a.bn >= 128 ? 1.5 * a.bn : (a.bn == 64 ? 112 : 94);
    printf("%-70s bn=%3d: %.1f cyc per K slice (pipe alone: %.0f)\n", name, a.bn, tot / 148 / slices, floor_c);
}

int main() {
    long long* d;
    cudaMalloc(&d, 296 * sizeof(long long));
    for (int bn : {128, 64, 32}) {
        Args a = {bn, 2, 4, 18, 3, 6, 18, 20, 2};
        run<0>("V0 conv_tc loop, one issuer (waits pass immediately)", a, d);
        run<6>("V6 two issuers alternate blocks, shared accumulator, no waits (order not deterministic)", a, d);
        run<7>("V7 two issuers alternate blocks, with waits (order not deterministic)", a, d);
        run<8>("V8 two issuers, strict order via named-barrier handshake, with waits", a, d);
    }
    return 0;
}
