// sm_100a building blocks for the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (alloc / mma / commit / ld) as inline PTX, and the host-side tensor-map encoder.
//
// The driver entry point cuTensorMapEncodeTiled is fetched at run time through the CUDA runtime
// (cudaGetDriverEntryPoint) so the shared library does not link libcuda.so and still loads on a
// machine without a GPU (the C-ABI symbol test runs there).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "evk_common.cuh"

namespace evk {

// ------------------------------------------------------------------ host: tensor maps
// rank <= 5; dims / strides innermost first; strides in BYTES for dims 1..rank-1 (dim 0 is contiguous).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes /*32, 64 or 128*/);

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// non-blocking poll (try_wait may suspend the thread for a hardware time limit when the phase is not complete yet)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a pipeline bug becomes a trapped launch ("unspecified launch failure") instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) __trap();
    }
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

// multicast variant: the box lands at the same shared-memory offset in every CTA of `mask`, and each of those
// CTAs' mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}

// one lane of a CONVERGED warp (all 32 lanes must execute this); the same lane every time
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- thread-block clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// all threads of all CTAs of the cluster
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05
__device__ __forceinline__ void tc_alloc(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrives on the mbarrier at the same offset in every CTA of `mask` once all prior MMAs of this thread retire
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate; issued by ONE thread for the CTA.
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::f8f6f4 (8-bit float operands, K = 32 per instruction, twice the kind::f16 rate), same accumulator
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators: thread i of the warp receives lane (base+i), columns col..col+31
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// 16-column variant (registers 0..15 of v)
__device__ __forceinline__ void tc_ld_32x16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 8-column variant (registers 0..7 of v)
__device__ __forceinline__ void tc_ld_32x8(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory written by TMA with 128B / 64B / 32B swizzle: rows of `row_bytes`
// (= swizzle span), 8-row core groups `8*row_bytes` apart (SBO).  SM100 descriptor (version 1).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t smem_addr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)((8u * row_bytes) >> 4) << 32;            // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    d |= layout << 61;                                       // swizzle mode
    return d;
}
// kind::f16 instruction descriptor: BF16 x BF16 -> FP32, both operands K-major, M x N tile
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// kind::f16 with FP16 operands, and kind::f8f6f4 with A = E5M2, B = E4M3 (the "mixed" operand decomposition, conv_tc.cu)
__device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ uint32_t umma_idesc_f8_e5m2_e4m3(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// ---- "mixed" split records (consumers: MIXED kernels).  A pixel of C channels (C % 64 == 0) is 2*C bytes in either plane:
// plane 0 = fp16(v) [C] (saturating at +-65504); plane 1 = per 64-channel chunk 64 bytes e5m2(v) then 64 bytes
// e5m2(256 (v - x16)), x16 = fp16(v).
// o = element offset of the record's first channel (2-byte units, = pixel * C + c0), c0 = that channel's index in its pixel.
__device__ __forceinline__ void mixed_cvt2(float v0, float v1, uint32_t& h2, uint32_t& x2, uint32_t& l2) {
    const float s0 = fminf(fmaxf(v0, -65504.0f), 65504.0f), s1 = fminf(fmaxf(v1, -65504.0f), 65504.0f);
    const __half2 h = __floats2half2_rn(s0, s1);
    h2 = *reinterpret_cast<const uint32_t*>(&h);
    const float2 hf = __half22float2(h);
    x2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(v0, v1), __NV_SATFINITE, __NV_E5M2);
    l2 = (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * 256.0f, (v1 - hf.y) * 256.0f), __NV_SATFINITE, __NV_E5M2);
}
__device__ __forceinline__ void store_mixed8(__nv_bfloat16* base, long long plane, size_t o, int c0, const float* f) {
    uint32_t h[4], x[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) mixed_cvt2(f[2 * i], f[2 * i + 1], h[i], x[i], l[i]);
    *reinterpret_cast<uint4*>(base + o) = make_uint4(h[0], h[1], h[2], h[3]);
    uint8_t* b1 = reinterpret_cast<uint8_t*>(base + plane) + 2 * o - (size_t)(c0 & 63);
    *reinterpret_cast<uint2*>(b1) = make_uint2(x[0] | (x[1] << 16), x[2] | (x[3] << 16));
    *reinterpret_cast<uint2*>(b1 + 64) = make_uint2(l[0] | (l[1] << 16), l[2] | (l[3] << 16));
}
// 4 consecutive channels (c0 % 4 == 0)
__device__ __forceinline__ void store_mixed4(__nv_bfloat16* base, long long plane, size_t o, int c0, const float* f) {
    uint32_t h[2], x[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) mixed_cvt2(f[2 * i], f[2 * i + 1], h[i], x[i], l[i]);
    *reinterpret_cast<uint2*>(base + o) = make_uint2(h[0], h[1]);
    uint8_t* b1 = reinterpret_cast<uint8_t*>(base + plane) + 2 * o - (size_t)(c0 & 63);
    *reinterpret_cast<uint32_t*>(b1) = x[0] | (x[1] << 16);
    *reinterpret_cast<uint32_t*>(b1 + 64) = l[0] | (l[1] << 16);
}
// inverse of store_mixed8 for readers outside the tensor-core kernels: v ~ x16 + xl8 (the e5m2(v) copy is not needed)
__device__ __forceinline__ void load_mixed8(const __nv_bfloat16* base, long long plane, size_t o, int c0, float* f) {
    const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(base + o));
    const uint2 l2 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint8_t*>(base + plane) + 2 * o - (size_t)(c0 & 63) + 64));
    const __half2* hp = reinterpret_cast<const __half2*>(&h4);
    const uint8_t* lb = reinterpret_cast<const uint8_t*>(&l2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 hf = __half22float2(hp[i]);
        const __half_raw r0 = {(unsigned short)(lb[2 * i] << 8)}, r1 = {(unsigned short)(lb[2 * i + 1] << 8)};      // e5m2 = the high byte of an fp16
        f[2 * i] = hf.x + __half2float(__half(r0)) * (1.0f / 256.0f);
        f[2 * i + 1] = hf.y + __half2float(__half(r1)) * (1.0f / 256.0f);
    }
}
}  // namespace evk
