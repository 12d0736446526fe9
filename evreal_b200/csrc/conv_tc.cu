// Tensor-core implicit-GEMM convolution for sm_100a: TMA-staged activation / weight tiles, tcgen05.mma with the
// fp32 accumulator in tensor memory, and the ConvLayer / ConvLSTM epilogues fused behind tcgen05.ld.
//
// Reference semantics: model/submodules.py:8-35 (ConvLayer), :152-184 (ResidualBlock), :187-245 (ConvLSTM),
// :69-97 (the 5x5 convolution of UpsampleConvLayer).  fp32 parity (1e-4, north_star) rules out single-pass
// bf16/tf32 operands (2e-3 / 3e-4), so every operand travels as TWO bf16 planes (hi = bf16(v),
// lo = bf16(v - hi)) and each K step issues three MMAs into the same TMEM accumulator:
//     D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo          (error ~ 2^-16 relative; measured 4-6e-6 end to end)
//
// GEMM view: M = 128 output pixels (a TH x TW spatial patch of one image), N = BN output channels,
// K = kh*kw*(c1+c2) walked as (tap, channel chunk of BK).  For one K block the A tile is the input patch
// shifted by the tap offset: ONE tiled TMA load per plane from the 5-D tensor [plane, n, y, x, c] with
// out-of-bounds zero fill doing the convolution's zero padding (and element strides doing stride 2), landing in
// shared memory in the 128B-swizzled K-major layout tcgen05.mma consumes.  cat(x, h) is two tensor maps.
//
// Warp roles (256 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (thread r owns accumulator row r = one output pixel).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "conv.cuh"
#include "tc.cuh"

namespace evk {

struct TcArgs {
    int N, Hout, Wout, stride, pad, kh, kw;
    int th, tw, tiles_x, tiles_y;
    int chunks1, chunks2;
    int bn, n_tiles, m_tiles, cout;
    int epi, act;
    int stages;
    int acc_stride;         // TMEM columns per accumulator stage
    uint32_t tmem_cols;
    const float* bias;
    const float* res;
    float* y;
    __nv_bfloat16* ys; long long ys_plane;
    const float* c_prev; float* c_new; float* h_new;
    __nv_bfloat16* hs_new; long long hs_plane;
};

struct TcPlan {
    CUtensorMap tm_x1, tm_x2, tm_w;
    TcArgs a;
    int bk;
    dim3 grid;
    size_t smem;
};

constexpr int kTcThreads = 384;      // 12 warps: TMA, MMA, TMEM alloc, (idle), 8 epilogue
constexpr int kEpiWarps = 8;

// sigmoid / tanh on the SFU (ex2.approx, rcp.approx): absolute error ~2e-7, far inside the 1e-4 parity budget
__device__ __forceinline__ float fast_sigmoid(float x) { return __frcp_rn(1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ float fast_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_SIGMOID: return fast_sigmoid(v);
        case ACT_TANH: return fast_tanh(v);
        default: return v;
    }
}

// Persistent, warp-specialised: every CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...  (N tile fastest
// so CTAs that run concurrently share the activation patch in L2).  The smem ring (TMA -> MMA) runs continuously
// across tiles; the accumulator is double-buffered in tensor memory so the epilogue of tile i overlaps the MMAs of
// tile i+1.
template <int BK>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
               const __grid_constant__ CUtensorMap tm_w, const TcArgs a) {
    constexpr uint32_t ROW_BYTES = BK * 2;
    constexpr uint32_t A_BYTES = 128 * ROW_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)a.bn * ROW_BYTES;
    const uint32_t stage_bytes = 2 * A_BYTES + 2 * b_bytes;
    const uint32_t bar_full = base + (uint32_t)a.stages * stage_bytes;
    const uint32_t bar_empty = bar_full + 8u * a.stages;
    const uint32_t bar_tfull = bar_empty + 8u * a.stages;     // [2] accumulator ready
    const uint32_t bar_tempty = bar_tfull + 16u;              // [2] accumulator drained by the epilogue
    const uint32_t slot = bar_tempty + 16u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int chunks = a.chunks1 + a.chunks2;
    const int KB = a.kh * a.kw * chunks;
    const int total_tiles = a.m_tiles * a.n_tiles;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x1);
        if (a.chunks2) tma_prefetch_desc(&tm_x2);
        tma_prefetch_desc(&tm_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < a.stages; ++s) {
            mbar_init(bar_full + 8u * s, 1);
            mbar_init(bar_empty + 8u * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8u * i, 1);
            mbar_init(bar_tempty + 8u * i, kEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 2) tc_alloc(slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            uint32_t cnt = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int nt = t % a.n_tiles, mt = t / a.n_tiles;
                const int img = mt / tiles_per_img, rem = mt - img * tiles_per_img;
                const int oy0 = (rem / a.tiles_x) * a.th, ox0 = (rem % a.tiles_x) * a.tw;
                const int n0 = nt * a.bn;
                for (int kb = 0; kb < KB; ++kb, ++cnt) {
                    const uint32_t s = cnt % (uint32_t)a.stages;
                    const uint32_t ph = (cnt / (uint32_t)a.stages) & 1u;
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    mbar_expect_tx(bar_full + 8u * s, stage_bytes);
                    const int tap = kb / chunks, ch = kb - tap * chunks;
                    const int r = tap / a.kw, q = tap - r * a.kw;
                    const int ix0 = ox0 * a.stride - a.pad + q, iy0 = oy0 * a.stride - a.pad + r;
                    const bool first = ch < a.chunks1;
                    const CUtensorMap* m = first ? &tm_x1 : &tm_x2;
                    const int c0 = (first ? ch : ch - a.chunks1) * BK;
                    const uint32_t sa = base + s * stage_bytes;
                    tma_load_5d(sa, m, bar_full + 8u * s, c0, ix0, iy0, img, 0);
                    tma_load_5d(sa + A_BYTES, m, bar_full + 8u * s, c0, ix0, iy0, img, 1);
                    tma_load_3d(sa + 2 * A_BYTES, &tm_w, bar_full + 8u * s, kb * BK, n0, 0);
                    tma_load_3d(sa + 2 * A_BYTES + b_bytes, &tm_w, bar_full + 8u * s, kb * BK, n0, 1);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer.  Two MMAs per 16-deep K step: B_hi and B_lo are adjacent in the stage, so
            //   D[:, 0:2bn]  (+)= A_hi * [B_hi; B_lo]^T      (N = 2*bn: hi*hi | hi*lo)
            //   D[:, 0:bn]    += A_lo *  B_hi^T
            // (A is read from shared memory twice instead of three times; the epilogue adds the two halves.)
            const uint32_t idesc2 = umma_idesc_bf16(128, (uint32_t)(2 * a.bn));
            const uint32_t idesc1 = umma_idesc_bf16(128, (uint32_t)a.bn);
            uint32_t cnt = 0, it = 0;
            bool ready = false;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
                mbar_wait(bar_tempty + 8u * as, aph ^ 1u);          // epilogue drained this accumulator stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * (uint32_t)a.acc_stride;
                for (int kb = 0; kb < KB; ++kb, ++cnt) {
                    const uint32_t s = cnt % (uint32_t)a.stages;
                    const uint32_t ph = (cnt / (uint32_t)a.stages) & 1u;
                    if (!ready) mbar_wait(bar_full + 8u * s, ph);
                    tc_fence_after();
                    // poll the NEXT stage now: the round trip of the barrier read overlaps the MMA issue below
                    const uint32_t s1 = (cnt + 1) % (uint32_t)a.stages;
                    const uint32_t ph1 = ((cnt + 1) / (uint32_t)a.stages) & 1u;
                    ready = mbar_try_wait(bar_full + 8u * s1, ph1);
                    const uint32_t sa = base + s * stage_bytes;
                    const uint64_t ah = umma_desc_kmajor(sa, ROW_BYTES);
                    const uint64_t al = umma_desc_kmajor(sa + A_BYTES, ROW_BYTES);
                    const uint64_t bh = umma_desc_kmajor(sa + 2 * A_BYTES, ROW_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // advancing 16 elements (32 B) along K inside the swizzle atom = +2 in the 16-byte address field
                        tc_mma_bf16(d_tmem, ah + 2 * k, bh + 2 * k, idesc2, (kb | k) != 0 ? 1u : 0u);
                        tc_mma_bf16(d_tmem, al + 2 * k, bh + 2 * k, idesc1, 1u);
                    }
                    tc_commit(bar_empty + 8u * s);      // frees the smem slot when these MMAs retire
                }
                tc_commit(bar_tfull + 8u * as);         // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warp (4 + e) owns TMEM lane quadrant e % 4 (its hardware-accessible lanes) and
        // the 32-column chunks with parity e / 4.  Thread = accumulator row = output pixel.
        const int e = warp - 4;
        const int wq = e & 3, half = e >> 2;
        const int row = wq * 32 + lane;
        const int ly = row / a.tw, lx = row - ly * a.tw;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
            const int nt = t % a.n_tiles, mt = t / a.n_tiles;
            const int img = mt / tiles_per_img, rem = mt - img * tiles_per_img;
            const int oy = (rem / a.tiles_x) * a.th + ly, ox = (rem % a.tiles_x) * a.tw + lx;
            const int n0 = nt * a.bn;
            const bool valid = oy < a.Hout && ox < a.Wout;
            const size_t pix = ((size_t)img * a.Hout + oy) * a.Wout + ox;
            mbar_wait(bar_tfull + 8u * as, aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + as * (uint32_t)a.acc_stride + ((uint32_t)(wq * 32) << 16);
            const int nchunks = (a.bn + 31) / 32;
            for (int c = half; c < nchunks; c += 2) {
                const int j0 = c * 32;
                uint32_t v[32];
                __syncwarp();                     // tcgen05.ld is .sync.aligned: reconverge after the masked stores
                {
                    uint32_t u[32];
                    tc_ld_32x32(t_row + (uint32_t)j0, v);
                    tc_ld_32x32(t_row + (uint32_t)(a.bn + j0), u);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
                }
                if (c + 2 >= nchunks) {           // last TMEM read of this warp for this tile: release the stage
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * as);
                }
                const int nb = n0 + j0;
                if (!valid || nb >= a.cout) continue;
                if (a.epi == EPI_LINEAR) {
                    const size_t o = pix * a.cout + nb;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (nb + g * 4 >= a.cout || j0 + g * 4 >= a.bn) break;
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                        float f[4] = {__uint_as_float(v[g * 4 + 0]) + b4.x, __uint_as_float(v[g * 4 + 1]) + b4.y,
                                      __uint_as_float(v[g * 4 + 2]) + b4.z, __uint_as_float(v[g * 4 + 3]) + b4.w};
                        if (a.res != nullptr) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.res + o + g * 4));
                            f[0] += r4.x; f[1] += r4.y; f[2] += r4.z; f[3] += r4.w;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) f[i] = fast_act(f[i], a.act);
                        if (a.y != nullptr) *reinterpret_cast<float4*>(a.y + o + g * 4) = make_float4(f[0], f[1], f[2], f[3]);
                        if (a.ys != nullptr) {
                            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_bf16(f[i], hi[i], lo[i]);
                            *reinterpret_cast<uint2*>(a.ys + o + g * 4) = *reinterpret_cast<uint2*>(hi);
                            *reinterpret_cast<uint2*>(a.ys + a.ys_plane + o + g * 4) = *reinterpret_cast<uint2*>(lo);
                        }
                    }
                } else {   // EPI_LSTM: packed column = channel*4 + {in, remember, out, cell}
                    const int C = a.cout >> 2;
                    const size_t o = pix * C + (nb >> 2);
                    const float4 cp0 = *reinterpret_cast<const float4*>(a.c_prev + o);
                    const float4 cp1 = *reinterpret_cast<const float4*>(a.c_prev + o + 4);
                    const float cprev[8] = {cp0.x, cp0.y, cp0.z, cp0.w, cp1.x, cp1.y, cp1.z, cp1.w};
                    float cn[8], hn[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                        const float ig = fast_sigmoid(__uint_as_float(v[g * 4 + 0]) + b4.x);
                        const float fg = fast_sigmoid(__uint_as_float(v[g * 4 + 1]) + b4.y);
                        const float og = fast_sigmoid(__uint_as_float(v[g * 4 + 2]) + b4.z);
                        const float cg = fast_tanh(__uint_as_float(v[g * 4 + 3]) + b4.w);
                        const float cell = __fadd_rn(__fmul_rn(fg, cprev[g]), __fmul_rn(ig, cg));
                        cn[g] = cell;
                        hn[g] = og * fast_tanh(cell);
                    }
                    *reinterpret_cast<float4*>(a.c_new + o) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(a.c_new + o + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                    *reinterpret_cast<float4*>(a.h_new + o) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(a.h_new + o + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                    if (a.hs_new != nullptr) {
                        __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) split_bf16(hn[i], hi[i], lo[i]);
                        *reinterpret_cast<uint4*>(a.hs_new + o) = *reinterpret_cast<uint4*>(hi);
                        *reinterpret_cast<uint4*>(a.hs_new + a.hs_plane + o) = *reinterpret_cast<uint4*>(lo);
                    }
                }
            }
            // a warp with no chunk of its parity (bn <= 32 and half == 1) still has to release the stage
            if (half >= nchunks) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8u * as);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------ host side
static void* g_encode = nullptr;

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
    if (!g_encode) {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        EVK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        EVK_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, EVK_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        g_encode = fn;
    }
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides[i]; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    const CUresult r = ((Fn)g_encode)(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs,
                                      bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EVK_REQUIRE(r == CUDA_SUCCESS, EVK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u %u %u)", (int)r,
                rank, box[0], box[1], rank > 2 ? box[2] : 0);
    return EVK_OK;
}

static int pick_bk(const ConvParams& p) {
    if (p.c1 % 64 == 0 && p.c2 % 64 == 0) return 64;
    if (p.c1 % 32 == 0 && p.c2 % 32 == 0) return 32;
    return 0;
}

static bool g_tc_stride2 = true;   // EVK_TC_STRIDE2=0 routes stride-2 convolutions to the fp32 SIMT kernel

bool tc_eligible(const ConvParams& p) {
    static bool env_read = false;
    if (!env_read) {
        const char* e = getenv("EVK_TC_STRIDE2");
        if (e && e[0] == '0') g_tc_stride2 = false;
        env_read = true;
    }
    if (p.epi != EPI_LINEAR && p.epi != EPI_LSTM) return false;
    if (pick_bk(p) == 0 || p.c1 == 0) return false;
    if (p.stride != 1 && !(p.stride == 2 && g_tc_stride2)) return false;
    if (p.cout % 4 != 0) return false;
    if (p.epi == EPI_LSTM && p.cout % 32 != 0) return false;
    return true;
}

// N tile: the divisor of cout_pad (multiple of 16, <= 128 so that [B_hi; B_lo] is one N <= 256 operand) that
// minimises  waves(CTAs / 148) * (fixed per-CTA cost + K steps * cycles per step), cycles per 16-deep K step from
// the tcgen05 floor (128*N/256) and the 128 B/clk shared-memory operand read.
static int pick_bn(int cout_pad, long m_tiles, long k16_steps, int granule) {
    int best = 0;
    double best_cost = 0.0;
    for (int bn = 128; bn >= 16; bn -= 16) {
        if (cout_pad % bn != 0 || bn % granule != 0) continue;
        const double mma1 = std::max((double)bn, 32.0 + bn / 2.0);          // N = 2*bn
        const double mma2 = std::max(bn / 2.0, 32.0 + bn / 4.0);            // N = bn
        const double per_cta = 8000.0 + (double)k16_steps * (mma1 + mma2) + 40.0 * bn;
        const long ctas = m_tiles * (cout_pad / bn);
        const double cost = (double)((ctas + kNumSMs - 1) / kNumSMs) * per_cta;
        if (best == 0 || cost < best_cost) { best = bn; best_cost = cost; }
    }
    return best;
}

void pack_weights_tc(const float* w_kc, int K, int cout, int cout_pad, std::vector<__nv_bfloat16>& out) {
    out.assign((size_t)2 * cout_pad * K, __float2bfloat16(0.0f));
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < cout; ++n) {
            const float v = w_kc[(size_t)k * cout + n];
            const __nv_bfloat16 hi = __float2bfloat16(v);
            const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
            out[(size_t)n * K + k] = hi;
            out[(size_t)cout_pad * K + (size_t)n * K + k] = lo;
        }
}

int tc_plan_create(ConvParams& p) {
    EVK_REQUIRE(tc_eligible(p), EVK_ERR_ARG, "conv_tc: shape not eligible for the tensor-core path");
    EVK_REQUIRE(p.x1s && p.w_tc && (p.c2 == 0 || p.x2s), EVK_ERR_ARG, "conv_tc: split operands missing");
    const int bk = pick_bk(p);
    const int cout_pad = p.cout_pad;
    EVK_REQUIRE(cout_pad >= p.cout && cout_pad % 16 == 0, EVK_ERR_ARG, "conv_tc: cout_pad=%d must be a multiple of 16 >= cout", cout_pad);
    // spatial M tile: 128 pixels, least padded area
    const int cand[4][2] = {{8, 16}, {4, 32}, {16, 8}, {2, 64}};
    int th = 8, tw = 16;
    long best = -1;
    for (auto& c : cand) {
        if ((c[1] - 1) * p.stride + 1 > 256) continue;
        const long area = (long)ceil_div(p.Hout, c[0]) * c[0] * ceil_div(p.Wout, c[1]) * c[1];
        if (best < 0 || area < best) { best = area; th = c[0]; tw = c[1]; }
    }
    const long m_tiles = (long)ceil_div(p.Hout, th) * ceil_div(p.Wout, tw) * p.N;
    const int bn = pick_bn(cout_pad, m_tiles, (long)p.kh * p.kw * (p.c1 + p.c2) / 16, p.epi == EPI_LSTM ? 32 : 16);
    EVK_REQUIRE(bn >= 16, EVK_ERR_ARG, "conv_tc: no N tile for cout_pad=%d", cout_pad);
    TcPlan* pl = new TcPlan();
    pl->bk = bk;
    TcArgs& a = pl->a;
    a.N = p.N; a.Hout = p.Hout; a.Wout = p.Wout; a.stride = p.stride; a.pad = p.pad; a.kh = p.kh; a.kw = p.kw;
    a.th = th; a.tw = tw; a.tiles_x = ceil_div(p.Wout, tw); a.tiles_y = ceil_div(p.Hout, th);
    a.chunks1 = p.c1 / bk; a.chunks2 = p.c2 / bk;
    a.bn = bn; a.n_tiles = cout_pad / bn; a.m_tiles = (int)m_tiles; a.cout = p.cout; a.epi = p.epi; a.act = p.act;
    a.bias = p.bias; a.res = p.res; a.y = p.y; a.ys = p.ys;
    a.ys_plane = (long long)p.N * p.Hout * p.Wout * p.cout;
    a.c_prev = p.c_prev; a.c_new = p.c_new; a.h_new = p.h_new; a.hs_new = p.hs_new;
    a.hs_plane = (long long)p.N * p.Hout * p.Wout * (p.cout / 4);
    const uint32_t row_bytes = bk * 2;
    const size_t stage_bytes = 2 * (size_t)128 * row_bytes + 2 * (size_t)bn * row_bytes;
    int stages = (int)((227 * 1024 - 2048) / stage_bytes);
    stages = std::max(2, std::min(stages, 6));
    a.stages = stages;
    // per accumulator stage: [hi*hi + lo*hi | hi*lo] = 2*bn columns, read in 32-column windows (bn%32 tail -> pad)
    a.acc_stride = (bn + (bn + 31) / 32 * 32 + 31) / 32 * 32;
    uint32_t cols = 32;
    while ((int)cols < 2 * a.acc_stride) cols <<= 1;
    EVK_REQUIRE(cols <= 512, EVK_ERR_ARG, "conv_tc: accumulator does not fit tensor memory (bn=%d)", bn);
    a.tmem_cols = cols;
    pl->smem = stages * stage_bytes + 1024 + 16 * stages + 64;
    const long total_tiles = m_tiles * (cout_pad / bn);
    pl->grid = dim3((unsigned)std::min<long>(total_tiles, kNumSMs));
    // activations: [plane, n, y, x, c]
    auto act_map = [&](CUtensorMap* m, const __nv_bfloat16* base, int C) -> int {
        const uint64_t dims[5] = {(uint64_t)C, (uint64_t)p.Win, (uint64_t)p.Hin, (uint64_t)p.N, 2};
        const uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)p.Win * C * 2, (uint64_t)p.Hin * p.Win * C * 2,
                                 (uint64_t)p.N * p.Hin * p.Win * C * 2};
        const uint32_t box[5] = {(uint32_t)bk, (uint32_t)((tw - 1) * p.stride + 1), (uint32_t)((th - 1) * p.stride + 1), 1, 1};
        const uint32_t es[5] = {1, (uint32_t)p.stride, (uint32_t)p.stride, 1, 1};
        return encode_tmap_bf16(m, base, 5, dims, str, box, es, (int)row_bytes);
    };
    int r = act_map(&pl->tm_x1, p.x1s, p.c1);
    if (r == EVK_OK) r = p.c2 ? act_map(&pl->tm_x2, p.x2s, p.c2) : act_map(&pl->tm_x2, p.x1s, p.c1);
    if (r == EVK_OK) {
        const uint64_t K = (uint64_t)p.kh * p.kw * (p.c1 + p.c2);
        const uint64_t dims[3] = {K, (uint64_t)cout_pad, 2};
        const uint64_t str[2] = {K * 2, (uint64_t)cout_pad * K * 2};
        const uint32_t box[3] = {(uint32_t)bk, (uint32_t)bn, 1};
        const uint32_t es[3] = {1, 1, 1};
        r = encode_tmap_bf16(&pl->tm_w, p.w_tc, 3, dims, str, box, es, (int)row_bytes);
    }
    if (r != EVK_OK) { delete pl; return r; }
    static bool attr_set = false;
    if (!attr_set) {
        EVK_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        EVK_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    p.tc = pl;
    return EVK_OK;
}

void tc_plan_destroy(TcPlan* plan) { delete plan; }

int launch_conv_tc(const ConvParams& p, cudaStream_t st) {
    EVK_REQUIRE(p.tc != nullptr, EVK_ERR_STATE, "conv_tc: no plan");
    const TcPlan& pl = *p.tc;
    if (pl.bk == 64)
        conv_tc_kernel<64><<<pl.grid, kTcThreads, pl.smem, st>>>(pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a);
    else
        conv_tc_kernel<32><<<pl.grid, kTcThreads, pl.smem, st>>>(pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ------------------------------------------------------------------ fp32 -> split planes
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 hi, lo;
        split_bf16(src[i], hi, lo);
        dst[i] = hi;
        dst[n + i] = lo;
    }
}

int launch_split(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t st) {
    EVK_REQUIRE(src && dst && n > 0, EVK_ERR_ARG, "split: bad argument");
    split_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 2368), 256, 0, st>>>(src, dst, n);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
