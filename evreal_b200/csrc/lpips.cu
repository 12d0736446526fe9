// Stage 3, LPIPS: learned perceptual distance between a reconstruction and its reference frame.
//
// Reference semantics: utils/eval_metrics.py:100-156 (PyIqaMetricFactory: grey frame repeated to 3 channels by
// cv2torch(num_ch=3), utils/eval_utils.py:46-54; queue of 4 frames; iqa_metric(img, ref)) around pyiqa's LPIPS, a
// third-party dependency that is NOT in the reference tree (requirements.txt:7, unpinned).  Its published algorithm
// (richzhang/PerceptualSimilarity v0.1, restated in SURVEY A.3): x -> 2x-1 -> (x - shift)/scale; backbone features
// at five ReLU taps (AlexNet relu1..5: 64/192/384/256/256 channels; VGG16 relu1_2, 2_2, 3_3, 4_3, 5_3:
// 64/128/256/512/512); per tap unit-normalise over channels (x / (||x||_2 + 1e-10)), squared difference,
// non-negative 1x1 "lin" weights, spatial mean; sum over taps.  The weights do not exist offline: parity is pinned
// only against oracle/lpips.py (a torch-CPU restatement) with seeded weights -- "parity unpinned" against pyiqa.
//
// The backbone runs on the same convolution kernels as the reconstruction networks (north_star: "SSIM/LPIPS reusing
// the same conv kernel"): tcgen05 split-bf16 implicit GEMM wherever the shape qualifies (all AlexNet layers but the
// 11x11 stride-4 stem, all VGG16 layers but the Cin=3 stem), fp32 CUDA-core kernel otherwise.
#include <map>
#include <string>
#include <vector>
#include <cmath>

#include "conv.cuh"
#include "tc.cuh"

namespace evk {

struct LpLayer {
    int kind = 0;                 // 0 conv (+ReLU), 1 max pool
    ConvParams cp;
    const float* in = nullptr; float* out = nullptr; __nv_bfloat16* out_s = nullptr;
    int mixed = 0;              // max pooling: out_s is written in the mixed-operand format (conv.cuh, ConvParams::mixed)
    int C = 0, H = 0, W = 0, k = 0, stride = 0, Ho = 0, Wo = 0;
    double flops = 0.0;
};

struct LpTap { const float* feat = nullptr; int C = 0, H = 0, W = 0; const float* lin = nullptr; };

// grey [n,H,W] in [0,1] (img rows 0..batch-1, ref rows batch..2*batch-1) -> scaled 3-channel NHWC4 (4th channel 0)
__global__ void __launch_bounds__(256) lpips_prep_kernel(const float* __restrict__ img, const float* __restrict__ ref, int n, int batch,
                                                         int64_t pixels, float* __restrict__ out) {
    const int64_t total = (int64_t)2 * batch * pixels;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t im = i / pixels, pix = i - im * pixels;
        const int b = (int)(im % batch);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < n) {
            // (clamped to [0,1]: the clip of EvalMetricsTracker.update, utils/eval_metrics.py:253-255 -- a no-op for in-range frames)
            const float v = fminf(fmaxf((im < batch ? img : ref)[(int64_t)b * pixels + pix], 0.0f), 1.0f);
            const float s = 2.0f * v - 1.0f;
            o.x = (s - (-0.030f)) / 0.458f;
            o.y = (s - (-0.088f)) / 0.448f;
            o.z = (s - (-0.188f)) / 0.450f;
        }
        reinterpret_cast<float4*>(out)[i] = o;
    }
}

__device__ __forceinline__ float3 lpips_scale(float v) {
    const float s = 2.0f * fminf(fmaxf(v, 0.0f), 1.0f) - 1.0f;
    return make_float3((s - (-0.030f)) / 0.458f, (s - (-0.088f)) / 0.448f, (s - (-0.188f)) / 0.450f);
}

// Tensor-core form of the AlexNet stem (11x11, stride 4, padding 2, 3 -> 64): space-to-depth.  Block (by, bx) of the
// output holds input rows 4*by - 2 .. 4*by + 1 and columns 4*bx - 2 .. 4*bx + 1 (zero outside the image) as 4*4*3 = 48
// "channels" (padded to 64: one 128-byte K row); output pixel (y, x) of the stem then reads exactly blocks y .. y+2 by
// x .. x+2 -- an ordinary unpadded stride-1 3x3 convolution with 64 input channels (taps beyond row / column 10 get zero
// weights).  Written directly as split-bf16 planes [2][N][Hb][Wb][64].
__global__ void __launch_bounds__(256) lpips_prep_s2d_kernel(const float* __restrict__ img, const float* __restrict__ ref, int n, int batch,
                                                             int H, int W, int Hb, int Wb, __nv_bfloat16* __restrict__ out) {
    const int64_t total = (int64_t)2 * batch * Hb * Wb * 16;           // one thread per (block, sub-pixel)
    const int64_t plane = (int64_t)2 * batch * Hb * Wb * 64;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int sp = (int)(i % 16);
        const int bx = (int)((i / 16) % Wb);
        const int by = (int)((i / ((int64_t)16 * Wb)) % Hb);
        const int im = (int)(i / ((int64_t)16 * Wb * Hb));
        const int b = im % batch;
        const int y = 4 * by - 2 + sp / 4, x = 4 * bx - 2 + sp % 4;
        float3 v = make_float3(0.f, 0.f, 0.f);
        if (b < n && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W)
            v = lpips_scale((im < batch ? img : ref)[((int64_t)b * H + y) * W + x]);
        const int64_t o = (((int64_t)im * Hb + by) * Wb + bx) * 64 + sp * 3;
        const float f[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            __nv_bfloat16 hi, lo;
            split_bf16(f[c], hi, lo);
            out[o + c] = hi;
            out[plane + o + c] = lo;
        }
        if (sp == 15) {                       // channels 48..63 stay zero
#pragma unroll
            for (int c = 48; c < 64; ++c) { out[o - 45 + c] = __float2bfloat16(0.f); out[plane + o - 45 + c] = __float2bfloat16(0.f); }
        }
    }
}

// Tensor-core form of the VGG16 stem (3x3, 3 -> 64): the row-window layout of the reconstruction networks' head
// convolution (conv.cuh, ConvParams::kw_packed): [2][N][H][W + 8][8], pixel x at column x + 1, channels 3..7 and the pad
// columns zero.
__global__ void __launch_bounds__(256) lpips_prep_rowwin_kernel(const float* __restrict__ img, const float* __restrict__ ref, int n, int batch,
                                                                int H, int W, __nv_bfloat16* __restrict__ out) {
    const int Wp = W + 8;
    const int64_t total = (int64_t)2 * batch * H * Wp;
    const int64_t plane = total * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xp = (int)(i % Wp);
        const int y = (int)((i / Wp) % H);
        const int im = (int)(i / ((int64_t)Wp * H));
        const int b = im % batch, x = xp - 1;
        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { hi[c] = __float2bfloat16(0.f); lo[c] = __float2bfloat16(0.f); }
        if (b < n && (unsigned)x < (unsigned)W) {
            const float3 v = lpips_scale((im < batch ? img : ref)[((int64_t)b * H + y) * W + x]);
            split_bf16(v.x, hi[0], lo[0]); split_bf16(v.y, hi[1], lo[1]); split_bf16(v.z, hi[2], lo[2]);
        }
        *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out + plane + i * 8) = *reinterpret_cast<const uint4*>(lo);
    }
}

// NHWC max pooling (no padding, floor mode): torch.nn.MaxPool2d(k, stride); writes fp32 and the split-bf16 copy
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, __nv_bfloat16* __restrict__ ys,
                                                      int N, int H, int W, int C4, int k, int s, int Ho, int Wo, int mixed) {
    const int64_t total = (int64_t)N * Ho * Wo * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int ox = (int)((i / C4) % Wo);
        const int oy = (int)((i / ((int64_t)C4 * Wo)) % Ho);
        const int n = (int)(i / ((int64_t)C4 * Wo * Ho));
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int r = 0; r < k; ++r)
            for (int q = 0; q < k; ++q) {
                const int iy = oy * s + r, ix = ox * s + q;
                if (iy >= H || ix >= W) continue;
                const float4 v = __ldg(reinterpret_cast<const float4*>(x) + (((int64_t)n * H + iy) * W + ix) * C4 + c);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        if (y != nullptr) reinterpret_cast<float4*>(y)[i] = m;
        if (ys != nullptr) {
            const float f[4] = {m.x, m.y, m.z, m.w};
            if (mixed) { store_mixed4(ys, (long long)total * 4, (size_t)i * 4, c * 4, f); continue; }      // mixed-operand consumer (tc.cuh)
            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16(f[e], hi[e], lo[e]);
            *reinterpret_cast<uint2*>(ys + i * 4) = *reinterpret_cast<uint2*>(hi);
            *reinterpret_cast<uint2*>(ys + total * 4 + i * 4) = *reinterpret_cast<uint2*>(lo);
        }
    }
}

// kTapChunks CTAs per (pair, tap), each over a contiguous range of pixels: sum over pixels of
// sum_c lin[c] * (f0/(|f0|+eps) - f1/(|f1|+eps))^2; fixed summation tree (partials are added in index order by lpips_sum_kernel).
// A pixel is handled by LPP = 16 or 32 lanes with 16-byte loads; its channels stay in registers between the norm pass and the
// distance pass (the features are read from memory exactly once: this reduction is HBM bound, 2 * C * 4 bytes per pixel pair).
// Specialised per channel count (a generic version that unrolled to the largest count was 2.8x slower at 64 channels).
constexpr int kTapChunks = 16;
template <int NVEC, int LPP>                        // float4 per lane and image; lanes per pixel (C = 4 * NVEC * LPP)
__global__ void __launch_bounds__(256) lpips_tap_kernel(const float* __restrict__ feat, const float* __restrict__ lin, int batch, int pixels,
                                                        double* __restrict__ part /*[batch][kTapChunks]*/) {
    constexpr int C4 = NVEC * LPP, GPW = 32 / LPP;      // float4 per pixel; pixels per warp and iteration
    const int p = blockIdx.x;
    const float4* f0 = reinterpret_cast<const float4*>(feat) + (size_t)p * pixels * C4;
    const float4* f1 = reinterpret_cast<const float4*>(feat) + (size_t)(batch + p) * pixels * C4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gl = lane % LPP, grp = lane / LPP;
    float4 w4[NVEC];
#pragma unroll
    for (int k = 0; k < NVEC; ++k) w4[k] = __ldg(reinterpret_cast<const float4*>(lin) + gl + k * LPP);
    const int per = (pixels + kTapChunks - 1) / kTapChunks;
    const int px0 = blockIdx.y * per, px1 = min(pixels, px0 + per);
    double acc = 0.0;
    for (int pxb = px0 + warp * GPW; pxb < px1; pxb += 8 * GPW) {
        const int px = pxb + grp;
        const bool ok = px < px1;
        float4 a4[NVEC], b4[NVEC];
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            a4[k] = ok ? __ldg(f0 + (size_t)px * C4 + gl + k * LPP) : make_float4(0.f, 0.f, 0.f, 0.f);
            b4[k] = ok ? __ldg(f1 + (size_t)px * C4 + gl + k * LPP) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            sa = fmaf(a4[k].x, a4[k].x, sa); sa = fmaf(a4[k].y, a4[k].y, sa); sa = fmaf(a4[k].z, a4[k].z, sa); sa = fmaf(a4[k].w, a4[k].w, sa);
            sb = fmaf(b4[k].x, b4[k].x, sb); sb = fmaf(b4[k].y, b4[k].y, sb); sb = fmaf(b4[k].z, b4[k].z, sb); sb = fmaf(b4[k].w, b4[k].w, sb);
        }
#pragma unroll
        for (int o = LPP >> 1; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
        const float na = sqrtf(sa) + 1e-10f, nb = sqrtf(sb) + 1e-10f;
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < NVEC; ++k) {
            float t;
            t = __fdividef(a4[k].x, na) - __fdividef(b4[k].x, nb); d = fmaf(w4[k].x, t * t, d);
            t = __fdividef(a4[k].y, na) - __fdividef(b4[k].y, nb); d = fmaf(w4[k].y, t * t, d);
            t = __fdividef(a4[k].z, na) - __fdividef(b4[k].z, nb); d = fmaf(w4[k].z, t * t, d);
            t = __fdividef(a4[k].w, na) - __fdividef(b4[k].w, nb); d = fmaf(w4[k].w, t * t, d);
        }
#pragma unroll
        for (int o = LPP >> 1; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (ok && gl == 0) acc += (double)d;
    }
    if (LPP == 16) acc += __shfl_xor_sync(0xffffffffu, acc, 16);     // the leader of the warp's second pixel hands its sum to lane 0
    __shared__ double sm[8];
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sm[w];
        part[(size_t)p * kTapChunks + blockIdx.y] = s;
    }
}

// (first version, kept for A/B: a warp per pixel, lanes over channels, scalar loads, two passes over memory)
__global__ void __launch_bounds__(256) lpips_tap_kernel_v1(const float* __restrict__ feat, const float* __restrict__ lin, int batch, int pixels,
                                                        int C, double* __restrict__ part /*[batch][kTapChunks]*/) {
    const int p = blockIdx.x;
    const float* f0 = feat + (size_t)p * pixels * C;
    const float* f1 = feat + (size_t)(batch + p) * pixels * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (pixels + kTapChunks - 1) / kTapChunks;
    const int px0 = blockIdx.y * per, px1 = min(pixels, px0 + per);
    double acc = 0.0;
    for (int px = px0 + warp; px < px1; px += 8) {       // a warp per pixel, lanes over channels
        const float* a = f0 + (size_t)px * C;
        const float* b = f1 + (size_t)px * C;
        float sa = 0.f, sb = 0.f;
        for (int c = lane; c < C; c += 32) { sa = fmaf(a[c], a[c], sa); sb = fmaf(b[c], b[c], sb); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sb += __shfl_xor_sync(0xffffffffu, sb, o); }
        const float na = sqrtf(sa) + 1e-10f, nb = sqrtf(sb) + 1e-10f;
        float d = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float t = a[c] / na - b[c] / nb;
            d = fmaf(lin[c], t * t, d);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        acc += (double)d;
    }
    __shared__ double sm[8];
    if (lane == 0) sm[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sm[w];
        part[(size_t)p * kTapChunks + blockIdx.y] = s;
    }
}

struct LpPixels { int n[5]; };
__global__ void lpips_sum_kernel(const double* __restrict__ part, int taps, int batch, int n, LpPixels pixels, double* __restrict__ scores) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    double s = 0.0;
    for (int t = 0; t < taps; ++t) {
        double a = 0.0;
        for (int c = 0; c < kTapChunks; ++c) a += part[((size_t)t * batch + p) * kTapChunks + c];
        s += a / (double)pixels.n[t];                     // spatial average of the tap
    }
    scores[p] = s;
}

}  // namespace evk

using namespace evk;

struct evk_lpips {
    int backbone = 0, batch = 0, H = 0, W = 0, precision = 0;
    bool finalized = false;
    std::map<std::string, std::vector<float>> sd;
    std::map<std::string, std::vector<int64_t>> shapes;
    std::vector<void*> allocs;
    std::vector<LpLayer> layers;
    std::vector<LpTap> taps;
    std::vector<TcPlan*> plans;
    float* input = nullptr;
    __nv_bfloat16* input_s = nullptr;     // tensor-core stems: space-to-depth (AlexNet) / row-window (VGG16) split-bf16 input
    int stem_hb = 0, stem_wb = 0;         // AlexNet: blocks of the space-to-depth input
    double* part = nullptr;
    double flops = 0.0;

    DeviceArena arena;                    // (evk_common.cuh)
    void* dalloc(size_t bytes) {
        void* p = arena.alloc(bytes);
        allocs.push_back(p);
        return p;
    }
    const std::vector<float>* find(const std::vector<std::string>& names, std::string* found = nullptr) const {
        for (const std::string& n : names) {
            auto it = sd.find(n);
            if (it != sd.end()) { if (found) *found = n; return &it->second; }
        }
        return nullptr;
    }
};

namespace evk {

struct LpSpec { int idx; int cin, cout, k, stride, pad; int pool_k, pool_s; int tap; };   // pool (if any) runs BEFORE the conv

// torchvision `features` indices; lpips slice names "net.slice{s}.{idx}"
static const LpSpec kAlex[] = {{0, 3, 64, 11, 4, 2, 0, 0, 0},  {3, 64, 192, 5, 1, 2, 3, 2, 1}, {6, 192, 384, 3, 1, 1, 3, 2, 2},
                               {8, 384, 256, 3, 1, 1, 0, 0, 3}, {10, 256, 256, 3, 1, 1, 0, 0, 4}};
static const LpSpec kVgg[] = {{0, 3, 64, 3, 1, 1, 0, 0, -1},    {2, 64, 64, 3, 1, 1, 0, 0, 0},    {5, 64, 128, 3, 1, 1, 2, 2, -1},
                              {7, 128, 128, 3, 1, 1, 0, 0, 1},  {10, 128, 256, 3, 1, 1, 2, 2, -1}, {12, 256, 256, 3, 1, 1, 0, 0, -1},
                              {14, 256, 256, 3, 1, 1, 0, 0, 2}, {17, 256, 512, 3, 1, 1, 2, 2, -1}, {19, 512, 512, 3, 1, 1, 0, 0, -1},
                              {21, 512, 512, 3, 1, 1, 0, 0, 3}, {24, 512, 512, 3, 1, 1, 2, 2, -1}, {26, 512, 512, 3, 1, 1, 0, 0, -1},
                              {28, 512, 512, 3, 1, 1, 0, 0, 4}};

static int slice_of(int backbone, int idx) {
    if (backbone == 0) return idx < 2 ? 1 : idx < 5 ? 2 : idx < 8 ? 3 : idx < 10 ? 4 : 5;
    return idx < 4 ? 1 : idx < 9 ? 2 : idx < 16 ? 3 : idx < 23 ? 4 : 5;
}

static int lpips_build(evk_lpips* l) {
    const LpSpec* specs = l->backbone == 0 ? kAlex : kVgg;
    const int n_specs = l->backbone == 0 ? 5 : 13;
    const int N = 2 * l->batch;
    int H = l->H, W = l->W, C = 4;
    const bool tc = l->precision == 0;
    // tensor-core stems: AlexNet always (space-to-depth), VGG16 when two output pixels can share a GEMM row (even width)
    const bool stem_tc = tc && (l->backbone == 0 || W % 2 == 0) && getenv("EVK_LPIPS_SIMT_STEM") == nullptr;
    const float* x = nullptr;
    __nv_bfloat16* xs = nullptr;
    if (!stem_tc) {
        l->input = (float*)l->dalloc(sizeof(float) * (size_t)N * H * W * 4);
        EVK_REQUIRE(l->input != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
        x = l->input;
    }
    l->taps.assign(5, LpTap());
    // mixed operands (conv.cuh, ConvParams::mixed): every non-stem convolution has cin % 64 == 0 and reads a companion written by
    // the previous convolution's epilogue or by the max pooling, so the whole chain after the stem runs MIXED (EVK_MIXED=0: bf16x3)
    const char* mxe = getenv("EVK_MIXED");
    const bool mixed_on = tc && !(mxe && mxe[0] == '0');
    auto layer_mixed = [&](int i) { return mixed_on && i > 0 && i < n_specs && specs[i].cin % 64 == 0 && (specs[i].cout + 15) / 16 * 16 % 32 == 0; };
    for (int i = 0; i < n_specs; ++i) {
        const LpSpec& s = specs[i];
        const bool next_tc = tc;                                   // every non-stem convolution qualifies for the tensor-core kernel
        if (s.pool_k) {
            LpLayer pl; pl.kind = 1;
            pl.C = C; pl.H = H; pl.W = W; pl.k = s.pool_k; pl.stride = s.pool_s;
            pl.Ho = (H - s.pool_k) / s.pool_s + 1; pl.Wo = (W - s.pool_k) / s.pool_s + 1;
            EVK_REQUIRE(pl.Ho > 0 && pl.Wo > 0, EVK_ERR_ARG, "evk_lpips: image %dx%d is too small for the backbone", l->H, l->W);
            const size_t n_out = (size_t)N * pl.Ho * pl.Wo * C;
            pl.in = x;
            // the pooled map is read by the next convolution only: split planes for the tensor-core kernel, fp32 otherwise
            pl.out = next_tc ? nullptr : (float*)l->dalloc(sizeof(float) * n_out);
            pl.out_s = next_tc ? (__nv_bfloat16*)l->dalloc(sizeof(__nv_bfloat16) * 2 * n_out) : nullptr;
            EVK_REQUIRE(pl.out || pl.out_s, EVK_ERR_CUDA, "evk_lpips: out of device memory");
            pl.mixed = (pl.out_s != nullptr && layer_mixed(i)) ? 1 : 0;
            l->layers.push_back(pl);
            x = pl.out; xs = pl.out_s; H = pl.Ho; W = pl.Wo;
        }
        const std::string idx = std::to_string(s.idx);
        const std::string sl = "net.slice" + std::to_string(slice_of(l->backbone, s.idx)) + "." + idx;
        const std::vector<float>* w = l->find({sl + ".weight", "features." + idx + ".weight"});
        const std::vector<float>* b = l->find({sl + ".bias", "features." + idx + ".bias"});
        EVK_REQUIRE(w && b, EVK_ERR_KEY, "missing LPIPS backbone tensor '%s.weight' / '.bias'", sl.c_str());
        EVK_REQUIRE((int64_t)w->size() == (int64_t)s.cout * s.cin * s.k * s.k && (int)b->size() == s.cout, EVK_ERR_KEY,
                    "'%s': expected [%d,%d,%d,%d]", sl.c_str(), s.cout, s.cin, s.k, s.k);
        auto wat = [&](int n, int c, int r, int q) { return (*w)[(((size_t)n * s.cin + c) * s.k + r) * s.k + q]; };
        LpLayer cl; cl.kind = 0;
        ConvParams& p = cl.cp;
        p.N = N; p.epi = EPI_LINEAR; p.act = ACT_RELU;
        const int Ho = (H + 2 * s.pad - s.k) / s.stride + 1, Wo = (W + 2 * s.pad - s.k) / s.stride + 1;
        EVK_REQUIRE(Ho > 0 && Wo > 0, EVK_ERR_ARG, "evk_lpips: image %dx%d is too small for the backbone", l->H, l->W);
        const size_t n_out = (size_t)N * Ho * Wo * s.cout;
        // who reads this layer's output: a tap and a following max-pooling read fp32, a following tensor-core convolution
        // reads the split planes -- nothing else is written
        const bool last = i + 1 == n_specs;
        const bool next_pool = !last && specs[i + 1].pool_k != 0;
        const bool want_f32 = s.tap >= 0 || next_pool || (!last && !next_tc);
        const bool want_split = !last && !next_pool && next_tc;
        float* y = want_f32 ? (float*)l->dalloc(sizeof(float) * n_out) : nullptr;
        __nv_bfloat16* ys = want_split ? (__nv_bfloat16*)l->dalloc(sizeof(__nv_bfloat16) * 2 * n_out) : nullptr;
        EVK_REQUIRE((y || !want_f32) && (ys || !want_split), EVK_ERR_CUDA, "evk_lpips: out of device memory");
        std::vector<float> wk, bias(b->begin(), b->end());
        std::vector<__nv_bfloat16> wt;
        if (i == 0 && stem_tc && l->backbone == 0) {
            // space-to-depth stem (lpips_prep_s2d_kernel): 3x3 over 64 channels, channel (rr*4 + ss)*3 + c of block tap (r, q)
            // is input tap (4r + rr, 4q + ss)
            l->stem_hb = Ho + 2; l->stem_wb = Wo + 2;
            l->input_s = (__nv_bfloat16*)l->dalloc(sizeof(__nv_bfloat16) * 2 * (size_t)N * l->stem_hb * l->stem_wb * 64);
            EVK_REQUIRE(l->input_s != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
            wk.assign((size_t)9 * 64 * s.cout, 0.f);
            for (int n = 0; n < s.cout; ++n)
                for (int c = 0; c < 3; ++c)
                    for (int R = 0; R < s.k; ++R)
                        for (int Q = 0; Q < s.k; ++Q)
                            wk[((size_t)((R / 4) * 3 + Q / 4) * 64 + ((R % 4) * 4 + Q % 4) * 3 + c) * s.cout + n] = wat(n, c, R, Q);
            p.x1 = nullptr; p.x1s = l->input_s; p.c1 = 64; p.Hin = l->stem_hb; p.Win = l->stem_wb; p.kh = p.kw = 3; p.stride = 1; p.pad = 0;
            p.Hout = Ho; p.Wout = Wo; p.cout = s.cout; p.cout_pad = (s.cout + 15) / 16 * 16;
            pack_weights_tc(wk.data(), 9 * 64, s.cout, p.cout_pad, wt);
        } else if (i == 0 && stem_tc) {
            // row-window stem (lpips_prep_rowwin_kernel): two output pixels per GEMM row, like the networks' head convolution
            const int G = 2;
            l->input_s = (__nv_bfloat16*)l->dalloc(sizeof(__nv_bfloat16) * 2 * (size_t)N * H * (W + 8) * 8);
            EVK_REQUIRE(l->input_s != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
            std::vector<float> w3((size_t)s.k * s.k * 3 * s.cout), wr;
            for (int n = 0; n < s.cout; ++n)
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < s.k; ++r)
                        for (int q = 0; q < s.k; ++q) w3[((size_t)(r * s.k + q) * 3 + c) * s.cout + n] = wat(n, c, r, q);
            pack_head_weights_rowwin(w3.data(), s.k, s.k, 3, s.cout, G, wr);
            bias.resize((size_t)G * s.cout);
            for (int g = 1; g < G; ++g) std::copy(b->begin(), b->end(), bias.begin() + (size_t)g * s.cout);
            p.x1 = nullptr; p.x1s = l->input_s; p.c1 = 64; p.kw_packed = s.k; p.kw_group = G;
            p.Hin = p.Hout = H; p.Win = p.Wout = W / G; p.kh = s.k; p.kw = 1; p.stride = 1; p.pad = s.k / 2;
            p.cout = G * s.cout; p.cout_pad = G * s.cout;
            pack_weights_tc(wr.data(), s.k * 64, G * s.cout, p.cout_pad, wt);
        } else {
            const int cin_p = s.cin == 3 ? 4 : s.cin;          // the fp32 stem reads the NHWC4 input (4th channel zero)
            const int K = s.k * s.k * cin_p;
            wk.assign((size_t)K * s.cout, 0.f);
            for (int n = 0; n < s.cout; ++n)
                for (int c = 0; c < s.cin; ++c)
                    for (int r = 0; r < s.k; ++r)
                        for (int q = 0; q < s.k; ++q) wk[((size_t)(r * s.k + q) * cin_p + c) * s.cout + n] = wat(n, c, r, q);
            p.x1 = x; p.c1 = cin_p; p.Hin = H; p.Win = W; p.kh = p.kw = s.k; p.stride = s.stride; p.pad = s.pad;
            p.Hout = Ho; p.Wout = Wo; p.cout = s.cout;
            if (tc && xs != nullptr && tc_eligible(p)) {
                p.x1s = xs;
                p.cout_pad = (s.cout + 15) / 16 * 16;
                pack_weights_tc(wk.data(), K, s.cout, p.cout_pad, wt);
                if (layer_mixed(i)) {
                    EVK_REQUIRE(tc_mixed_capable(p), EVK_ERR_STATE, "evk_lpips: layer %d cannot run with mixed operands", i);
                    std::vector<__nv_bfloat16> wm;
                    std::vector<float> isc;
                    pack_weights_mixed(wk.data(), K, s.cout, p.cout_pad, wm, isc);
                    void* dm = l->dalloc(wm.size() * sizeof(__nv_bfloat16));
                    float* di = (float*)l->dalloc(isc.size() * sizeof(float));
                    EVK_REQUIRE(dm != nullptr && di != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
                    cudaMemcpy(dm, wm.data(), wm.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
                    cudaMemcpy(di, isc.data(), isc.size() * sizeof(float), cudaMemcpyHostToDevice);
                    p.w_mx = (const __nv_bfloat16*)dm; p.w_iscale = di; p.mixed = 1;
                }
            } else {
                EVK_REQUIRE(x != nullptr, EVK_ERR_STATE, "evk_lpips: layer %d has no fp32 input for the CUDA-core kernel", i);
                float* dw = (float*)l->dalloc(sizeof(float) * wk.size());
                EVK_REQUIRE(dw != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
                cudaMemcpy(dw, wk.data(), sizeof(float) * wk.size(), cudaMemcpyHostToDevice);
                p.w = dw;
                if (y == nullptr) {                              // the CUDA-core kernel writes fp32 only
                    y = (float*)l->dalloc(sizeof(float) * n_out);
                    EVK_REQUIRE(y != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
                }
            }
        }
        float* db = (float*)l->dalloc(sizeof(float) * bias.size());
        EVK_REQUIRE(db != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
        cudaMemcpy(db, bias.data(), sizeof(float) * bias.size(), cudaMemcpyHostToDevice);
        p.bias = db; p.y = y; p.ys = ys;
        p.ys_mixed = (ys != nullptr && layer_mixed(i + 1) && s.cout % 64 == 0) ? 1 : 0;
        if (!wt.empty()) {
            void* d = l->dalloc(wt.size() * sizeof(__nv_bfloat16));
            EVK_REQUIRE(d != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
            cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
            p.w_tc = (const __nv_bfloat16*)d;
            int r = tc_plan_create(p);
            if (r != EVK_OK) return r;
            l->plans.push_back(p.tc);
        }
        cl.flops = 2.0 * s.cout * s.cin * s.k * s.k * (double)N * Ho * Wo;
        l->flops += cl.flops;
        l->layers.push_back(cl);
        x = y; xs = ys; H = Ho; W = Wo; C = s.cout;
        if (s.tap >= 0) {
            const std::string t = std::to_string(s.tap);
            const std::vector<float>* lin = l->find({"lin" + t + ".model.1.weight", "lins." + t + ".model.1.weight"});
            EVK_REQUIRE(lin && (int)lin->size() == C, EVK_ERR_KEY, "missing LPIPS linear layer 'lin%s.model.1.weight' [1,%d,1,1]", t.c_str(), C);
            float* dl = (float*)l->dalloc(sizeof(float) * C);
            EVK_REQUIRE(dl != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
            cudaMemcpy(dl, lin->data(), sizeof(float) * C, cudaMemcpyHostToDevice);
            l->taps[s.tap] = LpTap{y, C, H, W, dl};
        }
    }
    l->part = (double*)l->dalloc(sizeof(double) * 5 * l->batch * kTapChunks);
    EVK_REQUIRE(l->part != nullptr, EVK_ERR_CUDA, "evk_lpips: out of device memory");
    return EVK_OK;
}

}  // namespace evk

extern "C" {

int evk_lpips_create(int backbone, int batch, int H, int W, int precision, evk_lpips** out) {
    EVK_REQUIRE(out && (backbone == 0 || backbone == 1) && batch >= 1 && H > 0 && W > 0, EVK_ERR_ARG,
                "evk_lpips_create: bad argument (backbone 0 = alex, 1 = vgg16)");
    int ndev = 0;
    EVK_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    EVK_REQUIRE(ndev > 0, EVK_ERR_CUDA, "evk_lpips_create: no CUDA device (there is no CPU implementation)");
    evk_lpips* l = new evk_lpips();
    l->backbone = backbone; l->batch = batch; l->H = H; l->W = W; l->precision = precision;
    *out = l;
    return EVK_OK;
}

int evk_lpips_load_tensor(evk_lpips* l, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    EVK_REQUIRE(l && name && host_data && ndim >= 0 && ndim <= 8, EVK_ERR_ARG, "evk_lpips_load_tensor: bad argument");
    EVK_REQUIRE(!l->finalized, EVK_ERR_STATE, "evk_lpips_load_tensor: already finalized");
    int64_t n = 1;
    std::vector<int64_t> sh;
    for (int i = 0; i < ndim; ++i) { sh.push_back(shape[i]); n *= shape[i]; }
    l->sd[name].assign(host_data, host_data + n);
    l->shapes[name] = sh;
    return EVK_OK;
}

int evk_lpips_finalize(evk_lpips* l) {
    EVK_REQUIRE(l && !l->finalized, EVK_ERR_STATE, "evk_lpips_finalize: bad state");
    int r = lpips_build(l);
    if (r != EVK_OK) return r;
    EVK_CHECK_CUDA(cudaDeviceSynchronize());
    l->sd.clear();
    l->finalized = true;
    return EVK_OK;
}

int evk_lpips_forward(evk_lpips* l, const float* img, const float* ref, int n, double* scores, void* stream) {
    EVK_REQUIRE(l && l->finalized, EVK_ERR_STATE, "evk_lpips_forward: not finalized");
    EVK_REQUIRE(img && ref && scores && n >= 1 && n <= l->batch, EVK_ERR_ARG, "evk_lpips_forward: bad argument (1 <= n <= %d)", l->batch);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t pixels = (int64_t)l->H * l->W;
    const int N = 2 * l->batch;
    if (l->input != nullptr) {
        lpips_prep_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(N * pixels, 256), 2368), 256, 0, st>>>(img, ref, n, l->batch, pixels, l->input);
    } else if (l->backbone == 0) {
        const int64_t total = (int64_t)N * l->stem_hb * l->stem_wb * 16;
        lpips_prep_s2d_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, 0, st>>>(img, ref, n, l->batch, l->H, l->W, l->stem_hb,
                                                                                                 l->stem_wb, l->input_s);
    } else {
        const int64_t total = (int64_t)N * l->H * (l->W + 8);
        lpips_prep_rowwin_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, 0, st>>>(img, ref, n, l->batch, l->H, l->W, l->input_s);
    }
    EVK_CHECK_CUDA(cudaGetLastError());
    for (const LpLayer& ly : l->layers) {
        if (ly.kind == 0) {
            int r = launch_conv(ly.cp, l->precision, st);
            if (r != EVK_OK) return r;
        } else {
            const int64_t total = (int64_t)N * ly.Ho * ly.Wo * (ly.C / 4);
            maxpool_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, 0, st>>>(ly.in, ly.out, ly.out_s, N, ly.H, ly.W,
                                                                                                ly.C / 4, ly.k, ly.stride, ly.Ho, ly.Wo, ly.mixed);
            EVK_CHECK_CUDA(cudaGetLastError());
        }
    }
    LpPixels px;
    for (int t = 0; t < 5; ++t) {
        const LpTap& tp = l->taps[t];
        px.n[t] = tp.H * tp.W;
        static const bool tap_v1 = getenv("EVK_LPIPS_TAP_V1") != nullptr;
        const dim3 tg((unsigned)n, kTapChunks);
        double* tpart = l->part + (size_t)t * l->batch * kTapChunks;
        const int px_n = tp.H * tp.W;
        if (tap_v1) lpips_tap_kernel_v1<<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tp.C, tpart);
        else switch (tp.C) {
            case 64: lpips_tap_kernel<1, 16><<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tpart); break;
            case 128: lpips_tap_kernel<1, 32><<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tpart); break;
            case 192: lpips_tap_kernel<3, 16><<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tpart); break;
            case 256: lpips_tap_kernel<2, 32><<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tpart); break;
            case 384: lpips_tap_kernel<3, 32><<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tpart); break;
            case 512: lpips_tap_kernel<4, 32><<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tpart); break;
            default: lpips_tap_kernel_v1<<<tg, 256, 0, st>>>(tp.feat, tp.lin, l->batch, px_n, tp.C, tpart); break;
        }
        EVK_CHECK_CUDA(cudaGetLastError());
    }
    lpips_sum_kernel<<<ceil_div(n, 64), 64, 0, st>>>(l->part, 5, l->batch, n, px, scores);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

double evk_lpips_flops(evk_lpips* l) { return l ? l->flops : 0.0; }

int evk_lpips_num_tc_layers(evk_lpips* l) { return l ? (int)l->plans.size() : EVK_ERR_ARG; }

int evk_lpips_destroy(evk_lpips* l) {
    if (!l) return EVK_OK;
    for (TcPlan* pl : l->plans) tc_plan_destroy(pl);
    l->arena.release();
    delete l;
    return EVK_OK;
}

}  // extern "C"
