// SPADE-E2VID (model/spade_e2v.py, Unet6) glue kernels: everything around its convolutions and ConvLSTMs, which run on
// conv_tc.cu like every other network.
//
//   skip add               x + x_k before the pixel-shuffle decoders and the recurrent last decoder (:160-162)
//   UpConvLayer3           conv3x3 (C -> 4*Cout, no bias) -> PixelShuffle(2) -> SPADE -> ReLU (:79-110): the shuffle, the
//                          parameter-free BatchNorm (eval mode: running statistics), the modulation
//                          normalized * (1 + gamma) + beta (:44-76) and the ReLU are ONE elementwise kernel here; gamma and
//                          beta come from one convolution with 2*Cout outputs
//   prediction             conv_img(relu(x + head)) -> bn_img -> sigmoid = prev_recs (3 channels); image = their mean (:166-171)
//   first frame            x_org = x[:, :3] shifted / scaled to [0, 1] IN PLACE (:140-145; x_org is a view of the input, so
//                          the head convolution sees the change) -- per sample here (the reference only runs batch 1)
#include <algorithm>

#include "conv.cuh"
#include "tc.cuh"

namespace evk {

// out = x + s (fp32, optional) and its split-bf16 planes (optional)
__global__ void __launch_bounds__(256) add_split_kernel(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ out,
                                                        __nv_bfloat16* __restrict__ out_s, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i), b = __ldg(reinterpret_cast<const float4*>(s) + i);
        const float f[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
        if (out != nullptr) reinterpret_cast<float4*>(out)[i] = make_float4(f[0], f[1], f[2], f[3]);
        if (out_s != nullptr) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16(f[e], hi[e], lo[e]);
            *reinterpret_cast<uint2*>(out_s + i * 4) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(out_s + n4 * 4 + i * 4) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

int launch_add_split(const float* x, const float* s, float* out, __nv_bfloat16* out_s, int64_t n, cudaStream_t st) {
    EVK_REQUIRE(x && s && (out || out_s) && n > 0 && n % 4 == 0, EVK_ERR_ARG, "add_split: bad argument");
    add_split_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n / 4, 256), 2368), 256, 0, st>>>(x, s, out, out_s, n / 4);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// c0 [N,h,w,4*C] (conv output, channel c*4 + i*2 + j = phase (i, j) of output channel c: nn.PixelShuffle(2)),
// gb [N,2h,2w,2*C] (gamma | beta), alpha / shift [C] (eval BatchNorm without affine: v * alpha + shift, alpha = 1/sqrt(var+eps),
// shift = -mean * alpha) -> out [N,2h,2w,C] = relu((v * alpha + shift) * (1 + gamma) + beta), fp32 and / or split planes.
// A thread takes 16 consecutive input channels (4 output channels x 4 phases) of one input pixel.
__global__ void __launch_bounds__(256) spade_shuffle_kernel(const float* __restrict__ c0, const float* __restrict__ gb, const float* __restrict__ alpha,
                                                            const float* __restrict__ shift, float* __restrict__ out, __nv_bfloat16* __restrict__ out_s,
                                                            int N, int h, int w, int C) {
    const int C4 = C / 4;
    const int64_t total = (int64_t)N * h * w * C4;
    const int64_t plane = (int64_t)N * (2 * h) * (2 * w) * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int cg = (int)(t % C4);
        const int x = (int)((t / C4) % w);
        const int y = (int)((t / ((int64_t)C4 * w)) % h);
        const int n = (int)(t / ((int64_t)C4 * w * h));
        const float4* src = reinterpret_cast<const float4*>(c0 + ((((int64_t)n * h + y) * w + x) * 4 * C + (int64_t)cg * 16));
        float v[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float4 a = __ldg(src + q); v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w; }
        const float4 al = __ldg(reinterpret_cast<const float4*>(alpha) + cg), sh = __ldg(reinterpret_cast<const float4*>(shift) + cg);
        const float a4[4] = {al.x, al.y, al.z, al.w}, s4[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int Y = 2 * y + (ph >> 1), X = 2 * x + (ph & 1);
            const int64_t opix = ((int64_t)n * (2 * h) + Y) * (2 * w) + X;
            const float4 g = __ldg(reinterpret_cast<const float4*>(gb + opix * 2 * C) + cg);
            const float4 b = __ldg(reinterpret_cast<const float4*>(gb + opix * 2 * C + C) + cg);
            const float g4[4] = {g.x, g.y, g.z, g.w}, b4[4] = {b.x, b.y, b.z, b.w};
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float nrm = fmaf(v[c * 4 + ph], a4[c], s4[c]);
                o[c] = fmaxf(__fadd_rn(__fmul_rn(nrm, __fadd_rn(1.0f, g4[c])), b4[c]), 0.0f);
            }
            const int64_t oo = opix * C + (int64_t)cg * 4;
            if (out != nullptr) *reinterpret_cast<float4*>(out + oo) = make_float4(o[0], o[1], o[2], o[3]);
            if (out_s != nullptr) {
                __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_bf16(o[e], hi[e], lo[e]);
                *reinterpret_cast<uint2*>(out_s + oo) = *reinterpret_cast<const uint2*>(hi);
                *reinterpret_cast<uint2*>(out_s + plane + oo) = *reinterpret_cast<const uint2*>(lo);
            }
        }
    }
}

int launch_spade_shuffle(const float* c0, const float* gb, const float* alpha, const float* shift, float* out, __nv_bfloat16* out_s, int N,
                         int h, int w, int C, cudaStream_t st) {
    EVK_REQUIRE(c0 && gb && alpha && shift && (out || out_s) && C % 4 == 0, EVK_ERR_ARG, "spade_shuffle: bad argument");
    const int64_t total = (int64_t)N * h * w * (C / 4);
    spade_shuffle_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, 0, st>>>(c0, gb, alpha, shift, out, out_s, N, h, w, C);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// t = relu(x + head) [N,H,W,32]; p_k = sigmoid(sum_c w[c][k] * t[c] + b[k]), k < 3 (conv_img + bn_img folded);
// prev [N,3,H,W] = p; image [N,1,H,W] = (p_0 + p_1 + p_2) / 3
__global__ void __launch_bounds__(256) spade_pred_kernel(const float* __restrict__ x, const float* __restrict__ head, const float* __restrict__ w,
                                                         float b0, float b1, float b2, float* __restrict__ prev, float* __restrict__ image,
                                                         int N, int64_t HW, int C) {
    const int64_t total = (int64_t)N * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float4* xp = reinterpret_cast<const float4*>(x + i * C);
        const float4* hp = reinterpret_cast<const float4*>(head + i * C);
        float acc[3] = {0.f, 0.f, 0.f};
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 a = __ldg(xp + c4), h4 = __ldg(hp + c4);
            const float t[4] = {fmaxf(a.x + h4.x, 0.f), fmaxf(a.y + h4.y, 0.f), fmaxf(a.z + h4.z, 0.f), fmaxf(a.w + h4.w, 0.f)};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[k] = fmaf(t[e], __ldg(w + (c4 * 4 + e) * 3 + k), acc[k]);
        }
        const float p0 = sigmoidf_(acc[0] + b0), p1 = sigmoidf_(acc[1] + b1), p2 = sigmoidf_(acc[2] + b2);
        const int64_t n = i / HW, pix = i - n * HW;
        prev[(n * 3 + 0) * HW + pix] = p0;
        prev[(n * 3 + 1) * HW + pix] = p1;
        prev[(n * 3 + 2) * HW + pix] = p2;
        image[i] = __fdiv_rn(__fadd_rn(__fadd_rn(p0, p1), p2), 3.0f);
    }
}

int launch_spade_pred(const float* x, const float* head, const float* w, const float* bias_host3, float* prev, float* image, int N, int64_t HW,
                      int C, cudaStream_t st) {
    EVK_REQUIRE(x && head && w && prev && image && C % 4 == 0, EVK_ERR_ARG, "spade_pred: bad argument");
    spade_pred_kernel<<<(unsigned)std::min<int64_t>(ceil_div64((int64_t)N * HW, 256), 2368), 256, 0, st>>>(x, head, w, bias_host3[0], bias_host3[1],
                                                                                                         bias_host3[2], prev, image, N, HW, C);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// first frame after reset_states(): per sample, over the first three bins of the (padded) event tensor [N,bins,H,W]:
// v -= min; if (max > 0) v /= max   (in place), and a copy into prev [N,3,H,W]
__global__ void __launch_bounds__(1024) spade_first_frame_kernel(float* __restrict__ in, float* __restrict__ prev, int bins, int64_t HW) {
    __shared__ float s_min[32], s_max[32];
    float* v = in + (int64_t)blockIdx.x * bins * HW;
    float* o = prev + (int64_t)blockIdx.x * 3 * HW;
    const int64_t n = 3 * HW;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += 1024) { const float a = v[i]; mn = fminf(mn, a); mx = fmaxf(mx, a); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = s_min[0]; mx = s_max[0];
    for (int k = 1; k < 32; ++k) { mn = fminf(mn, s_min[k]); mx = fmaxf(mx, s_max[k]); }
    const float range = __fsub_rn(mx, mn);                    // the maximum after the shift
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        float a = __fsub_rn(v[i], mn);
        if (range > 0.0f) a = __fdiv_rn(a, range);
        v[i] = a;
        o[i] = a;
    }
}

int launch_spade_first_frame(float* in, float* prev, int N, int bins, int64_t HW, cudaStream_t st) {
    EVK_REQUIRE(in && prev && bins >= 3, EVK_ERR_ARG, "spade_first_frame: needs at least 3 bins");
    spade_first_frame_kernel<<<N, 1024, 0, st>>>(in, prev, bins, HW);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
