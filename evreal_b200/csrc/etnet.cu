// ET-Net (model/eitr/u_trans.py, transformer_encoder.py, transformer_decoder.py) token-path kernels.  Tokens are the pixels of
// the 1/8-resolution map in row-major order, so a token tensor [N, L, 256] IS an NHWC activation [N, h, w, 256]: every linear
// layer (attention in / out projections, feed-forward) is a 1x1 convolution on the tensor-core kernel of conv_tc.cu with the
// residual add in its epilogue.  What remains is here: LayerNorm, multi-head attention (fp32: softmax(q k^T / sqrt(d)) v),
// the sine position table and the average of the six token sets.
#include <algorithm>
#include <cmath>
#include <vector>

#include "conv.cuh"
#include "tc.cuh"

namespace evk {

// LayerNorm over C = 256 channels (nn.LayerNorm, eps 1e-5, biased variance): a warp per token, 8 channels per lane.
__global__ void __launch_bounds__(256) layernorm256_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ out, __nv_bfloat16* __restrict__ out_s, int64_t T) {
    const int lane = threadIdx.x & 31;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + lane * 2 + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane * 2), b1 = __ldg(reinterpret_cast<const float4*>(beta) + lane * 2 + 1);
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < T; t += (int64_t)gridDim.x * 8) {
        const float4* src = reinterpret_cast<const float4*>(x + t * 256) + lane * 2;
        const float4 a0 = __ldg(src), a1 = __ldg(src + 1);
        float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.0f / 256.0f);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / 256.0f) + 1e-5f);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf((v[i] - mean) * rstd, gg[i], bb[i]);
        const int64_t o = t * 256 + lane * 8;
        if (out != nullptr) {
            *reinterpret_cast<float4*>(out + o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(out + o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (out_s != nullptr) {
            __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_bf16(v[i], hi[i], lo[i]);
            *reinterpret_cast<uint4*>(out_s + o) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(out_s + T * 256 + o) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

int launch_layernorm256(const float* x, const float* gamma, const float* beta, float* out, __nv_bfloat16* out_s, int64_t T, cudaStream_t st) {
    EVK_REQUIRE(x && gamma && beta && (out || out_s) && T > 0, EVK_ERR_ARG, "layernorm: bad argument");
    layernorm256_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(T, 8), 2368), 256, 0, st>>>(x, gamma, beta, out, out_s, T);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// nn.MultiheadAttention core (eval, no masks), 8 heads of 32: out[n, i, h*32 + d] = sum_j softmax_j(q_i . k_j / sqrt(32)) v_j[d].
// q / k / v are row-strided views (the packed in-projection output): element (n, token, h*32 + d) at ptr[(n*L + token)*stride + h*32 + d].
// CTA = 128 queries of one (sample, head), a thread per query; keys / values stream through shared memory in tiles of 32
// (online softmax: one rescale per tile).  fp32 throughout -- the probabilities are the one place where a bf16 split would
// need a fourth product.
constexpr int kHd = 32, kKeyTile = 32;
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ q, int q_stride, const float* __restrict__ k, const float* __restrict__ v,
                                                        int kv_stride, float* __restrict__ out, __nv_bfloat16* __restrict__ out_s, int Lq, int Lk,
                                                        int64_t out_plane, float scale) {
    __shared__ __align__(16) float sk[kKeyTile][kHd], sv[kKeyTile][kHd];
    const int n = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * 128 + threadIdx.x;
    const bool ok = qi < Lq;
    float qr[kHd], acc[kHd];
    {
        const float4* qp = reinterpret_cast<const float4*>(q + ((int64_t)n * Lq + (ok ? qi : 0)) * q_stride + h * kHd);
#pragma unroll
        for (int d = 0; d < kHd / 4; ++d) {
            const float4 t = __ldg(qp + d);
            qr[4 * d] = t.x * scale; qr[4 * d + 1] = t.y * scale; qr[4 * d + 2] = t.z * scale; qr[4 * d + 3] = t.w * scale;
        }
    }
#pragma unroll
    for (int d = 0; d < kHd; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < Lk; j0 += kKeyTile) {
        __syncthreads();
        for (int i = threadIdx.x; i < kKeyTile * (kHd / 4); i += 128) {
            const int r = i / (kHd / 4), c = i % (kHd / 4);
            float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
            if (j0 + r < Lk) {
                kk = __ldg(reinterpret_cast<const float4*>(k + ((int64_t)n * Lk + j0 + r) * kv_stride + h * kHd) + c);
                vv = __ldg(reinterpret_cast<const float4*>(v + ((int64_t)n * Lk + j0 + r) * kv_stride + h * kHd) + c);
            }
            *reinterpret_cast<float4*>(&sk[r][c * 4]) = kk;
            *reinterpret_cast<float4*>(&sv[r][c * 4]) = vv;
        }
        __syncthreads();
        const int nk = min(kKeyTile, Lk - j0);
        float s[kKeyTile];
        float tmax = -INFINITY;
#pragma unroll
        for (int j = 0; j < kKeyTile; ++j) {
            float d0 = 0.f;
#pragma unroll
            for (int d = 0; d < kHd; ++d) d0 = fmaf(qr[d], sk[j][d], d0);
            s[j] = j < nk ? d0 : -INFINITY;
            tmax = fmaxf(tmax, s[j]);
        }
        const float m_new = fmaxf(m, tmax);
        const float corr = __expf(m - m_new);               // exp(-inf) = 0 on the first tile
        l *= corr;
#pragma unroll
        for (int d = 0; d < kHd; ++d) acc[d] *= corr;
#pragma unroll
        for (int j = 0; j < kKeyTile; ++j) {
            const float p = __expf(s[j] - m_new);            // 0 for the padded keys
            l += p;
#pragma unroll
            for (int d = 0; d < kHd; ++d) acc[d] = fmaf(p, sv[j][d], acc[d]);
        }
        m = m_new;
    }
    if (!ok) return;
    const float inv = 1.0f / l;
    const int64_t o = ((int64_t)n * Lq + qi) * 256 + h * kHd;
#pragma unroll
    for (int d = 0; d < kHd; ++d) acc[d] *= inv;
    if (out != nullptr) {
#pragma unroll
        for (int d = 0; d < kHd / 4; ++d) *reinterpret_cast<float4*>(out + o + 4 * d) = make_float4(acc[4 * d], acc[4 * d + 1], acc[4 * d + 2], acc[4 * d + 3]);
    }
    if (out_s != nullptr) {
#pragma unroll
        for (int d = 0; d < kHd / 8; ++d) {
            __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16(acc[8 * d + e], hi[e], lo[e]);
            *reinterpret_cast<uint4*>(out_s + o + 8 * d) = *reinterpret_cast<const uint4*>(hi);
            *reinterpret_cast<uint4*>(out_s + out_plane + o + 8 * d) = *reinterpret_cast<const uint4*>(lo);
        }
    }
}

int launch_attention(const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* out, __nv_bfloat16* out_s, int N, int Lq,
                     int Lk, cudaStream_t st) {
    EVK_REQUIRE(q && k && v && (out || out_s) && N > 0 && Lq > 0 && Lk > 0 && q_stride % 4 == 0 && kv_stride % 4 == 0, EVK_ERR_ARG, "attention: bad argument");
    dim3 grid((unsigned)ceil_div(Lq, 128), 8, (unsigned)N);
    attention_kernel<<<grid, 128, 0, st>>>(q, q_stride, k, v, kv_stride, out, out_s, Lq, Lk, (int64_t)N * Lq * 256, 1.0f / sqrtf((float)kHd));
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// out[n, t, c] = x[n, t, c] + pos[t, c]   (TransformerEncoder.with_embed: the sine table is added once, before the first layer)
__global__ void __launch_bounds__(256) add_pos_kernel(const float* __restrict__ x, const float* __restrict__ pos, float* __restrict__ out, int64_t n4,
                                                      int64_t per_sample4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i), b = __ldg(reinterpret_cast<const float4*>(pos) + i % per_sample4);
        reinterpret_cast<float4*>(out)[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

int launch_add_pos(const float* x, const float* pos, float* out, int N, int64_t per_sample, cudaStream_t st) {
    EVK_REQUIRE(x && pos && out && per_sample % 4 == 0, EVK_ERR_ARG, "add_pos: bad argument");
    const int64_t n4 = (int64_t)N * per_sample / 4;
    add_pos_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n4, 256), 2368), 256, 0, st>>>(x, pos, out, n4, per_sample / 4);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// (hs0 + hs1 + hs2 + hc0 + hc1 + hc2) / 6 in the reference's left-to-right order (u_trans.py:108)
struct Six { const float* p[6]; };
__global__ void __launch_bounds__(256) avg6_kernel(const __grid_constant__ Six s, float* __restrict__ out, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldg(reinterpret_cast<const float4*>(s.p[0]) + i);
#pragma unroll
        for (int k = 1; k < 6; ++k) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(s.p[k]) + i);
            a.x = __fadd_rn(a.x, b.x); a.y = __fadd_rn(a.y, b.y); a.z = __fadd_rn(a.z, b.z); a.w = __fadd_rn(a.w, b.w);
        }
        reinterpret_cast<float4*>(out)[i] = make_float4(__fdiv_rn(a.x, 6.0f), __fdiv_rn(a.y, 6.0f), __fdiv_rn(a.z, 6.0f), __fdiv_rn(a.w, 6.0f));
    }
}

int launch_avg6(const float* const* six, float* out, int64_t n, cudaStream_t st) {
    EVK_REQUIRE(six && out && n % 4 == 0, EVK_ERR_ARG, "avg6: bad argument");
    Six s;
    for (int k = 0; k < 6; ++k) { s.p[k] = six[k]; EVK_REQUIRE(six[k] != nullptr, EVK_ERR_ARG, "avg6: null input"); }
    avg6_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n / 4, 256), 2368), 256, 0, st>>>(s, out, n / 4);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// model/eitr/position_encoding.py:15-23: table[pos][j] = sin / cos(pos / 10000^(2*(j/2)/d)), float64 then float32
void sine_position_table(int n, int d, std::vector<float>& out) {
    out.resize((size_t)n * d);
    for (int p = 0; p < n; ++p)
        for (int j = 0; j < d; ++j) {
            const double ang = (double)p / std::pow(10000.0, 2.0 * (double)(j / 2) / (double)d);
            out[(size_t)p * d + j] = (float)((j & 1) ? std::cos(ang) : std::sin(ang));
        }
}

}  // namespace evk
