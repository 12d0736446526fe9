// HyperE2VID dynamic decoder: context fusion input, per-pixel atom generation
// and per-pixel dynamic convolution.  The reference materialises a
// [N, 256*25, h, w] im2col (nn.functional.unfold) and permutes it for a bmm
// (model/hyper/hyper_dynamic.py:87-91, 60 % of its CPU time); here the 5x5
// neighbourhood is read from a shared-memory halo tile instead and only the
// [N,h,w,C*A] intermediate that feeds the 1x1 compositional conv is written.
#include "hyper.cuh"
#include "tc.cuh"

namespace evk {

// (0) context = interpolate(cat(ev, prev), scale 0.25, bilinear, align_corners=False)
// src = 4*(dst+0.5)-0.5 = 4*dst+1.5 -> taps 4*dst+1 and 4*dst+2 with weight 0.5 each.
__global__ void __launch_bounds__(256) hyper_context_kernel(const HyperParams p) {
    const int h = p.H / 4, w = p.W / 4;
    const int64_t total = (int64_t)p.N * h * w * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % 8);
        const int x = (int)((i / 8) % w);
        const int y = (int)((i / (8 * (int64_t)w)) % h);
        const int n = (int)(i / (8 * (int64_t)w * h));
        float v = 0.f;
        if (ch <= p.bins) {
            const float* src = (ch < p.bins) ? p.ev_nchw + ((size_t)n * p.bins + ch) * p.H * p.W : p.prev + (size_t)n * p.H * p.W;
            const int y0 = 4 * y + 1, x0 = 4 * x + 1;
            const float v00 = src[(size_t)y0 * p.W + x0], v01 = src[(size_t)y0 * p.W + x0 + 1];
            const float v10 = src[(size_t)(y0 + 1) * p.W + x0], v11 = src[(size_t)(y0 + 1) * p.W + x0 + 1];
            v = 0.5f * (0.5f * v00 + 0.5f * v01) + 0.5f * (0.5f * v10 + 0.5f * v11);
        }
        p.ctx[i] = v;
    }
}

// (1) atoms[pix][a][l] = sum_k coef[pix][a*K+k] * bases[k][l]
__global__ void __launch_bounds__(256) hyper_atoms_kernel(const HyperParams p) {
    extern __shared__ float sb[];   // bases [K][L]
    for (int i = threadIdx.x; i < p.K * p.L; i += blockDim.x) sb[i] = p.bases[i];
    __syncthreads();
    const int64_t total = (int64_t)p.N * p.h * p.w * p.A * p.L;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(i % p.L);
        const int a = (int)((i / p.L) % p.A);
        const int64_t pix = i / ((int64_t)p.L * p.A);
        const float* c = p.coef + pix * (p.A * p.K) + a * p.K;
        float acc = 0.f;
        for (int k = 0; k < p.K; ++k) acc = fmaf(c[k], sb[k * p.L + l], acc);
        p.atoms[i] = acc;
    }
}

// (2) dynamic per-pixel convolution.  CTA = 8x4 pixels x 32 channels.
constexpr int kATW = 8, kATH = 4, kACH = 32;

template <int A, int KS>
__global__ void __launch_bounds__(256) hyper_apply_kernel(const HyperParams p) {
    constexpr int L = KS * KS, R = KS / 2;
    constexpr int HW_ = kATW + 2 * R, HH_ = kATH + 2 * R;
    __shared__ __align__(16) float s_x[HH_ * HW_][kACH];
    __shared__ float s_a[kATH * kATW][A * L];
    const int tid = threadIdx.x;
    const int chunks = p.C / kACH;
    const int n = blockIdx.z / chunks, c0 = (blockIdx.z % chunks) * kACH;
    const int y0 = blockIdx.y * kATH, x0 = blockIdx.x * kATW;
    for (int i = tid; i < HH_ * HW_ * (kACH / 4); i += 256) {
        const int c4 = i % (kACH / 4), pix = i / (kACH / 4);
        const int gy = y0 + pix / HW_ - R, gx = x0 + pix % HW_ - R;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w)
            v = __ldg(reinterpret_cast<const float4*>(p.xu + (((size_t)n * p.h + gy) * p.w + gx) * p.C + c0) + c4);
        *reinterpret_cast<float4*>(&s_x[pix][c4 * 4]) = v;
    }
    for (int i = tid; i < kATH * kATW * A * L; i += 256) {
        const int q = i % (A * L), pix = i / (A * L);
        const int gy = y0 + pix / kATW, gx = x0 + pix % kATW;
        s_a[pix][q] = (gy < p.h && gx < p.w) ? p.atoms[(((size_t)n * p.h + gy) * p.w + gx) * (A * L) + q] : 0.f;
    }
    __syncthreads();
    const int pix = tid / 8, lane = tid % 8;           // 32 pixels x 8 channel lanes (4 channels each)
    const int py = pix / kATW, px = pix % kATW;
    const int gy = y0 + py, gx = x0 + px;
    float acc[4][A];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int a = 0; a < A; ++a) acc[c][a] = 0.f;
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const int dy = l / KS, dx = l % KS;
        const float4 xv = *reinterpret_cast<const float4*>(&s_x[(py + dy) * HW_ + px + dx][lane * 4]);
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float at = s_a[pix][a * L + l];
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c][a] = fmaf(at, xs[c], acc[c][a]);
        }
    }
    if (gy < p.h && gx < p.w) {
        const size_t off = (((size_t)n * p.h + gy) * p.w + gx) * ((size_t)p.C * A) + (size_t)(c0 + lane * 4) * A;
        // the thread's 4 channels x A atoms are 4*A contiguous values of the [.., C*A] pixel record: vector stores
        float vals[4 * A];
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int a = 0; a < A; ++a) vals[c * A + a] = acc[c][a];
        static_assert((4 * A) % 8 == 0, "vector stores need 4*A to be a multiple of 8");
        if (p.inter != nullptr) {
#pragma unroll
            for (int q = 0; q < 4 * A; q += 4)
                *reinterpret_cast<float4*>(p.inter + off + q) = make_float4(vals[q], vals[q + 1], vals[q + 2], vals[q + 3]);
        }
        if (p.inter_s != nullptr) {
            const size_t plane = (size_t)p.N * p.h * p.w * p.C * A;
#pragma unroll
            for (int q = 0; q < 4 * A; q += 8) {
                __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split_bf16(vals[q + e], hi[e], lo[e]);
                *reinterpret_cast<uint4*>(p.inter_s + off + q) = *reinterpret_cast<const uint4*>(hi);
                *reinterpret_cast<uint4*>(p.inter_s + plane + off + q) = *reinterpret_cast<const uint4*>(lo);
            }
        }
    }
}

// (3) dynamic convolution applied to U = conv1x1(xu) (see hyper.cuh).  A warp owns a strip of kSX x kSY pixels and all CO = 128
// output channels (4 per lane): every U value it loads (one 512-byte row of the warp per pixel and atom) feeds up to
// 5 x 4 = 20 of its pixels from registers, so U is read ~3x from L2 instead of 25x, and the atoms of the strip -- generated
// on the fly from the 72 basis coefficients -- are warp-uniform shared-memory broadcasts.
constexpr int kSX = 8, kSY = 4, kSWarps = 8;

template <int A, int KS, int K>
__global__ void __launch_bounds__(kSWarps * 32, 1) hyper_apply_u_kernel(const HyperParams p, int strips_x, int strips_y, int n_strips) {
    constexpr int L = KS * KS, R = KS / 2, LP = KS * 8;           // atoms of one (pixel, a): KS rows of 8 slots (KS used)
    constexpr int NC = A * K;                                      // basis coefficients per pixel
    extern __shared__ __align__(16) float s_dyn[];
    float* s_bases = s_dyn;                                         // [K][L]
    float (*s_at)[kSX * kSY][LP] = reinterpret_cast<float (*)[kSX * kSY][LP]>(s_dyn + ((K * L + 3) & ~3));
    float (*s_coef)[kSX * kSY][NC] = reinterpret_cast<float (*)[kSX * kSY][NC]>(&s_at[kSWarps][0][0]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < K * L; i += blockDim.x) s_bases[i] = p.bases[i];
    __syncthreads();
    const int CO = p.CO, CU = A * CO;
    const float4 bias = __ldg(reinterpret_cast<const float4*>(p.out_bias) + lane);
    for (int s = blockIdx.x * kSWarps + warp; s < n_strips; s += gridDim.x * kSWarps) {
        const int sx = s % strips_x, sy = (s / strips_x) % strips_y, n = s / (strips_x * strips_y);
        const int x0 = sx * kSX, y0 = sy * kSY;
        float4 acc[kSY * kSX];
#pragma unroll
        for (int i = 0; i < kSY * kSX; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // the strip's basis coefficients, once, coalesced: a strip row is kSX * NC contiguous floats
        __syncwarp();
        for (int py = 0; py < kSY; ++py) {
            const int gy = min(y0 + py, p.h - 1);
            const int valid = min(kSX, p.w - x0) * NC;
            const float* src = p.coef + (((size_t)n * p.h + gy) * p.w + x0) * NC;
            for (int i = lane; i < kSX * NC; i += 32) s_coef[warp][py * kSX][i] = i < valid ? __ldg(src + i) : 0.f;
        }
#pragma unroll 1
        for (int a = 0; a < A; ++a) {
            __syncwarp();
            // atoms of this strip for atom index a: (pixel, l) pairs dealt over the lanes
            for (int i = lane; i < kSX * kSY * L; i += 32) {
                const int pix = i / L, l = i - pix * L;
                const float* c = &s_coef[warp][pix][a * K];
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) v = fmaf(c[k], s_bases[k * L + l], v);
                s_at[warp][pix][(l / KS) * 8 + (l % KS)] = v;
            }
            __syncwarp();
#pragma unroll
            for (int ry = 0; ry < kSY + 2 * R; ++ry) {
                const int gy = y0 - R + ry;
                float4 u[kSX + 2 * R];
                const bool row_ok = (unsigned)gy < (unsigned)p.h;
#pragma unroll
                for (int j = 0; j < kSX + 2 * R; ++j) {
                    const int gx = x0 - R + j;
                    u[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row_ok && (unsigned)gx < (unsigned)p.w)
                        u[j] = __ldg(reinterpret_cast<const float4*>(p.u + (((size_t)n * p.h + gy) * p.w + gx) * CU + (size_t)a * CO) + lane);
                }
#pragma unroll
                for (int py = 0; py < kSY; ++py) {
                    const int dy = ry - py;
                    if (dy < 0 || dy >= KS) continue;
#pragma unroll
                    for (int px = 0; px < kSX; ++px) {
                        const float4 w0 = *reinterpret_cast<const float4*>(&s_at[warp][py * kSX + px][dy * 8]);
                        const float w4 = s_at[warp][py * kSX + px][dy * 8 + 4];
                        const float wv[5] = {w0.x, w0.y, w0.z, w0.w, w4};
                        float4& o = acc[py * kSX + px];
#pragma unroll
                        for (int dx = 0; dx < KS; ++dx) {
                            o.x = fmaf(wv[dx], u[px + dx].x, o.x);
                            o.y = fmaf(wv[dx], u[px + dx].y, o.y);
                            o.z = fmaf(wv[dx], u[px + dx].z, o.z);
                            o.w = fmaf(wv[dx], u[px + dx].w, o.w);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int py = 0; py < kSY; ++py)
#pragma unroll
            for (int px = 0; px < kSX; ++px) {
                const int gy = y0 + py, gx = x0 + px;
                if (gy >= p.h || gx >= p.w) continue;
                const float4 o = acc[py * kSX + px];
                const float4 r = make_float4(fmaxf(o.x + bias.x, 0.f), fmaxf(o.y + bias.y, 0.f), fmaxf(o.z + bias.z, 0.f), fmaxf(o.w + bias.w, 0.f));
                reinterpret_cast<float4*>(p.y + (((size_t)n * p.h + gy) * p.w + gx) * CO)[lane] = r;
            }
    }
}

int launch_hyper(int which, const HyperParams& p, cudaStream_t st) {
    if (which == 0) {
        EVK_REQUIRE(p.H % 4 == 0 && p.W % 4 == 0 && p.bins + 1 <= 8, EVK_ERR_ARG, "hyper context: H, W must be multiples of 4 and bins <= 7");
        const int64_t total = (int64_t)p.N * (p.H / 4) * (p.W / 4) * 8;
        hyper_context_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1184), 256, 0, st>>>(p);
    } else if (which == 1) {
        const int64_t total = (int64_t)p.N * p.h * p.w * p.A * p.L;
        hyper_atoms_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, sizeof(float) * p.K * p.L, st>>>(p);
    } else if (which == 3) {
        EVK_REQUIRE(p.A == 6 && p.ks == 5 && p.K == 12 && p.CO == 128, EVK_ERR_ARG,
                    "hyper apply (re-associated): only num_atoms=6, kernel_size=5, 12 bases, 128 output channels is built (got A=%d ks=%d K=%d CO=%d)",
                    p.A, p.ks, p.K, p.CO);
        const int strips_x = ceil_div(p.w, kSX), strips_y = ceil_div(p.h, kSY), n_strips = strips_x * strips_y * p.N;
        const int blocks = std::min(ceil_div(n_strips, kSWarps), kNumSMs);
        const size_t smem = sizeof(float) * (((12 * 25 + 3) & ~3) + (size_t)kSWarps * kSX * kSY * (5 * 8 + 6 * 12));
        static bool attr_set[64] = {false};
        int dev = 0;
        EVK_CHECK_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            EVK_CHECK_CUDA(cudaFuncSetAttribute(hyper_apply_u_kernel<6, 5, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        hyper_apply_u_kernel<6, 5, 12><<<blocks, kSWarps * 32, smem, st>>>(p, strips_x, strips_y, n_strips);
    } else {
        EVK_REQUIRE(p.A == 6 && p.ks == 5 && p.C % kACH == 0, EVK_ERR_ARG,
                    "hyper apply: only num_atoms=6, kernel_size=5, C%%32==0 is built (got A=%d ks=%d C=%d)", p.A, p.ks, p.C);
        dim3 grid(ceil_div(p.w, kATW), ceil_div(p.h, kATH), p.N * (p.C / kACH));
        hyper_apply_kernel<6, 5><<<grid, 256, 0, st>>>(p);
    }
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
