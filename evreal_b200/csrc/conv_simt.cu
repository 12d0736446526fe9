// fp32 CUDA-core implicit-GEMM convolution with fused recurrent epilogues.
//
// This is the exact-fp32 path: it runs the layers whose channel counts are too
// small for the tensor-core tiles (head Cin=5, FireNet's 16-channel stack, the
// 1x1 prediction layer) and serves as the on-device cross-check for the
// tcgen05 split-bf16 kernels.  Reference semantics: model/submodules.py:8-35
// (ConvLayer), :152-184 (ResidualBlock), :187-245 (ConvLSTM), :248-287 (ConvGRU).
#include "conv.cuh"
#include "tc.cuh"

namespace evk {

constexpr int kBK = 16;

template <int BM, int BN>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvParams p) {
    constexpr int TX = BN / 4;
    constexpr int A_PER_THREAD = BM * 4 / 256;   // float4 loads of the A tile per thread
    constexpr int B_F4 = kBK * BN / 4;           // float4 count of the B tile
    static_assert((BM / 4) * (BN / 4) == 256, "tile must map to 256 threads of 4x4 micro-tiles");
    __shared__ __align__(16) float As[2][kBK][BM];
    __shared__ __align__(16) float Bs[2][kBK][BN];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int M = p.N * p.Hout * p.Wout;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int Cin = p.c1 + p.c2;
    const int chunks = (Cin + kBK - 1) / kBK;
    const int iters = p.kh * p.kw * chunks;

    // fixed per-thread gather coordinates for the A tile
    int a_img[A_PER_THREAD], a_iy[A_PER_THREAD], a_ix[A_PER_THREAD], a_kc[A_PER_THREAD], a_pix[A_PER_THREAD];
#pragma unroll
    for (int j = 0; j < A_PER_THREAD; ++j) {
        const int idx = tid + j * 256;
        a_pix[j] = idx % BM;
        a_kc[j] = idx / BM;
        const int m = m0 + a_pix[j];
        if (m < M) {
            const int ox = m % p.Wout;
            const int oy = (m / p.Wout) % p.Hout;
            a_img[j] = m / (p.Wout * p.Hout);
            a_iy[j] = oy * p.stride - p.pad;
            a_ix[j] = ox * p.stride - p.pad;
        } else {
            a_img[j] = -1;
            a_iy[j] = a_ix[j] = 0;
        }
    }

    float4 ra[A_PER_THREAD];
    float4 rb;
    auto gload = [&](int it) {
        const int tap = it / chunks;
        const int c0 = (it - tap * chunks) * kBK;
        const int r = tap / p.kw, s = tap - r * p.kw;
#pragma unroll
        for (int j = 0; j < A_PER_THREAD; ++j) {
            const int c = c0 + a_kc[j] * 4;
            const int iy = a_iy[j] + r, ix = a_ix[j] + s;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a_img[j] >= 0 && (unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win && c < Cin) {
                const size_t pix = ((size_t)a_img[j] * p.Hin + iy) * p.Win + ix;
                const float* src = (c < p.c1) ? (p.x1 + pix * p.c1 + c) : (p.x2 + pix * p.c2 + (c - p.c1));
                v = __ldg(reinterpret_cast<const float4*>(src));
            }
            ra[j] = v;
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < B_F4) {
            const int kr = tid / TX, nc = tid % TX;
            const int c = c0 + kr, n = n0 + nc * 4;
            if (c < Cin && n < p.cout)
                rb = __ldg(reinterpret_cast<const float4*>(p.w + ((size_t)tap * Cin + c) * p.cout + n));
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int j = 0; j < A_PER_THREAD; ++j) {
            const int k = a_kc[j] * 4;
            As[buf][k + 0][a_pix[j]] = ra[j].x;
            As[buf][k + 1][a_pix[j]] = ra[j].y;
            As[buf][k + 2][a_pix[j]] = ra[j].z;
            As[buf][k + 3][a_pix[j]] = ra[j].w;
        }
        if (tid < B_F4) *reinterpret_cast<float4*>(&Bs[buf][tid / TX][(tid % TX) * 4]) = rb;
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int it = 0; it < iters; ++it) {
        if (it + 1 < iters) gload(it + 1);
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < iters) sstore(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ------------------------------------------------------------ epilogue
    const int n = n0 + tx * 4;
    if (n >= p.cout) return;
    const float4 bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const float bb[4] = {bias4.x, bias4.y, bias4.z, bias4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bb[j];
        if (p.epi == EPI_LINEAR) {
            if (p.res != nullptr) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.res + (size_t)m * p.cout + n));
                v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], p.act);
            if (p.y != nullptr) *reinterpret_cast<float4*>(p.y + (size_t)m * p.cout + n) = make_float4(v[0], v[1], v[2], v[3]);
            if (p.ys != nullptr) {
                __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
                *reinterpret_cast<uint2*>(p.ys + (size_t)m * p.cout + n) = *reinterpret_cast<uint2*>(hi);
                *reinterpret_cast<uint2*>(p.ys + (size_t)M * p.cout + (size_t)m * p.cout + n) = *reinterpret_cast<uint2*>(lo);
            }
        } else if (p.epi == EPI_LSTM) {
            const int C = p.cout >> 2, ch = n >> 2;
            const size_t o = (size_t)m * C + ch;
            const float ig = sigmoidf_(v[0]), fg = sigmoidf_(v[1]), og = sigmoidf_(v[2]), cg = tanhf(v[3]);
            const float cell = __fadd_rn(__fmul_rn(fg, p.c_prev[o]), __fmul_rn(ig, cg));
            p.c_new[o] = cell;
            const float hid = og * tanhf(cell);
            p.h_new[o] = hid;
            if (p.hs_new != nullptr) {
                __nv_bfloat16 hi, lo;
                split_bf16(hid, hi, lo);
                p.hs_new[o] = hi;
                p.hs_new[(size_t)M * C + o] = lo;
            }
        } else if (p.epi == EPI_GRU_UR) {
            const int C = p.cout >> 1, ch = n >> 1;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const size_t o = (size_t)m * C + ch + q;
                const float u = sigmoidf_(v[2 * q]), r = sigmoidf_(v[2 * q + 1]);
                p.u_out[o] = u;
                p.hr_out[o] = p.h_prev[o] * r;
            }
        } else {   // EPI_GRU_OUT
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const size_t o = (size_t)m * p.cout + n + j;
                const float u = p.u_in[o];
                const float cand = tanhf(v[j]);
                p.h_new[o] = __fadd_rn(__fmul_rn(p.h_prev[o], __fsub_rn(1.0f, u)), __fmul_rn(cand, u));
            }
        }
    }
}

int launch_conv_simt(const ConvParams& p, cudaStream_t st) {
    EVK_REQUIRE(p.c1 % 4 == 0 && p.c2 % 4 == 0 && p.cout % 4 == 0, EVK_ERR_ARG,
                "conv_simt: channel counts must be multiples of 4 (c1=%d c2=%d cout=%d)", p.c1, p.c2, p.cout);
    EVK_REQUIRE(p.x1 && p.w && p.bias && (p.c2 == 0 || p.x2), EVK_ERR_ARG, "conv_simt: null tensor");
    const int M = p.N * p.Hout * p.Wout;
    if (p.cout <= 16) {
        dim3 grid(ceil_div(M, 256), ceil_div(p.cout, 16));
        conv_simt_kernel<256, 16><<<grid, 256, 0, st>>>(p);
    } else if (p.cout <= 32) {
        dim3 grid(ceil_div(M, 128), ceil_div(p.cout, 32));
        conv_simt_kernel<128, 32><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid(ceil_div(M, 64), ceil_div(p.cout, 64));
        conv_simt_kernel<64, 64><<<grid, 256, 0, st>>>(p);
    }
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ---------------------------------------------------------------------------
// head: NCHW event tensor (Cin = num_bins, not a multiple of 4) -> NHWC, ReLU.
// One thread per output pixel, 16x16 pixel tile with halo and all weights in
// shared memory.  model/unet.py:77-82 (5x5) and model/legacy.py:52-58 (3x3).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ y, __nv_bfloat16* __restrict__ ys, size_t plane, int cin, int H, int W, int k, int cout) {
    extern __shared__ __align__(16) float smem[];
    const int tw = 16 + k - 1;
    float* wsm = smem;                                  // [k*k*cin][cout]
    float* tile = smem + (size_t)k * k * cin * cout;    // [cin][tw][tw+1]
    const int n = blockIdx.z;
    const int y0 = blockIdx.y * 16, x0 = blockIdx.x * 16;
    const int tid = threadIdx.x, pad = k / 2;
    for (int i = tid; i < k * k * cin * cout; i += 256) wsm[i] = w[i];
    for (int i = tid; i < cin * tw * tw; i += 256) {
        const int c = i / (tw * tw), rem = i % (tw * tw);
        const int r = rem / tw, s = rem % tw;
        const int gy = y0 + r - pad, gx = x0 + s - pad;
        float v = 0.f;
        if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) v = x[(((size_t)n * cin + c) * H + gy) * W + gx];
        tile[(c * tw + r) * (tw + 1) + s] = v;
    }
    __syncthreads();
    const int ty = tid / 16, tx = tid % 16;
    const int oy = y0 + ty, ox = x0 + tx;
    if (oy >= H || ox >= W) return;
    float* out = y + (((size_t)n * H + oy) * W + ox) * cout;
    for (int co0 = 0; co0 < cout; co0 += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = (co0 + j < cout) ? bias[co0 + j] : 0.f;
        for (int r = 0; r < k; ++r)
            for (int s = 0; s < k; ++s)
                for (int c = 0; c < cin; ++c) {
                    const float v = tile[(c * tw + ty + r) * (tw + 1) + tx + s];
                    const float* wr = wsm + ((size_t)(r * k + s) * cin + c) * cout + co0;
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        if (co0 + j4 * 4 < cout) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wr + j4 * 4);
                            acc[j4 * 4 + 0] = fmaf(v, w4.x, acc[j4 * 4 + 0]);
                            acc[j4 * 4 + 1] = fmaf(v, w4.y, acc[j4 * 4 + 1]);
                            acc[j4 * 4 + 2] = fmaf(v, w4.z, acc[j4 * 4 + 2]);
                            acc[j4 * 4 + 3] = fmaf(v, w4.w, acc[j4 * 4 + 3]);
                        }
                    }
                }
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4)
            if (co0 + j4 * 4 < cout) {
                const float4 o4 = make_float4(fmaxf(acc[j4 * 4 + 0], 0.f), fmaxf(acc[j4 * 4 + 1], 0.f),
                                              fmaxf(acc[j4 * 4 + 2], 0.f), fmaxf(acc[j4 * 4 + 3], 0.f));
                *reinterpret_cast<float4*>(out + co0 + j4 * 4) = o4;
                if (ys != nullptr) {
                    const float f[4] = {o4.x, o4.y, o4.z, o4.w};
                    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_bf16(f[e], hi[e], lo[e]);
                    const size_t o = (size_t)(out - y) + co0 + j4 * 4;
                    *reinterpret_cast<uint2*>(ys + o) = *reinterpret_cast<uint2*>(hi);
                    *reinterpret_cast<uint2*>(ys + plane + o) = *reinterpret_cast<uint2*>(lo);
                }
            }
    }
}

int launch_head_conv(const float* x, const float* w, const float* bias, float* y, __nv_bfloat16* ys, int N, int cin, int H,
                     int W, int k, int cout, cudaStream_t st) {
    EVK_REQUIRE(cout % 4 == 0 && (k == 3 || k == 5 || k == 1 || k == 7), EVK_ERR_ARG, "head_conv: unsupported cout=%d k=%d", cout, k);
    const int tw = 16 + k - 1;
    const size_t smem = sizeof(float) * ((size_t)k * k * cin * cout + (size_t)cin * tw * (tw + 1));
    EVK_REQUIRE(smem <= 200 * 1024, EVK_ERR_ARG, "head_conv: weights do not fit in shared memory (%zu B)", smem);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        EVK_CHECK_CUDA(cudaFuncSetAttribute(head_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(ceil_div(W, 16), ceil_div(H, 16), N);
    head_conv_kernel<<<grid, 256, smem, st>>>(x, w, bias, y, ys, (size_t)N * H * W * cout, cin, H, W, k, cout);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ---------------------------------------------------------------------------
// prediction layer: 1x1 conv of (x + skip) down to the image channel, folded
// BN, optional sigmoid.  model/unet.py:136-138, eval.py:143.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pred_kernel(const float* __restrict__ x, const float* __restrict__ skip, const float* __restrict__ w, float bias,
            float* __restrict__ y, int64_t pixels, int cin, int sigmoid) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
        const float4* xp = reinterpret_cast<const float4*>(x + i * cin);
        const float4* sp = skip ? reinterpret_cast<const float4*>(skip + i * cin) : nullptr;
        const float4* wp = reinterpret_cast<const float4*>(w);
        float acc = 0.f;
        for (int c = 0; c < cin / 4; ++c) {
            float4 a = __ldg(xp + c);
            if (sp) {
                const float4 b = __ldg(sp + c);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            const float4 ww = __ldg(wp + c);
            acc = fmaf(a.x, ww.x, acc);
            acc = fmaf(a.y, ww.y, acc);
            acc = fmaf(a.z, ww.z, acc);
            acc = fmaf(a.w, ww.w, acc);
        }
        acc += bias;
        y[i] = sigmoid ? sigmoidf_(acc) : acc;
    }
}

int launch_pred(const float* x, const float* skip, const float* w, float bias, float* y, int64_t pixels, int cin,
                int sigmoid, cudaStream_t st) {
    EVK_REQUIRE(cin % 4 == 0 && pixels > 0, EVK_ERR_ARG, "pred: cin=%d must be a multiple of 4", cin);
    pred_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(pixels, 256), 2368), 256, 0, st>>>(x, skip, w, bias, y, pixels, cin, sigmoid);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ---------------------------------------------------------------------------
// y = bilinear_x2(x + skip): ATen upsample_bilinear2d, align_corners=False:
// src = 0.5*(dst+0.5)-0.5 clamped at 0; i1 = min(i0+1, size-1); the horizontal
// lerp is applied inside the vertical one like ATen's CPU kernel.
// ---------------------------------------------------------------------------
// One thread = one INPUT pixel x 8 channels -> the 2x2 output block it maps to: the 3x3 input neighbourhood (x + skip)
// is read once (9 loads instead of 16 for four independent outputs), every output evaluates ATen's expression
// hy*(hx*v00 + lx*v01) + ly*(hx*v10 + lx*v11) with ATen's source indices / weights (clamped at the borders; a
// zero-weight tap may come from a clamped duplicate, 0 * finite = 0 either way).
__global__ void __launch_bounds__(256)
upsample2x_add_kernel(const float* __restrict__ x, const float* __restrict__ skip, float* __restrict__ y,
                      __nv_bfloat16* __restrict__ ys, int N, int H, int W, int C8) {
    const int Ho = 2 * H, Wo = 2 * W, C = C8 * 8;
    const int64_t total = (int64_t)N * H * W * C8;
    const int64_t out_elems = (int64_t)N * Ho * Wo * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const int ix = (int)((i / C8) % W);
        const int iy = (int)((i / ((int64_t)C8 * W)) % H);
        const int n = (int)(i / ((int64_t)C8 * W * H));
        const int ys3[3] = {max(iy - 1, 0), iy, min(iy + 1, H - 1)};
        const int xs3[3] = {max(ix - 1, 0), ix, min(ix + 1, W - 1)};
        float v[3][3][8];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const size_t o = (((size_t)n * H + ys3[r]) * W + xs3[q]) * C + c;
                const float4 a0 = __ldg(reinterpret_cast<const float4*>(x + o));
                const float4 a1 = __ldg(reinterpret_cast<const float4*>(x + o + 4));
                float t[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                if (skip) {
                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(skip + o));
                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(skip + o + 4));
                    t[0] += b0.x; t[1] += b0.y; t[2] += b0.z; t[3] += b0.w;
                    t[4] += b1.x; t[5] += b1.y; t[6] += b1.z; t[7] += b1.w;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[r][q][e] = t[e];
            }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            // output row 2*iy + a: source sy = iy - 0.25 (a = 0; clamped to 0 on the first row) or iy + 0.25 (a = 1)
            const int r0 = a == 0 ? 0 : 1;                       // index into ys3 of the upper source row
            const float ly = a == 0 ? (iy == 0 ? 0.f : 0.75f) : 0.25f;
            const float hy = 1.f - ly;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int q0 = b == 0 ? 0 : 1;
                const float lx = b == 0 ? (ix == 0 ? 0.f : 0.75f) : 0.25f;
                const float hx = 1.f - lx;
                // (first output row / column: ATen's source index is 0 with weight 1, i.e. v[1][.]; ys3[0] == ys3[1] == 0 there)
                float o8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    o8[e] = hy * (hx * v[r0][q0][e] + lx * v[r0][q0 + 1][e]) + ly * (hx * v[r0 + 1][q0][e] + lx * v[r0 + 1][q0 + 1][e]);
                const size_t oo = (((size_t)n * Ho + 2 * iy + a) * Wo + 2 * ix + b) * C + c;
                if (y != nullptr) {
                    *reinterpret_cast<float4*>(y + oo) = make_float4(o8[0], o8[1], o8[2], o8[3]);
                    *reinterpret_cast<float4*>(y + oo + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
                }
                if (ys != nullptr) {
                    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) split_bf16(o8[e], hi[e], lo[e]);
                    *reinterpret_cast<uint4*>(ys + oo) = *reinterpret_cast<const uint4*>(hi);
                    *reinterpret_cast<uint4*>(ys + out_elems + oo) = *reinterpret_cast<const uint4*>(lo);
                }
            }
        }
    }
}

// y[N,2H,2W,C]: y[2i,2j] = x[i,j] + skip[i,j], zero elsewhere -- the zero-insertion that turns ConvTranspose2d(k, stride 2,
// padding p, output_padding 1) into a stride-1 convolution with the flipped kernel and padding k-1-p (the extra zero
// row / column at the bottom / right is the output_padding).  model/submodules.py:38-66, model/unet.py:130-134.
__global__ void __launch_bounds__(256)
zero_insert2x_add_kernel(const float* __restrict__ x, const float* __restrict__ skip, float* __restrict__ y,
                         __nv_bfloat16* __restrict__ ys, int N, int H, int W, int C4) {
    const int Ho = 2 * H, Wo = 2 * W;
    const int64_t total = (int64_t)N * Ho * Wo * C4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int ox = (int)((i / C4) % Wo);
        const int oy = (int)((i / ((int64_t)C4 * Wo)) % Ho);
        const int n = (int)(i / ((int64_t)C4 * Wo * Ho));
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (((ox | oy) & 1) == 0) {
            const size_t o = (((size_t)n * H + (oy >> 1)) * W + (ox >> 1)) * C4 + c;
            v = __ldg(reinterpret_cast<const float4*>(x) + o);
            if (skip) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(skip) + o);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
        }
        if (y != nullptr) reinterpret_cast<float4*>(y)[i] = v;
        if (ys != nullptr) {
            const float f[4] = {v.x, v.y, v.z, v.w};
            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_bf16(f[e], hi[e], lo[e]);
            *reinterpret_cast<uint2*>(ys + i * 4) = *reinterpret_cast<uint2*>(hi);
            *reinterpret_cast<uint2*>(ys + total * 4 + i * 4) = *reinterpret_cast<uint2*>(lo);
        }
    }
}

int launch_zero_insert2x_add(const float* x, const float* skip, float* y, __nv_bfloat16* ys, int N, int H, int W, int C, cudaStream_t st) {
    EVK_REQUIRE(C % 4 == 0 && (y != nullptr || ys != nullptr), EVK_ERR_ARG, "zero_insert2x_add: bad argument (C=%d)", C);
    const int64_t total = (int64_t)N * 4 * H * W * (C / 4);
    zero_insert2x_add_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 4736), 256, 0, st>>>(x, skip, y, ys, N, H, W, C / 4);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

int launch_upsample2x_add(const float* x, const float* skip, float* y, __nv_bfloat16* ys, int N, int H, int W, int C,
                          cudaStream_t st) {
    EVK_REQUIRE(C % 8 == 0, EVK_ERR_ARG, "upsample2x_add: C=%d must be a multiple of 8", C);
    EVK_REQUIRE(y != nullptr || ys != nullptr, EVK_ERR_ARG, "upsample2x_add: no output");
    const int64_t total = (int64_t)N * H * W * (C / 8);
    upsample2x_add_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 4736), 256, 0, st>>>(x, skip, y, ys, N, H, W, C / 8);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
