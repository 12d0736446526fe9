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

// (3b) the same computation with the U tile staged in shared memory.  CTA = 8 rows x 16 columns of output pixels, all 128
// output channels; a stage = (atom a, 64-channel half): the 12 x 20 pixel halo tile of that slice of U (61 kB) arrives by
// cp.async (zero fill outside the image) into one of two buffers while the previous stage computes.  Warp w owns tile rows
// 2w', columns 8w'' (2 x 8 pixels); a lane owns 2 channels of the half.  Per stage a warp pulls its 6 x 12 pixel strip of U
// into registers once (72 conflict-free 8-byte loads) and then walks its 16 pixels: 7 broadcast 16-byte loads of the pixel's
// 25 atoms, 50 multiply-adds out of registers.  U crosses L2 -> SM 1.9x (halo) instead of 3x, every load is overlapped, and
// the multiply-add pipe sees ~3 instructions per shared-memory wavefront.
constexpr int kTX = 16, kTY = 8, kHX = kTX + 4, kHY = kTY + 4, kHalf = 64;

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool ok) {
    const int n = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

template <int A, int KS, int K>
__global__ void __launch_bounds__(256, 1) hyper_apply_u2_kernel(const HyperParams p, int tiles_x, int tiles_y, int n_tiles) {
    constexpr int L = KS * KS, LP = 28, NC = A * K;                 // 25 atoms padded to 28 floats (7 x 16 bytes)
    extern __shared__ __align__(16) float s_dyn[];
    float* s_u = s_dyn;                                               // [2][kHY][kHX][kHalf]
    float* s_at = s_u + 2 * kHY * kHX * kHalf;                        // [128 px][LP]
    float* s_coef = s_at + kTX * kTY * LP;                            // [128 px][NC]
    float* s_bases = s_coef + kTX * kTY * NC;                         // [K][L]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < K * L; i += 256) s_bases[i] = p.bases[i];
    const int CO = p.CO, CU = A * CO;
    const int wy = (warp >> 1) * 2, wx = (warp & 1) * 8;              // the warp's 2 x 8 pixels inside the tile
    const uint32_t s_u_addr = (uint32_t)__cvta_generic_to_shared(s_u);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int x0 = tx * kTX, y0 = ty * kTY;
        __syncthreads();                                              // previous tile done with s_coef / s_at / s_u
        // basis coefficients of the tile (coalesced: 8 rows of up to 16 * NC contiguous floats)
        for (int i = tid; i < kTY * kTX * NC; i += 256) {
            const int row = i / (kTX * NC), rem = i - row * (kTX * NC);
            const int gy = min(y0 + row, p.h - 1), gx = x0 + rem / NC;
            s_coef[i] = gx < p.w ? __ldg(p.coef + (((size_t)n * p.h + gy) * p.w + x0) * NC + rem) : 0.f;
        }
        auto issue = [&](int stage) {                                 // stage -> (a, half); 12 x 20 pixels x 16 chunks of 16 bytes
            const int a = stage >> 1, hf = stage & 1;
            const uint32_t dst0 = s_u_addr + (uint32_t)(stage & 1) * (kHY * kHX * kHalf * 4);
            for (int i = tid; i < kHY * kHX * (kHalf / 4); i += 256) {
                const int c4 = i % (kHalf / 4), pix = i / (kHalf / 4);
                const int gy = y0 - 2 + pix / kHX, gx = x0 - 2 + pix % kHX;
                const bool ok = (unsigned)gy < (unsigned)p.h && (unsigned)gx < (unsigned)p.w;
                const float* src = p.u + (((size_t)n * p.h + (ok ? gy : 0)) * p.w + (ok ? gx : 0)) * CU + (size_t)a * CO + hf * kHalf + c4 * 4;
                cp_async16_zfill(dst0 + (uint32_t)(pix * kHalf + c4 * 4) * 4, src, ok);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        float2 acc[2][16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc[0][i] = make_float2(0.f, 0.f); acc[1][i] = make_float2(0.f, 0.f); }
        issue(0);
#pragma unroll 1
        for (int a = 0; a < A; ++a)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                              // (unrolled: acc[hf] stays in registers)
            const int stage = 2 * a + hf;
            __syncthreads();                                          // everyone done with the buffer the next stage overwrites, and with s_at
            if (stage + 1 < 2 * A) issue(stage + 1);
            if (hf == 0) {                                            // atoms of this tile for atom index a
                for (int i = tid; i < kTX * kTY * L; i += 256) {
                    const int pix = i / L, l = i - pix * L;
                    const float* c = s_coef + pix * NC + a * K;
                    float v = 0.f;
#pragma unroll
                    for (int k = 0; k < K; ++k) v = fmaf(c[k], s_bases[k * L + l], v);
                    s_at[pix * LP + l] = v;
                }
            }
            if (stage + 1 < 2 * A) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            float2 u[6][12];
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    // (asm volatile: the strip is loaded ONCE per stage and lives in registers; left to itself the compiler re-reads
                    // shared memory for every tap and the kernel becomes shared-memory bound)
                    const uint32_t addr = s_u_addr + (uint32_t)(((stage & 1) * (kHY * kHX) + (wy + r) * kHX + wx + c) * kHalf + 2 * lane) * 4u;
                    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(u[r][c].x), "=f"(u[r][c].y) : "r"(addr));
                }
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 8; ++px) {
                    const float4* at4 = reinterpret_cast<const float4*>(s_at + ((wy + py) * kTX + wx + px) * LP);
                    float w[LP];
#pragma unroll
                    for (int q = 0; q < LP / 4; ++q) { const float4 t = at4[q]; w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w; }
                    float2 o = acc[hf][py * 8 + px];
#pragma unroll
                    for (int dy = 0; dy < KS; ++dy)
#pragma unroll
                        for (int dx = 0; dx < KS; ++dx) {
                            o.x = fmaf(w[dy * KS + dx], u[py + dy][px + dx].x, o.x);
                            o.y = fmaf(w[dy * KS + dx], u[py + dy][px + dx].y, o.y);
                        }
                    acc[hf][py * 8 + px] = o;
                }
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const float2 bias = __ldg(reinterpret_cast<const float2*>(p.out_bias + hf * kHalf) + lane);
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
                for (int px = 0; px < 8; ++px) {
                    const int gy = y0 + wy + py, gx = x0 + wx + px;
                    if (gy >= p.h || gx >= p.w) continue;
                    const float2 o = acc[hf][py * 8 + px];
                    reinterpret_cast<float2*>(p.y + (((size_t)n * p.h + gy) * p.w + gx) * CO + hf * kHalf)[lane] =
                        make_float2(fmaxf(o.x + bias.x, 0.f), fmaxf(o.y + bias.y, 0.f));
                }
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
        static const bool v1 = getenv("EVK_HYPER_APPLY_V1") != nullptr;
        if (!v1) {
            const int tiles_x = ceil_div(p.w, kTX), tiles_y = ceil_div(p.h, kTY), n_tiles = tiles_x * tiles_y * p.N;
            const size_t smem2 = sizeof(float) * ((size_t)2 * kHY * kHX * kHalf + (size_t)kTX * kTY * 28 + (size_t)kTX * kTY * 72 + 12 * 25);
            static bool attr2[64] = {false};
            int dev2 = 0;
            EVK_CHECK_CUDA(cudaGetDevice(&dev2));
            if (dev2 < 0 || dev2 >= 64 || !attr2[dev2]) {
                EVK_CHECK_CUDA(cudaFuncSetAttribute(hyper_apply_u2_kernel<6, 5, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                if (dev2 >= 0 && dev2 < 64) attr2[dev2] = true;
            }
            hyper_apply_u2_kernel<6, 5, 12><<<std::min(n_tiles, kNumSMs), 256, smem2, st>>>(p, tiles_x, tiles_y, n_tiles);
            EVK_CHECK_CUDA(cudaGetLastError());
            return EVK_OK;
        }
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
