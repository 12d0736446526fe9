// Polyphase ("phase-stacked") form of UpsampleConvLayer (model/submodules.py:69-97 behind model/unet.py:130-134):
//     y = ReLU(conv5x5_pad2(bilinear_x2(x + skip)) + b)
//
// bilinear x2 (align_corners=False) makes every upsampled sample a fixed 2-tap mix of low-resolution samples
// (0.25 / 0.75), so output pixel (2i+a, 2j+b) is a 5x5 convolution of the LOW-resolution map around (i, j) with
// weights that depend only on the phase (a, b):
//     Wc[a,b][ty][tx] = sum_{dy,dx} w[dy][dx] * cy[a][dy][ty] * cx[b][dx][tx]         (float64, once per model)
// The four phases are stacked along the GEMM's N dimension (N = 4 * cout), and the epilogue of the tensor-core kernel
// scatters column block `phase` of pixel (i, j) to output pixel (2i+a, 2j+b) ("pixel shuffle").  For the last decoder
// of the E2VID family (cout = 32) this turns a GEMM with N = 32 (MMAs at a quarter of the tensor pipe's width) on the
// full-resolution map into one with N = 128 on the quarter-resolution map, and the upsampled tensor is never written.
//
// Borders.  The reference zero-pads the UPSAMPLED map, while its bilinear kernel clamps at the border.  The uniform
// polyphase formula on a replicate-padded low-resolution map (add_pad_split_kernel writes it: 2 pixels of padding) gives
// exactly the clamped bilinear values inside the map, but "sees" u_ext = u[clamp] instead of 0 outside it.  The excess
//     corr[Y,X] = sum_{(dy,dx): (Y+dy, X+dx) outside} w[dy][dx] * u_ext[Y+dy][X+dx]
// touches only the two outermost output rows / columns.  Both u_ext rows (columns) outside a border are equal, so the
// excess of a border line is a 1x5 convolution along that line with pre-summed tap weights: two small launches of the
// same tensor-core kernel (ring_lines_kernel builds their inputs), whose results the decoder's epilogue adds before
// the activation.
#include <algorithm>
#include <cmath>
#include <vector>

#include "conv.cuh"
#include "tc.cuh"

namespace evk {

// c[d + 2][t]: coefficient of low-resolution sample (i + t - 2) in upsampled sample (2i + a + d), d in [-2, 2]
static void phase_coef(int a, double c[5][5]) {
    for (int d = 0; d < 5; ++d)
        for (int t = 0; t < 5; ++t) c[d][t] = 0.0;
    for (int d = -2; d <= 2; ++d) {
        const int s = a + d;
        const int m = (s + 4) / 2 - 2, r = (s + 4) & 1;      // floor division
        if (r == 0) { c[d + 2][m - 1 + 2] += 0.25; c[d + 2][m + 2] += 0.75; }
        else        { c[d + 2][m + 2] += 0.75; c[d + 2][m + 1 + 2] += 0.25; }
    }
}

void pack_weights_phase4(const float* w_kc, int cin, int cout, std::vector<float>& out) {
    out.assign((size_t)25 * cin * 4 * cout, 0.f);
    double cy[2][5][5], cx[2][5][5];
    for (int a = 0; a < 2; ++a) { phase_coef(a, cy[a]); phase_coef(a, cx[a]); }
    std::vector<double> acc((size_t)25);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int c = 0; c < cin; ++c)
                for (int n = 0; n < cout; ++n) {
                    std::fill(acc.begin(), acc.end(), 0.0);
                    for (int dy = 0; dy < 5; ++dy)
                        for (int dx = 0; dx < 5; ++dx) {
                            const double w = (double)w_kc[((size_t)(dy * 5 + dx) * cin + c) * cout + n];
                            for (int ty = 0; ty < 5; ++ty) {
                                if (cy[a][dy][ty] == 0.0) continue;
                                for (int tx = 0; tx < 5; ++tx) acc[ty * 5 + tx] += w * cy[a][dy][ty] * cx[b][dx][tx];
                            }
                        }
                    for (int t = 0; t < 25; ++t)
                        out[((size_t)t * cin + c) * (4 * cout) + (a * 2 + b) * cout + n] = (float)acc[t];
                }
}

// Ring correction weights as two 1x5 convolutions (horizontal / vertical border lines): [v][5*cin][4*cout], column block l
// = border line l.  Horizontal line l = output row {0, 1, Ho-2, Ho-1}: the taps whose ROW falls outside, summed per column
// offset s (both outside rows read the same u_ext row); vertical line l = output column {0, 1, Wo-2, Wo-1}: the taps whose
// COLUMN falls outside, per row offset.  Stored NEGATED: the decoder's epilogue adds the result.
void pack_weights_ring(const float* w_kc, int cin, int cout, std::vector<float>& out) {
    out.assign((size_t)2 * 5 * cin * 4 * cout, 0.f);
    static const int dsets[4][2] = {{-2, -1}, {-2, -2}, {2, 2}, {1, 2}};
    for (int v = 0; v < 2; ++v)
        for (int l = 0; l < 4; ++l)
            for (int s = 0; s < 5; ++s)
                for (int c = 0; c < cin; ++c)
                    for (int n = 0; n < cout; ++n) {
                        double acc = 0.0;
                        for (int j = 0; j < 2; ++j) {
                            if (j == 1 && dsets[l][1] == dsets[l][0]) break;
                            const int d = dsets[l][j] + 2;
                            const int tap = v ? s * 5 + d : d * 5 + s;
                            acc += (double)w_kc[((size_t)tap * cin + c) * cout + n];
                        }
                        out[(((size_t)v * 5 + s) * cin + c) * (4 * cout) + l * cout + n] = (float)-acc;
                    }
}

// ------------------------------------------------------------------ (x + skip) -> replicate-padded split-bf16 planes
__global__ void __launch_bounds__(256)
add_pad_split_kernel(const float* __restrict__ x, const float* __restrict__ skip, __nv_bfloat16* __restrict__ out, long long plane,
                     int N, int H, int W, int C8, int mixed) {
    const int Hp = H + 4, Wp = W + 4;
    const int64_t total = (int64_t)N * Hp * Wp * C8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const int px = (int)((i / C8) % Wp);
        const int py = (int)((i / ((int64_t)C8 * Wp)) % Hp);
        const int n = (int)(i / ((int64_t)C8 * Wp * Hp));
        const int sy = min(max(py - 2, 0), H - 1), sx = min(max(px - 2, 0), W - 1);
        const size_t src = (((size_t)n * H + sy) * W + sx) * (size_t)(C8 * 8) + (size_t)c8 * 8;
        float4 v0 = __ldg(reinterpret_cast<const float4*>(x + src));
        float4 v1 = __ldg(reinterpret_cast<const float4*>(x + src + 4));
        if (skip != nullptr) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(skip + src));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(skip + src + 4));
            v0.x += s0.x; v0.y += s0.y; v0.z += s0.z; v0.w += s0.w;
            v1.x += s1.x; v1.y += s1.y; v1.z += s1.z; v1.w += s1.w;
        }
        const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        const size_t o = i * 8;
        if (mixed) { store_mixed8(out, plane, o, c8 * 8, f); continue; }     // mixed-operand consumer (tc.cuh)
        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16(f[e], hi[e], lo[e]);
        *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out + plane + o) = *reinterpret_cast<const uint4*>(lo);
    }
}

int launch_add_pad_split(const float* x, const float* skip, __nv_bfloat16* out, int N, int H, int W, int C, cudaStream_t st, int mixed) {
    EVK_REQUIRE(x && out && C % 8 == 0 && (!mixed || C % 64 == 0), EVK_ERR_ARG, "add_pad_split: bad argument (C=%d)", C);
    const int64_t total = (int64_t)N * (H + 4) * (W + 4) * (C / 8);
    const long long plane = (long long)N * (H + 4) * (W + 4) * C;
    add_pad_split_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 4736), 256, 0, st>>>(x, skip, out, plane, N, H, W, C / 8, mixed);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ------------------------------------------------------------------ border ring correction: source lines
// The excess of a border pixel is a 1x5 convolution ALONG the border of one line of u_ext just outside the map (rows -1 /
// Ho for the horizontal borders, columns -1 / Wo for the vertical ones; pack_weights_ring).  This kernel writes those
// lines as split-bf16 "images" whose rows are independent lines -- [2][R = 2*N][L + 4][C], row = side * N + image, 2
// positions of padding -- so that the two corrections run on the tensor-core convolution kernel as ordinary 1x5 layers
// with 4*cout output columns (border line l in column block l).  Horizontal lines carry u_ext over [-2, Wo + 2); vertical
// lines are ZERO outside [0, Ho): the taps whose row is outside as well belong to the horizontal correction.
__global__ void __launch_bounds__(256)
ring_lines_kernel(const __nv_bfloat16* __restrict__ xp, long long xp_plane, __nv_bfloat16* __restrict__ lh, __nv_bfloat16* __restrict__ lv,
                  int N, int H, int W, int C8, int mixed) {
    const int Ho = 2 * H, Wo = 2 * W, Wp = W + 4, C = C8 * 8;
    const int64_t nh = (int64_t)2 * N * (Wo + 4) * C8, nv = (int64_t)2 * N * (Ho + 4) * C8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nh + nv; i += (int64_t)gridDim.x * blockDim.x) {
        const bool vert = i >= nh;
        const int64_t k = vert ? i - nh : i;
        const int Lp = (vert ? Ho : Wo) + 4;
        const int c8 = (int)(k % C8);
        const int q = (int)((k / C8) % Lp) - 2;
        const int row = (int)(k / ((int64_t)C8 * Lp));          // side * N + image
        const int side = row / N, n = row - side * N;
        const int fixed = side ? (vert ? Wo : Ho) : -1;
        const int Y = vert ? q : fixed, X = vert ? fixed : q;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!vert || (q >= 0 && q < Ho)) {
            const int my = (Y + 4) / 2 - 2, ry = (Y + 4) & 1, mx = (X + 4) / 2 - 2, rx = (X + 4) & 1;
            const int rowA = (ry ? my : my - 1) + 2, colA = (rx ? mx : mx - 1) + 2;
            const float wyA = ry ? 0.75f : 0.25f, wxA = rx ? 0.75f : 0.25f;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const float wgt = (dy ? 1.0f - wyA : wyA) * (dx ? 1.0f - wxA : wxA);
                    const size_t o = (((size_t)n * (H + 4) + rowA + dy) * Wp + colA + dx) * (size_t)C + (size_t)c8 * 8;
                    if (mixed) {
                        float e8[8];
                        load_mixed8(xp, xp_plane, o, c8 * 8, e8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaf(wgt, e8[e], v[e]);
                        continue;
                    }
                    const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(xp + o));
                    const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(xp + xp_plane + o));
                    const __nv_bfloat16* hb = reinterpret_cast<const __nv_bfloat16*>(&h4);
                    const __nv_bfloat16* lb = reinterpret_cast<const __nv_bfloat16*>(&l4);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(wgt, __bfloat162float(hb[e]) + __bfloat162float(lb[e]), v[e]);
                }
        }
        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) split_bf16(v[e], hi[e], lo[e]);
        __nv_bfloat16* dst = vert ? lv : lh;
        const int64_t plane = (vert ? nv : nh) * 8;
        *reinterpret_cast<uint4*>(dst + k * 8) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(dst + plane + k * 8) = *reinterpret_cast<const uint4*>(lo);
    }
}

int launch_ring_lines(const __nv_bfloat16* xp, __nv_bfloat16* lines_h, __nv_bfloat16* lines_v, int N, int H, int W, int C, cudaStream_t st, int mixed) {
    EVK_REQUIRE(xp && lines_h && lines_v && C % 8 == 0 && H >= 2 && W >= 2, EVK_ERR_ARG, "ring_lines: bad argument (C=%d)", C);
    const int64_t total = (int64_t)2 * N * (2 * W + 4 + 2 * H + 4) * (C / 8);
    ring_lines_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, 0, st>>>(
        xp, (long long)N * (H + 4) * (W + 4) * C, lines_h, lines_v, N, H, W, C / 8, mixed);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
