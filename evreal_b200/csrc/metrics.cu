// Stage 3 + post-processing: fused clip + MSE + Gaussian SSIM, and the
// percentile ("robust") normalisation that sits between the network and the
// metrics for E2VID.
//
// Reference semantics: utils/eval_metrics.py:77-97,253-255 (scikit-image
// mean_squared_error / structural_similarity, sigma 1.5, 11 taps, float32 maps,
// float64 mean over the interior crop), utils/eval_utils.py:15-35 and
// eval.py:380-395 (np.percentile linear interpolation).
//
// HBM traffic is 2*H*W*4 bytes per frame (345 kB at 240x180): these kernels
// are launch-latency bound; everything per frame is one pass.
#include "evk_common.cuh"

namespace evk {

// ----------------------------------------------------------------- SSIM/MSE
constexpr int kTapR = 5;                 // int(3.5*1.5 + 0.5)
constexpr int kTaps = 2 * kTapR + 1;
constexpr int kTH = 16, kTW = 32;        // owned pixels per CTA
constexpr int kIH = kTH + 2 * kTapR, kIW = kTW + 2 * kTapR;

// The taps travel as a kernel argument (a __constant__ symbol is per device and would need per-device uploads).
struct GaussTaps { double w[kTaps]; };

static GaussTaps make_taps() {
    GaussTaps t;
    double s = 0.0;
    const double sigma = 1.5;
    for (int i = -kTapR; i <= kTapR; ++i) { t.w[i + kTapR] = exp(-0.5 / (sigma * sigma) * (double)(i * i)); s += t.w[i + kTapR]; }
    for (int i = 0; i < kTaps; ++i) t.w[i] /= s;
    return t;
}

// scipy correlate1d, symmetric branch: centre tap first, then pairs from the
// outside in, float64 accumulate, float32 store.
__device__ __forceinline__ float sym_filter(const float* v, int stride, const double* c_taps) {
    double acc = (double)v[0] * c_taps[kTapR];
#pragma unroll
    for (int j = kTapR; j >= 1; --j) acc += ((double)v[-j * stride] + (double)v[j * stride]) * c_taps[kTapR - j];
    return (float)acc;
}

__global__ void __launch_bounds__(256)
mse_ssim_kernel(const float* __restrict__ img, const float* __restrict__ ref, int H, int W, int clip,
                double* __restrict__ sums /* [n][2] */, const __grid_constant__ GaussTaps taps) {
    const double* c_taps = taps.w;
    __shared__ float sx[kIH][kIW], sy[kIH][kIW];          // x = ref, y = img (skimage argument order)
    __shared__ float vert[5][kTH][kIW];
    __shared__ double red[2][8];
    const int n = blockIdx.z;
    const float* X = ref + (size_t)n * H * W;
    const float* Y = img + (size_t)n * H * W;
    const int y0 = blockIdx.y * kTH, x0 = blockIdx.x * kTW;
    const int tid = threadIdx.x;

    for (int i = tid; i < kIH * kIW; i += 256) {
        const int r = i / kIW, c = i % kIW;
        const int gy = y0 + r - kTapR, gx = x0 + c - kTapR;
        float a = 0.0f, b = 0.0f;
        if ((unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W) {
            a = X[(size_t)gy * W + gx];
            b = Y[(size_t)gy * W + gx];
            if (clip) { a = fminf(fmaxf(a, 0.0f), 1.0f); b = fminf(fmaxf(b, 0.0f), 1.0f); }
        }
        sx[r][c] = a;
        sy[r][c] = b;
    }
    __syncthreads();

    // MSE over owned pixels: float32 difference squared, float64 accumulation
    double mse = 0.0;
    for (int i = tid; i < kTH * kTW; i += 256) {
        const int r = i / kTW, c = i % kTW;
        if (y0 + r < H && x0 + c < W) {
            const float d = __fsub_rn(sx[r + kTapR][c + kTapR], sy[r + kTapR][c + kTapR]);
            mse += (double)__fmul_rn(d, d);
        }
    }

    // vertical pass (axis 0 first, like scipy.ndimage.gaussian_filter)
    for (int i = tid; i < kTH * kIW; i += 256) {
        const int r = i / kIW, c = i % kIW;
        float col[5][kTaps];
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
            const float a = sx[r + k][c], b = sy[r + k][c];
            col[0][k] = a;
            col[1][k] = b;
            col[2][k] = __fmul_rn(a, a);
            col[3][k] = __fmul_rn(b, b);
            col[4][k] = __fmul_rn(a, b);
        }
#pragma unroll
        for (int m = 0; m < 5; ++m) vert[m][r][c] = sym_filter(&col[m][kTapR], 1, c_taps);
    }
    __syncthreads();

    // horizontal pass + SSIM map on interior owned pixels
    const float C1 = (float)(0.01 * 0.01), C2 = (float)(0.03 * 0.03);
    double ssim = 0.0;
    for (int i = tid; i < kTH * kTW; i += 256) {
        const int r = i / kTW, c = i % kTW;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= kTapR && gy < H - kTapR && gx >= kTapR && gx < W - kTapR) {
            const float ux = sym_filter(&vert[0][r][c + kTapR], 1, c_taps);
            const float uy = sym_filter(&vert[1][r][c + kTapR], 1, c_taps);
            const float uxx = sym_filter(&vert[2][r][c + kTapR], 1, c_taps);
            const float uyy = sym_filter(&vert[3][r][c + kTapR], 1, c_taps);
            const float uxy = sym_filter(&vert[4][r][c + kTapR], 1, c_taps);
            const float vx = __fsub_rn(uxx, __fmul_rn(ux, ux));
            const float vy = __fsub_rn(uyy, __fmul_rn(uy, uy));
            const float vxy = __fsub_rn(uxy, __fmul_rn(ux, uy));
            const float A1 = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, ux), uy), C1);
            const float A2 = __fadd_rn(__fmul_rn(2.0f, vxy), C2);
            const float B1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), C1);
            const float B2 = __fadd_rn(__fadd_rn(vx, vy), C2);
            ssim += (double)__fdiv_rn(__fmul_rn(A1, A2), __fmul_rn(B1, B2));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mse += __shfl_xor_sync(0xffffffffu, mse, o);
        ssim += __shfl_xor_sync(0xffffffffu, ssim, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = mse; red[1][tid >> 5] = ssim; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) { mse += red[0][w]; ssim += red[1][w]; }
        atomicAdd(&sums[2 * n + 0], mse);
        atomicAdd(&sums[2 * n + 1], ssim);
    }
}

__global__ void mse_ssim_finalize_kernel(double* sums, int n, double inv_all, double inv_interior) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        sums[2 * i + 0] *= inv_all;
        sums[2 * i + 1] *= inv_interior;
    }
}

int mse_ssim(const float* img, const float* ref, int n, int H, int W, int clip, double* scores, cudaStream_t st) {
    EVK_REQUIRE(n > 0 && H >= kTaps && W >= kTaps, EVK_ERR_ARG,
                "evk_mse_ssim: images must be at least %dx%d (got n=%d %dx%d)", kTaps, kTaps, n, H, W);
    static const GaussTaps taps = make_taps();
    EVK_CHECK_CUDA(cudaMemsetAsync(scores, 0, sizeof(double) * 2 * n, st));
    dim3 grid(ceil_div(W, kTW), ceil_div(H, kTH), n);
    mse_ssim_kernel<<<grid, 256, 0, st>>>(img, ref, H, W, clip, scores, taps);
    EVK_CHECK_CUDA(cudaGetLastError());
    const double inv_all = 1.0 / ((double)H * W);
    const double inv_int = 1.0 / ((double)(H - 2 * kTapR) * (W - 2 * kTapR));
    mse_ssim_finalize_kernel<<<ceil_div(n, 128), 128, 0, st>>>(scores, n, inv_all, inv_int);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ---------------------------------------------------- percentile normalise
// One CTA per image: exact order statistics by 4-pass 8-bit radix select on
// order-preserving integer keys (warp-private histograms), for the two ranks
// floor(q*(n-1)) of q_min and q_max at once; a fifth scan finds the successors
// (rank+1).  Then out = (v - P_lo) / (P_hi - P_lo) in float32.
constexpr int kSelThreads = 1024;
constexpr int kSelWarps = kSelThreads / 32;

__device__ __forceinline__ unsigned int float_key(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ float load_val(const float* v, int i, int apply_exp) {
    const float a = v[i];
    return apply_exp ? expf(a) : a;
}

__global__ void __launch_bounds__(kSelThreads)
percentile_normalize_kernel(const float* __restrict__ img, float* __restrict__ out, int numel, double q_lo, double q_hi,
                            int apply_exp) {
    extern __shared__ unsigned int hist[];          // [2][kSelWarps][256]
    __shared__ unsigned int s_prefix[2], s_rank[2];
    __shared__ unsigned int s_le[2], s_next[2];
    __shared__ float s_p[2];
    const float* v = img + (size_t)blockIdx.x * numel;
    float* o = out + (size_t)blockIdx.x * numel;
    const int tid = threadIdx.x, warp = tid >> 5;

    // numpy method 'linear' (alpha = beta = 1): virtual index = n*q + (alpha + q*(1-alpha-beta)) - 1,
    // evaluated in float64 in numpy's own operation order; gamma = its fractional part
    const double qf[2] = {q_lo / 100.0, q_hi / 100.0};
    const double vi[2] = {(double)numel * qf[0] + (1.0 + qf[0] * (1.0 - 1.0 - 1.0)) - 1.0,
                          (double)numel * qf[1] + (1.0 + qf[1] * (1.0 - 1.0 - 1.0)) - 1.0};
    if (tid < 2) {
        double f = floor(vi[tid]);
        if (f < 0) f = 0;
        if (f > numel - 1) f = numel - 1;
        s_rank[tid] = (unsigned int)f;   // rank still to find inside the current prefix bucket
        s_prefix[tid] = 0;
    }
    __syncthreads();

    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < 2 * kSelWarps * 256; i += kSelThreads) hist[i] = 0;
        __syncthreads();
        const unsigned int pre0 = s_prefix[0], pre1 = s_prefix[1];
        const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        unsigned int* h0 = hist + (0 * kSelWarps + warp) * 256;
        unsigned int* h1 = hist + (1 * kSelWarps + warp) * 256;
        for (int i = tid; i < numel; i += kSelThreads) {
            const unsigned int k = float_key(load_val(v, i, apply_exp));
            const unsigned int d = (k >> shift) & 0xffu;
            if ((k & mask) == pre0) atomicAdd(&h0[d], 1u);
            if ((k & mask) == pre1) atomicAdd(&h1[d], 1u);
        }
        __syncthreads();
        // fold warp-private histograms: thread t (< 512) owns bin t%256 of target t/256
        if (tid < 512) {
            const int tgt = tid >> 8, bin = tid & 255;
            unsigned int c = 0;
            for (int w = 0; w < kSelWarps; ++w) c += hist[(tgt * kSelWarps + w) * 256 + bin];
            hist[(tgt * kSelWarps) * 256 + bin] = c;
        }
        __syncthreads();
        if (warp < 2) {
            // warp t finds the bin holding rank r of target t: lane l owns bins 8l..8l+7 (a serial scan of 256 bins by one
            // thread cost ~8k cycles per pass)
            const unsigned int* h = hist + (warp * kSelWarps) * 256;
            const int l = tid & 31;
            unsigned int c[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = h[8 * l + j]; sum += c[j]; }
            unsigned int incl = sum;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (l >= off) incl += t;
            }
            const unsigned int r = s_rank[warp];
            unsigned int acc = incl - sum;                         // elements in the bins before this lane's
            const bool mine = (acc <= r && r < incl) || (l == 31 && r >= incl);
            __syncwarp();
            if (mine) {
                int b = 0;
                for (; b < 8; ++b) {
                    if (acc + c[b] > r) break;
                    acc += c[b];
                }
                if (b > 7) { b = 7; }
                s_rank[warp] = r - acc;
                s_prefix[warp] |= ((unsigned int)(8 * l + b) << shift);
            }
        }
        __syncthreads();
    }
    // s_prefix[t] is now the key of the rank-th smallest value.  Successor scan.
    if (tid < 2) { s_le[tid] = 0; s_next[tid] = 0xffffffffu; }
    __syncthreads();
    {
        const unsigned int k0 = s_prefix[0], k1 = s_prefix[1];
        unsigned int le0 = 0, le1 = 0, nx0 = 0xffffffffu, nx1 = 0xffffffffu;
        for (int i = tid; i < numel; i += kSelThreads) {
            const unsigned int k = float_key(load_val(v, i, apply_exp));
            if (k <= k0) le0++; else nx0 = min(nx0, k);
            if (k <= k1) le1++; else nx1 = min(nx1, k);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            le0 += __shfl_xor_sync(0xffffffffu, le0, off);
            le1 += __shfl_xor_sync(0xffffffffu, le1, off);
            nx0 = min(nx0, __shfl_xor_sync(0xffffffffu, nx0, off));
            nx1 = min(nx1, __shfl_xor_sync(0xffffffffu, nx1, off));
        }
        if ((tid & 31) == 0) {
            atomicAdd(&s_le[0], le0);
            atomicAdd(&s_le[1], le1);
            atomicMin(&s_next[0], nx0);
            atomicMin(&s_next[1], nx1);
        }
    }
    __syncthreads();
    if (tid < 2) {
        double f = floor(vi[tid]);
        if (f < 0) f = 0;
        const unsigned int rank = (unsigned int)f;
        const double gamma = vi[tid] - f;
        const float a = key_float(s_prefix[tid]);
        // rank+1 is the same value when duplicates extend past it, else the next distinct key
        float b = a;
        if (rank + 1 < (unsigned int)numel) b = (s_le[tid] > rank + 1) ? a : key_float(s_next[tid]);
        // numpy _lerp: a + (b-a)*t, and b - (b-a)*(1-t) when t >= 0.5
        const double d = (double)__fsub_rn(b, a);
        double r = (double)a + d * gamma;
        if (gamma >= 0.5) r = (double)b - d * (1.0 - gamma);
        s_p[tid] = (float)r;
    }
    __syncthreads();
    const float lo = s_p[0];
    const float range = __fsub_rn(s_p[1], lo);
    for (int i = tid; i < numel; i += kSelThreads) o[i] = __fdiv_rn(__fsub_rn(load_val(v, i, apply_exp), lo), range);
}

int percentile_normalize(const float* img, float* out, int n, int numel, double q_lo, double q_hi, int apply_exp,
                         cudaStream_t st) {
    EVK_REQUIRE(n > 0 && numel > 0, EVK_ERR_ARG, "evk_percentile_normalize: empty input");
    EVK_REQUIRE(q_lo >= 0 && q_hi <= 100 && q_lo <= q_hi, EVK_ERR_ARG, "evk_percentile_normalize: bad percentiles");
    const size_t smem = sizeof(unsigned int) * 2 * kSelWarps * 256;   // 64 KB
    static bool attr_set[64] = {false};          // the attribute is per device
    int dev = 0;
    EVK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        EVK_CHECK_CUDA(cudaFuncSetAttribute(percentile_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    percentile_normalize_kernel<<<n, kSelThreads, smem, st>>>(img, out, numel, q_lo, q_hi, apply_exp);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// ------------------------------------------------------- histogram equalisation
// EvalMetricsTracker.histogram_equalization, hist_eq == 'global' (utils/eval_metrics.py:326-331):
// skimage.exposure.equalize_hist(img) -> img_as_float32.  scikit-image is a third-party dependency absent from the
// reference tree (parity unpinned); its published algorithm: hist = np.histogram(img, 256 bins over [min, max]),
// cdf = cumsum(hist) / numel, out = np.interp(img, bin_centers, cdf).  One CTA per image.
constexpr int kEqBins = 256;

__global__ void __launch_bounds__(1024)
equalize_hist_kernel(const float* __restrict__ img, float* __restrict__ out, int numel, int clip) {
    __shared__ float s_min[32], s_max[32];
    __shared__ unsigned int s_hist[kEqBins];
    __shared__ double s_cdf[kEqBins];
    __shared__ double s_lo, s_hi;
    const float* v = img + (size_t)blockIdx.x * numel;
    float* o = out + (size_t)blockIdx.x * numel;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    auto load = [&](int i) { const float a = v[i]; return clip ? fminf(fmaxf(a, 0.0f), 1.0f) : a; };
    float mn = INFINITY, mx = -INFINITY;
    for (int i = tid; i < numel; i += 1024) { const float a = load(i); mn = fminf(mn, a); mx = fmaxf(mx, a); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    if (lane == 0) { s_min[warp] = mn; s_max[warp] = mx; }
    if (tid < kEqBins) s_hist[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 32; ++w) { mn = fminf(mn, s_min[w]); mx = fmaxf(mx, s_max[w]); }
        double lo = (double)mn, hi = (double)mx;
        if (lo == hi) { lo -= 0.5; hi += 0.5; }          // np.histogram widens an empty range
        s_lo = lo; s_hi = hi;
    }
    __syncthreads();
    const double lo = s_lo, hi = s_hi, width = (hi - lo) / kEqBins, norm = (double)kEqBins / (hi - lo);
    auto edge = [&](int i) { return i == kEqBins ? hi : lo + (double)i * width; };          // np.linspace(lo, hi, 257)
    for (int i = tid; i < numel; i += 1024) {
        const double a = (double)load(i);
        int b = (int)((a - lo) * norm);
        b = min(max(b, 0), kEqBins - 1);
        if (a < edge(b)) --b;                                                        // numpy's edge corrections
        else if (b != kEqBins - 1 && a >= edge(b + 1)) ++b;
        atomicAdd(&s_hist[min(max(b, 0), kEqBins - 1)], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long c = 0;
        for (int b = 0; b < kEqBins; ++b) { c += s_hist[b]; s_cdf[b] = (double)c / (double)numel; }
    }
    __syncthreads();
    auto centre = [&](int j) { return (edge(j) + edge(j + 1)) * 0.5; };
    const double c0 = centre(0), c_last = centre(kEqBins - 1);
    for (int i = tid; i < numel; i += 1024) {
        const double a = (double)load(i);
        double r;
        // np.interp over the bin centres: clamped outside, linear inside
        if (a <= c0) r = s_cdf[0];
        else if (a >= c_last) r = s_cdf[kEqBins - 1];
        else {
            int j = min(max((int)floor((a - c0) / width), 0), kEqBins - 2);
            if (a < centre(j) && j > 0) --j;
            else if (a >= centre(j + 1) && j < kEqBins - 2) ++j;
            const double xj = centre(j), xj1 = centre(j + 1);
            r = (s_cdf[j + 1] - s_cdf[j]) / (xj1 - xj) * (a - xj) + s_cdf[j];
        }
        o[i] = (float)r;
    }
}

int equalize_hist(const float* img, float* out, int n, int numel, int clip, cudaStream_t st) {
    EVK_REQUIRE(n > 0 && numel > 0, EVK_ERR_ARG, "evk_equalize_hist: empty input");
    equalize_hist_kernel<<<n, 1024, 0, st>>>(img, out, numel, clip);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// EvalMetricsTracker.histogram_equalization, hist_eq == 'local' (utils/eval_metrics.py:332-339):
//     img_as_float32(skimage.filters.rank.equalize(img_as_ubyte(img), footprint=disk(55)))
// scikit-image is a third-party dependency absent from the reference tree (parity unpinned); its published algorithm
// (filters/rank/generic_cy.pyx, _kernel_equalize): for every pixel, over the footprint pixels that lie INSIDE the image
// (pop of them), out = uint8(255 * #{v <= g} / pop) with g the pixel's own grey level (double division, truncation);
// disk(r) = {dx^2 + dy^2 <= r^2}; img_as_ubyte = rint(v * 255) in float32, img_as_float32 = u8 * float32(1 / 255).
// CTA = 16 x 16 output pixels; the (16 + 2r)^2 neighbourhood sits in shared memory as uint16 (0xffff = outside the image).
constexpr int kEqTile = 16;

__global__ void __launch_bounds__(kEqTile * kEqTile)
equalize_local_kernel(const float* __restrict__ img, float* __restrict__ out, int H, int W, int r, int clip) {
    extern __shared__ unsigned short s_px[];
    const int span = kEqTile + 2 * r;
    int* s_hw = reinterpret_cast<int*>(s_px + ((size_t)span * span + 1) / 2 * 2);      // half width of the disk per row offset
    const float* v = img + (size_t)blockIdx.z * H * W;
    float* o = out + (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * kEqTile, y0 = blockIdx.y * kEqTile;
    const int tid = threadIdx.y * kEqTile + threadIdx.x;
    for (int i = tid; i < span * span; i += kEqTile * kEqTile) {
        const int yy = y0 - r + i / span, xx = x0 - r + i % span;
        unsigned short q = 0xffffu;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            float a = v[(size_t)yy * W + xx];
            if (clip) a = fminf(fmaxf(a, 0.0f), 1.0f);
            q = (unsigned short)fminf(fmaxf(rintf(a * 255.0f), 0.0f), 255.0f);
        }
        s_px[i] = q;
    }
    for (int d = tid; d <= 2 * r; d += kEqTile * kEqTile) {
        const int dy = d - r;
        int w = 0;
        while ((w + 1) * (w + 1) + dy * dy <= r * r) ++w;
        s_hw[d] = w;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const int cx = threadIdx.x + r, cy = threadIdx.y + r;
    const unsigned int g = s_px[cy * span + cx];
    unsigned int pop = 0, le = 0;
    for (int d = 0; d <= 2 * r; ++d) {
        const int w = s_hw[d];
        const unsigned short* row = s_px + (cy + d - r) * span + cx;
        for (int dx = -w; dx <= w; ++dx) {
            const unsigned int q = row[dx];
            pop += q != 0xffffu;
            le += q <= g;                 // (0xffff never is)
        }
    }
    const double e = pop ? (double)(255ull * le) / (double)pop : 0.0;
    o[(size_t)y * W + x] = (float)(unsigned int)e * (1.0f / 255.0f);
}

int equalize_local(const float* img, float* out, int n, int H, int W, int radius, int clip, cudaStream_t st) {
    EVK_REQUIRE(n > 0 && H > 0 && W > 0 && radius >= 1 && radius <= 100, EVK_ERR_ARG, "evk_equalize_local: bad argument (radius 1..100)");
    EVK_REQUIRE(img != out, EVK_ERR_ARG, "evk_equalize_local: in-place operation is not supported (every output reads a neighbourhood)");
    const int span = kEqTile + 2 * radius;
    const size_t smem = ((size_t)span * span + 1) / 2 * 2 * sizeof(unsigned short) + (size_t)(2 * radius + 1) * sizeof(int);
    static bool attr_dev[64] = {false};
    int dev = 0;
    EVK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_dev[dev]) {
        EVK_CHECK_CUDA(cudaFuncSetAttribute(equalize_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_dev[dev] = true;
    }
    EVK_REQUIRE(smem <= 200 * 1024, EVK_ERR_ARG, "evk_equalize_local: radius too large for shared memory");
    dim3 grid((unsigned)ceil_div(W, kEqTile), (unsigned)ceil_div(H, kEqTile), (unsigned)n);
    equalize_local_kernel<<<grid, dim3(kEqTile, kEqTile), smem, st>>>(img, out, H, W, radius, clip);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
