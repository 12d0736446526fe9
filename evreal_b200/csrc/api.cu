// C ABI entry points (include/evreal_b200.h) for the stateless stages, error
// reporting, and the conv dispatcher.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "conv.cuh"

namespace evk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int voxelize_f32(const float*, const float*, const float*, const float*, int64_t, int, int, int, float*, int*, cudaStream_t);
int voxelize_raw(const int16_t*, const double*, const uint8_t*, int64_t, int, int, int, float*, int*, cudaStream_t);
int normalize_pad(const float*, float*, int, int, int, int, int, int, int, cudaStream_t);
int crop(const float*, float*, int, int, int, int, int, int, cudaStream_t);
int u8_to_f32(const uint8_t*, float*, int64_t, cudaStream_t);
int quantize_u8(const float*, uint8_t*, int64_t, cudaStream_t);
int searchsorted_f64(const double*, int64_t, const double*, int64_t, int, long long*, cudaStream_t);
int equalize_hist(const float*, float*, int, int, int, cudaStream_t);
int equalize_local(const float*, float*, int, int, int, int, int, cudaStream_t);
int voxelize_raw_batch(const evk_event_window*, int, int, int, int, float*, int*, cudaStream_t);
int u8_to_f32_batch(const uint8_t* const*, int, int64_t, float*, cudaStream_t);
int mse_ssim(const float*, const float*, int, int, int, int, double*, cudaStream_t);
int percentile_normalize(const float*, float*, int, int, double, double, int, cudaStream_t);

int launch_conv(const ConvParams& p, int precision, cudaStream_t st) {
    if (precision == 0 && p.tc != nullptr) return launch_conv_tc(p, st);
    return launch_conv_simt(p, st);
}

}  // namespace evk

extern "C" {

int evk_version(void) { return 100; }
const char* evk_last_error(void) { return evk::g_err; }

int evk_voxelize(const float* x, const float* y, const float* t, const float* p, int64_t n, int num_bins, int H, int W,
                 float* grid, int* oob_count, void* stream) {
    EVK_REQUIRE(x && y && t && p && grid, EVK_ERR_ARG, "evk_voxelize: null pointer");
    return evk::voxelize_f32(x, y, t, p, n, num_bins, H, W, grid, oob_count, (cudaStream_t)stream);
}

int evk_voxelize_raw(const int16_t* xy, const double* t, const uint8_t* pol, int64_t n, int num_bins, int H, int W,
                     float* grid, int* oob_count, void* stream) {
    EVK_REQUIRE(xy && t && pol && grid, EVK_ERR_ARG, "evk_voxelize_raw: null pointer");
    return evk::voxelize_raw(xy, t, pol, n, num_bins, H, W, grid, oob_count, (cudaStream_t)stream);
}

int evk_voxelize_raw_batch(const evk_event_window* windows, int n_windows, int num_bins, int H, int W, float* grids,
                           int* oob_count, void* stream) {
    return evk::voxelize_raw_batch(windows, n_windows, num_bins, H, W, grids, oob_count, (cudaStream_t)stream);
}

int evk_u8_to_f32_batch(const uint8_t* const* frames, int n_frames, int64_t numel, float* out, void* stream) {
    return evk::u8_to_f32_batch(frames, n_frames, numel, out, (cudaStream_t)stream);
}

int evk_stage_windows_h2d(const evk_event_window* host_windows, int n_windows, int16_t* st_xy, double* st_t, uint8_t* st_pol,
                          int64_t stride_events, void* stream) {
    EVK_REQUIRE(host_windows && st_xy && st_t && st_pol && n_windows > 0 && stride_events > 0, EVK_ERR_ARG,
                "evk_stage_windows_h2d: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int b = 0; b < n_windows; ++b) {
        const evk_event_window& w = host_windows[b];
        if (w.n <= 0) continue;
        EVK_REQUIRE(w.n <= stride_events && w.xy && w.t && w.pol, EVK_ERR_ARG, "evk_stage_windows_h2d: window %d (%lld events) does not fit", b, (long long)w.n);
        EVK_CHECK_CUDA(cudaMemcpyAsync(st_xy + (size_t)b * stride_events * 2, w.xy, (size_t)w.n * 4, cudaMemcpyHostToDevice, st));
        EVK_CHECK_CUDA(cudaMemcpyAsync(st_t + (size_t)b * stride_events, w.t, (size_t)w.n * 8, cudaMemcpyHostToDevice, st));
        EVK_CHECK_CUDA(cudaMemcpyAsync(st_pol + (size_t)b * stride_events, w.pol, (size_t)w.n, cudaMemcpyHostToDevice, st));
    }
    return EVK_OK;
}

int evk_stage_frames_h2d(const uint8_t* const* host_frames, int n_frames, int64_t numel, uint8_t* dst, void* stream) {
    EVK_REQUIRE(host_frames && dst && n_frames > 0 && numel > 0, EVK_ERR_ARG, "evk_stage_frames_h2d: bad argument");
    for (int b = 0; b < n_frames; ++b) {
        EVK_REQUIRE(host_frames[b] != nullptr, EVK_ERR_ARG, "evk_stage_frames_h2d: frame %d is null", b);
        EVK_CHECK_CUDA(cudaMemcpyAsync(dst + (size_t)b * numel, host_frames[b], (size_t)numel, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    }
    return EVK_OK;
}

int evk_normalize_pad(const float* in, float* out, int n_samples, int C, int H, int W, int Hp, int Wp, int do_normalize,
                      void* stream) {
    EVK_REQUIRE(in && out, EVK_ERR_ARG, "evk_normalize_pad: null pointer");
    return evk::normalize_pad(in, out, n_samples, C, H, W, Hp, Wp, do_normalize, (cudaStream_t)stream);
}

int evk_crop(const float* in, float* out, int n, int C, int Hp, int Wp, int H, int W, void* stream) {
    EVK_REQUIRE(in && out, EVK_ERR_ARG, "evk_crop: null pointer");
    return evk::crop(in, out, n, C, Hp, Wp, H, W, (cudaStream_t)stream);
}

int evk_u8_to_f32(const uint8_t* in, float* out, int64_t numel, void* stream) {
    EVK_REQUIRE(in && out, EVK_ERR_ARG, "evk_u8_to_f32: null pointer");
    return evk::u8_to_f32(in, out, numel, (cudaStream_t)stream);
}

int evk_equalize_local(const float* img, float* out, int n_images, int H, int W, int radius, int clip, void* stream) {
    EVK_REQUIRE(img && out, EVK_ERR_ARG, "evk_equalize_local: null pointer");
    return evk::equalize_local(img, out, n_images, H, W, radius, clip, (cudaStream_t)stream);
}

int evk_equalize_hist(const float* img, float* out, int n_images, int numel, int clip, void* stream) {
    EVK_REQUIRE(img && out, EVK_ERR_ARG, "evk_equalize_hist: null pointer");
    return evk::equalize_hist(img, out, n_images, numel, clip, (cudaStream_t)stream);
}

int evk_quantize_u8(const float* in, uint8_t* out, int64_t numel, void* stream) {
    EVK_REQUIRE(in && out, EVK_ERR_ARG, "evk_quantize_u8: null pointer");
    return evk::quantize_u8(in, out, numel, (cudaStream_t)stream);
}

int evk_searchsorted_f64(const double* t, int64_t n, const double* values, int64_t m, int right, int64_t* out, void* stream) {
    EVK_REQUIRE((t || n == 0) && values && out, EVK_ERR_ARG, "evk_searchsorted_f64: null pointer");
    return evk::searchsorted_f64(t, n, values, m, right, reinterpret_cast<long long*>(out), (cudaStream_t)stream);
}

int evk_mse_ssim(const float* img, const float* ref, int n_images, int H, int W, int clip, double* scores, void* stream) {
    EVK_REQUIRE(img && ref && scores, EVK_ERR_ARG, "evk_mse_ssim: null pointer");
    return evk::mse_ssim(img, ref, n_images, H, W, clip, scores, (cudaStream_t)stream);
}

int evk_percentile_normalize(const float* img, float* out, int n_images, int numel, double q_min, double q_max,
                             int apply_exp, void* stream) {
    EVK_REQUIRE(img && out, EVK_ERR_ARG, "evk_percentile_normalize: null pointer");
    return evk::percentile_normalize(img, out, n_images, numel, q_min, q_max, apply_exp, (cudaStream_t)stream);
}

int evk_pack_layer_weights(int kind, const float* w_oihw_host, int Cout, int Cin, int kh, int kw, int group, float* out_host,
                           int64_t out_cap, int64_t* out_len) {
    using namespace evk;
    EVK_REQUIRE(w_oihw_host && out_host && out_len && Cout > 0 && Cin > 0 && kh > 0 && kw > 0, EVK_ERR_ARG, "evk_pack_layer_weights: bad argument");
    // torch [Cout][Cin][kh][kw] -> the GEMM layout of every packer's input: [(r*kw + s)*Cin + c][Cout]
    std::vector<float> w_kc((size_t)kh * kw * Cin * Cout), out;
    for (int n = 0; n < Cout; ++n)
        for (int c = 0; c < Cin; ++c)
            for (int r = 0; r < kh; ++r)
                for (int q = 0; q < kw; ++q)
                    w_kc[((size_t)(r * kw + q) * Cin + c) * Cout + n] = w_oihw_host[(((size_t)n * Cin + c) * kh + r) * kw + q];
    switch (kind) {
        case 0:
            EVK_REQUIRE(kh == 5 && kw == 5, EVK_ERR_ARG, "evk_pack_layer_weights: phase stacking is built for 5x5 kernels");
            pack_weights_phase4(w_kc.data(), Cin, Cout, out);
            break;
        case 1:
            EVK_REQUIRE(kh == 5 && kw == 5, EVK_ERR_ARG, "evk_pack_layer_weights: border lines are built for 5x5 kernels");
            pack_weights_ring(w_kc.data(), Cin, Cout, out);
            break;
        case 2:
            EVK_REQUIRE(kw == 5, EVK_ERR_ARG, "evk_pack_layer_weights: pixel pairs are built for 5-tap rows");
            pack_weights_pixel_pair(w_kc.data(), kh, kw, Cin, Cout, out);
            break;
        case 3:
            EVK_REQUIRE(Cin % 16 == 0 && group >= 1 && group + kw - 1 <= 4, EVK_ERR_ARG, "evk_pack_layer_weights: window mode needs 16-channel tensors and group + kw - 1 <= 4");
            pack_weights_window(w_kc.data(), kh, kw, Cin, 16, Cout, group, out);
            break;
        case 4: {
            // mixed operands (conv.cuh, ConvParams::mixed), DECODED: [3][K][Cout] = w16 / S, wl8 / S, 256 w8 / S
            const int K = kh * kw * Cin, cp = (Cout + 31) / 32 * 32;
            EVK_REQUIRE(K % 64 == 0, EVK_ERR_ARG, "evk_pack_layer_weights: mixed operands need kh*kw*Cin %% 64 == 0");
            std::vector<__nv_bfloat16> packed;
            std::vector<float> isc;
            pack_weights_mixed(w_kc.data(), K, Cout, cp, packed, isc);
            const uint16_t* p0 = reinterpret_cast<const uint16_t*>(packed.data());
            const uint8_t* p1 = reinterpret_cast<const uint8_t*>(packed.data() + (size_t)cp * K);
            out.assign((size_t)3 * K * Cout, 0.f);
            for (int n = 0; n < Cout; ++n)
                for (int k = 0; k < K; ++k) {
                    __half_raw hr; hr.x = p0[(size_t)n * K + k];
                    const uint8_t* row = p1 + ((size_t)n * K + (size_t)(k / 64) * 64) * 2;
                    const __half_raw l = __nv_cvt_fp8_to_halfraw(row[k % 64], __NV_E4M3), w8 = __nv_cvt_fp8_to_halfraw(row[64 + k % 64], __NV_E4M3);
                    out[((size_t)0 * K + k) * Cout + n] = __half2float(__half(hr)) * isc[n];
                    out[((size_t)1 * K + k) * Cout + n] = __half2float(__half(l)) * isc[n];
                    out[((size_t)2 * K + k) * Cout + n] = __half2float(__half(w8)) * 256.0f * isc[n];
                }
            break;
        }
        default:
            EVK_REQUIRE(false, EVK_ERR_ARG, "evk_pack_layer_weights: unknown kind %d", kind);
    }
    *out_len = (int64_t)out.size();
    EVK_REQUIRE((int64_t)out.size() <= out_cap, EVK_ERR_ARG, "evk_pack_layer_weights: output needs %lld elements", (long long)out.size());
    std::copy(out.begin(), out.end(), out_host);
    return EVK_OK;
}

int evk_conv2d_nhwc(const float* x, int N, int H, int W, int Cin, const float* w_oihw_host, const float* bias_host, int Cout,
                    int k, int stride, int pad, int act, const float* res, int precision, float* y, void* stream) {
    using namespace evk;
    EVK_REQUIRE(x && w_oihw_host && y, EVK_ERR_ARG, "evk_conv2d_nhwc: null pointer");
    EVK_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && k > 0 && stride > 0 && pad >= 0 && Cin % 4 == 0 && Cout % 4 == 0,
                EVK_ERR_ARG, "evk_conv2d_nhwc: bad shape (channels must be multiples of 4)");
    cudaStream_t st = (cudaStream_t)stream;
    const int K = k * k * Cin;
    std::vector<float> wk((size_t)K * Cout), b(Cout, 0.f);
    for (int n = 0; n < Cout; ++n) {
        if (bias_host) b[n] = bias_host[n];
        for (int c = 0; c < Cin; ++c)
            for (int r = 0; r < k; ++r)
                for (int q = 0; q < k; ++q)
                    wk[((size_t)(r * k + q) * Cin + c) * Cout + n] = w_oihw_host[(((size_t)n * Cin + c) * k + r) * k + q];
    }
    float *dw = nullptr, *db = nullptr;
    __nv_bfloat16 *dwt = nullptr, *dxs = nullptr;
    int rc = EVK_OK;
    auto cleanup = [&]() { cudaFree(dw); cudaFree(db); cudaFree(dwt); cudaFree(dxs); };
    EVK_CHECK_CUDA(cudaMalloc(&dw, wk.size() * 4));
    EVK_CHECK_CUDA(cudaMalloc(&db, b.size() * 4));
    EVK_CHECK_CUDA(cudaMemcpyAsync(dw, wk.data(), wk.size() * 4, cudaMemcpyHostToDevice, st));
    EVK_CHECK_CUDA(cudaMemcpyAsync(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice, st));
    ConvParams p;
    p.x1 = x; p.c1 = Cin; p.N = N; p.Hin = H; p.Win = W; p.kh = p.kw = k; p.stride = stride; p.pad = pad;
    p.Hout = (H + 2 * pad - k) / stride + 1; p.Wout = (W + 2 * pad - k) / stride + 1;
    p.w = dw; p.bias = db; p.cout = Cout; p.epi = EPI_LINEAR; p.act = act; p.res = res; p.y = y;
    if (precision == 0 && tc_eligible(p)) {
        std::vector<__nv_bfloat16> wt;
        if (Cout == 32 && stride == 1 && p.Hout % 2 == 0 && res == nullptr && getenv("EVK_NO_ROW_PAIR") == nullptr) {
            std::vector<float> w2;          // same rule as the network builder (model.cu): row-pair form of the last decoder
            pack_weights_row_pair(wk.data(), k, k, Cin, Cout, w2);
            p.row_pair = 1;
            p.cout_pad = 2 * Cout;
            pack_weights_tc(w2.data(), (k + 1) * k * Cin, 2 * Cout, p.cout_pad, wt);
        } else {
            p.cout_pad = (Cout + 15) / 16 * 16;
            pack_weights_tc(wk.data(), K, Cout, p.cout_pad, wt);
        }
        const int64_t nx = (int64_t)N * H * W * Cin;
        if (cudaMalloc(&dwt, wt.size() * 2) != cudaSuccess || cudaMalloc(&dxs, (size_t)nx * 4) != cudaSuccess) {
            cleanup();
            set_error("evk_conv2d_nhwc: out of device memory");
            return EVK_ERR_CUDA;
        }
        cudaMemcpyAsync(dwt, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice, st);
        rc = launch_split(x, dxs, nx, st);
        p.w_tc = dwt; p.x1s = dxs;
        if (rc == EVK_OK) rc = tc_plan_create(p);
        if (rc == EVK_OK) rc = launch_conv_tc(p, st);
    } else {
        rc = launch_conv_simt(p, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (p.tc) tc_plan_destroy(p.tc);
    cleanup();
    if (rc != EVK_OK) return rc;
    EVK_CHECK_CUDA(e);
    return EVK_OK;
}

}  // extern "C"
