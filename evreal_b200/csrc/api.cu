// C ABI entry points (include/evreal_b200.h) for the stateless stages, error
// reporting, and the conv dispatcher.
#include <cstdarg>

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
int mse_ssim(const float*, const float*, int, int, int, int, double*, cudaStream_t);
int percentile_normalize(const float*, float*, int, int, double, double, int, cudaStream_t);

int launch_conv(const ConvParams& p, int precision, cudaStream_t st) {
    (void)precision;
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

int evk_mse_ssim(const float* img, const float* ref, int n_images, int H, int W, int clip, double* scores, void* stream) {
    EVK_REQUIRE(img && ref && scores, EVK_ERR_ARG, "evk_mse_ssim: null pointer");
    return evk::mse_ssim(img, ref, n_images, H, W, clip, scores, (cudaStream_t)stream);
}

int evk_percentile_normalize(const float* img, float* out, int n_images, int numel, double q_min, double q_max,
                             int apply_exp, void* stream) {
    EVK_REQUIRE(img && out, EVK_ERR_ARG, "evk_percentile_normalize: null pointer");
    return evk::percentile_normalize(img, out, n_images, numel, q_min, q_max, apply_exp, (cudaStream_t)stream);
}

}  // extern "C"
