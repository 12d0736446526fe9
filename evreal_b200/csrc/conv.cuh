// Convolution problem descriptor shared by the SIMT (fp32 CUDA-core) and the
// tcgen05 (split-bf16 tensor-core) implicit-GEMM kernels.
//
// All activations are NHWC.  The GEMM view is
//   M = N*Hout*Wout output pixels, Ncol = cout (packed), K = kh*kw*(c1+c2)
// with k = (r*kw + s)*(c1+c2) + c and torch.cat((x1, x2), 1) expressed as two
// channel ranges so it is never materialised.
#pragma once
#include <cuda_bf16.h>
#include <vector>

#include "evk_common.cuh"

namespace evk {

enum Epilogue : int {
    EPI_LINEAR = 0,   // y = act(acc + bias [+ res])                       ConvLayer / ResidualBlock / decoder
    EPI_LSTM = 1,     // packed cout = 4*C (ch*4 + {in,remember,out,cell}) ConvLSTM (model/submodules.py:231-243)
    EPI_GRU_UR = 2,   // packed cout = 2*C (ch*2 + {update,reset})         ConvGRU  (model/submodules.py:281-282)
    EPI_GRU_OUT = 3,  // cout = C: o = tanh(.), h' = h(1-u) + o*u          ConvGRU  (model/submodules.py:283-285)
};

struct ConvParams {
    // inputs
    const float* x1 = nullptr; int c1 = 0;
    const float* x2 = nullptr; int c2 = 0;
    int N = 0, Hin = 0, Win = 0, Hout = 0, Wout = 0;
    int kh = 0, kw = 0, stride = 1, pad = 0;
    // weights: SIMT layout [K][cout] fp32 (cout contiguous); bias [cout] (BatchNorm folded)
    const float* w = nullptr;
    const float* bias = nullptr;
    int cout = 0;
    int epi = EPI_LINEAR;
    int act = ACT_NONE;
    const float* res = nullptr;     // EPI_LINEAR: residual added before the activation, NHWC [.,cout]
    float* y = nullptr;             // EPI_LINEAR output NHWC [.,cout]
    // recurrent epilogues (C = hidden channels)
    const float* c_prev = nullptr; float* c_new = nullptr;   // LSTM cell (may alias: pointwise)
    float* h_new = nullptr;                                   // LSTM / GRU_OUT new hidden state
    const float* h_prev = nullptr;                            // GRU_UR / GRU_OUT previous hidden state
    float* u_out = nullptr; float* hr_out = nullptr;          // GRU_UR outputs: update gate, h*reset
    const float* u_in = nullptr;                              // GRU_OUT input: update gate
    // ---- tensor-core path (conv_tc.cu).  "Split" tensors are two bf16 planes [2][N,H,W,C]: hi = bf16(v),
    // lo = bf16(v - hi); the three products hi*hi + lo*hi + hi*lo accumulate in fp32 in tensor memory.
    const __nv_bfloat16* x1s = nullptr;   // split companions of x1 / x2 (required by the TC path)
    const __nv_bfloat16* x2s = nullptr;
    __nv_bfloat16* ys = nullptr;          // optional split copy of y (EPI_LINEAR) for a tensor-core consumer
    __nv_bfloat16* hs_new = nullptr;      // optional split copy of h_new (EPI_LSTM, EPI_GRU_OUT)
    __nv_bfloat16* hrs_out = nullptr;     // optional split copy of hr_out (EPI_GRU_UR)
    const __nv_bfloat16* w_tc = nullptr;  // weights [2][cout_pad][K] bf16 (hi, lo), K-major
    int cout_pad = 0;                     // rows of w_tc (cout rounded up to a multiple of 16)
    // "row-window" input (the Cin <= 8 head convolution on tensor cores): x1s is the packed tensor
    // [2][N][Hin][Win + 8][8] bf16 (pixel x at padded column x + kw/2, channels >= cin and the pad columns zero) and the
    // GEMM sees, for every pixel, the 64 contiguous values of the 8 pixels starting there as its "channels" (a TMA
    // tensor map whose pixel stride, 16 B, is smaller than its 128 B row): the kw taps of one kernel row are ONE K
    // chunk, so the layer runs as a (kh x 1) convolution with c1 = 64.  kw_packed = the real kw (<= 8).
    int kw_packed = 0;
    // kw_group = G > 1: one GEMM row computes G consecutive output pixels (G + kw - 1 <= 8: their taps all lie in the same
    // 8-pixel window) as G * cout packed columns -- the output [N,H,W,cout] is the same memory as [N,H,W/G,G*cout], so
    // the layer is an ordinary one on a W/G-wide image (Win = Wout = W/G, cout = G * real cout, bias repeated G times)
    // whose window rows are G pixels apart.  MMAs G times wider, G times fewer tiles.
    int kw_group = 0;
    // "window" mode for 16-channel layers (FireNet): the same row-window K layout over ordinary activations.  x1s / x2s are
    // ROW-PADDED split tensors [2][N][Hin][win_wp][win_c] (pixel x at padded column x + kw_packed/2, pad columns zero), 64 / win_c
    // pixels make one 128-byte K row (a 32-byte K row -- BK = 16, 32-byte swizzle -- costs ~3x per MMA), and kw_group
    // output pixels share a window (kw_group + kw_packed - 1 <= 64 / win_c).  c1 (and c2) = 64; weights from pack_weights_window.
    int win_c = 0, win_wp = 0;
    // horizontal stride / padding when they differ from the vertical ones (stride, pad).  Used by the "pixel pair" form of
    // a stride-2 layer with 32 input channels (first encoder): [N,H,W,32] is the same memory as [N,H,W/2,64], and output x
    // of a 5-tap stride-2 row reads exactly the pixel pairs x-1, x, x+1 -- a kh x 3 convolution with 64 channels, horizontal
    // stride 1 and padding 1 (weights from pack_weights_pixel_pair), whose K rows are 128 bytes instead of 64.
    int stride_x = 0, pad_x = -1;
    // split OUTPUTS (ys / hs_new / hrs_out) written row-padded for window-mode consumers: padded width, channels per pixel, left pad
    int s_wp = 0, s_c = 0, s_left = 0;
    // "row pair" form of a stride-1 layer with cout == 32 (any such layer that is not a phase-stacked decoder, e.g. the last
    // decoder with EVK_NO_POLY=1 or a TransposedConvLayer decoder): the GEMM computes output rows 2y and 2y+1
    // together as N = 64 columns of a (kh+1) x kw convolution with vertical stride 2 (weights of row 2y+1 shifted one tap
    // down) -- an MMA with N <= 64 costs the same ~50 cycles as one with N = 32, so this halves the MMA count per
    // output pixel at 6/5 of the taps.  w_tc then holds the stacked weights [2][64][(kh+1)*kw*cin] (pack_weights_row_pair).
    int row_pair = 0;
    // "phase-stacked" form of UpsampleConvLayer (poly.cu): the GEMM runs on the replicate-padded LOW-resolution map
    // (x1s = [2][N][Hin][Win][c1], Hin = H + 4, pad = 0, Hout x Wout = H x W) with N = 4 * cout composite columns
    // (column block a*2+b = output phase); the epilogue writes column block (a, b) of GEMM row (i, j) to output pixel
    // (2i+a, 2j+b) of the [N, 2*Hout, 2*Wout, cout] result and adds the border corrections (minus the excess of the taps
    // that fall outside the upsampled map) on the two outermost rows / columns before the activation.
    // w_tc = pack_weights_phase4 -> pack_weights_tc, cout_pad = 4*cout.  2 = one row phase per N tile (bn = 2*cout), the
    // all-zero tap row of each skipped.
    int phase4 = 0;
    const float* ring_h = nullptr;        // [2 (top, bottom)][N][2*Wout][4*cout]: column block l = output row {0, 1, Ho-2, Ho-1}
    const float* ring_v = nullptr;        // [2 (left, right)][N][2*Hout][4*cout]: column block l = output column {0, 1, Wo-2, Wo-1}
    // prediction layer fused into the epilogue (EPI_LINEAR, cout <= 32, one N tile): out[pix] = act(sum_c w[c] *
    // (y[pix,c] + skip[pix,c]) + b) -- model/unet.py:136-138; y itself need not be stored
    const float* pred_w = nullptr; const float* pred_skip = nullptr; float* pred_out = nullptr;
    float pred_bias = 0.f; int pred_sigmoid = 0;
    // skip operand of the fused prediction layer as split-bf16 planes (hi + lo = the value to 2^-17): when set it replaces
    // pred_skip, and the producer of the skip tensor (the head) need not write its fp32 copy at all
    const __nv_bfloat16* pred_skip_s = nullptr; long long pred_skip_plane = 0;
    // "mixed" operand decomposition (conv_tc.cu, MIXED kernels): x.w ~ x16.w16 + x8.wl8 + xl8.w8 -- one fp16 product and two fp8
    // products at twice the rate (2 tensor-pipe units instead of the 3 of bf16x3; measured 1.5-2.5e-5 end to end on the shipped
    // E2VID weights against 5e-6, tools/mixed_numerics_probe.py).  mixed = 1: x1s / x2s are mixed-format companions (plane 0 =
    // fp16(v), plane 1 = per 64-channel chunk [e5m2(v) | e5m2(256 (v - x16))]) and the kernel reads w_mx / w_iscale.
    // Plain layers only (c1, c2 multiples of 64; no row-window / pixel-pair / row-pair forms).
    int mixed = 0;
    const __nv_bfloat16* w_mx = nullptr;  // [2][cout_pad][K] 2-byte slots: plane 0 fp16(w S[n]), plane 1 per 64-k chunk [e4m3(wl S[n]) | e4m3(w S[n] / 256)]
    const float* w_iscale = nullptr;      // [cout_pad] 1 / S[n] (S[n] a power of two)
    int ys_mixed = 0, hs_mixed = 0;       // the split copy this layer writes (ys / hs_new) is in the mixed format
    struct TcPlan* tc = nullptr;          // tensor maps + tiling, built once per layer (tc_plan_create)
};

// ---- tensor-core implicit GEMM (TMA + tcgen05 + TMEM), conv_tc.cu
bool tc_eligible(const ConvParams& p);
int tc_plan_create(ConvParams& p);          // fills p.tc; EVK_ERR_ARG when the shape does not qualify
void tc_plan_destroy(TcPlan* plan);
int launch_conv_tc(const ConvParams& p, cudaStream_t st);
// fp32 -> split bf16 planes (elementwise): dst[0..n) = hi, dst[n..2n) = lo
int launch_split(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t st);
// host: fp32 [K][cout] (K-major rows of the SIMT layout) -> bf16 [2][cout_pad][K]
void pack_weights_tc(const float* w_kc, int K, int cout, int cout_pad, std::vector<__nv_bfloat16>& out);

// host: fp32 [K][cout] -> mixed-format weights (ConvParams::w_mx, K % 64 == 0) and the inverse column scales
void pack_weights_mixed(const float* w_kc, int K, int cout, int cout_pad, std::vector<__nv_bfloat16>& out, std::vector<float>& iscale);
bool tc_mixed_capable(const ConvParams& p);      // layer shape can run as a MIXED kernel
// fp32 NHWC [pixels][C] (C % 64 == 0) -> mixed-format companion
int launch_split_mixed(const float* src, __nv_bfloat16* dst, int64_t pixels, int C, cudaStream_t st);
// host: fp32 [kh*kw*cin][cout] (kw = 5, stride 2) -> pixel-pair weights [kh*3*(2*cin)][cout] (see ConvParams::stride_x)
void pack_weights_pixel_pair(const float* w_kc, int kh, int kw, int cin, int cout, std::vector<float>& out);
// host: fp32 [kh*kw*cin][cout] -> row-pair weights [ (kh+1)*kw*cin ][2*cout] (see ConvParams::row_pair)
void pack_weights_row_pair(const float* w_kc, int kh, int kw, int cin, int cout, std::vector<float>& out);

// ---- phase-stacked decoder (poly.cu)
// host: fp32 [25*cin][cout] -> composite phase weights [25*cin][4*cout] (column = (a*2+b)*cout + n)
void pack_weights_phase4(const float* w_kc, int cin, int cout, std::vector<float>& out);
// host: NEGATED pre-summed out-of-bounds tap weights of the two border-line convolutions, [2][5*cin][4*cout]
void pack_weights_ring(const float* w_kc, int cin, int cout, std::vector<float>& out);
// out[2][N][H+4][W+4][C] = split_bf16(x + skip), replicate padding of 2
int launch_add_pad_split(const float* x, const float* skip, __nv_bfloat16* out, int N, int H, int W, int C, cudaStream_t st, int mixed = 0);
// u_ext just outside the four borders as split-bf16 line images [2][2N][2W+4][C] (horizontal) / [2][2N][2H+4][C] (vertical)
int launch_ring_lines(const __nv_bfloat16* xp, __nv_bfloat16* lines_h, __nv_bfloat16* lines_v, int N, int H, int W, int C, cudaStream_t st, int mixed = 0);

int launch_conv_simt(const ConvParams& p, cudaStream_t st);
// dispatcher: tensor-core split-bf16 kernel when the shape qualifies and precision == 0, else fp32 SIMT
int launch_conv(const ConvParams& p, int precision, cudaStream_t st);

// head ConvLayer on the NCHW event tensor (Cin = num_bins): NCHW in -> NHWC out, ReLU
int launch_head_conv(const float* x_nchw, const float* w /*[k*k*cin][cout]*/, const float* bias, float* y_nhwc,
                     __nv_bfloat16* ys /*optional split copy*/, int N, int cin, int H, int W, int k, int cout, cudaStream_t st);
// prediction ConvLayer: 1x1 conv on (x [+ skip]) NHWC -> channel 0 only, NCHW [N,1,H,W]; optional sigmoid
int launch_pred(const float* x, const float* skip, const float* w /*[cin]*/, float bias, float* y, int64_t pixels,
                int cin, int sigmoid, cudaStream_t st);
// y[N,2H,2W,C] = bilinear_x2(x + skip), align_corners=False (model/unet.py:130-134 + submodules.py:88)
int launch_upsample2x_add(const float* x, const float* skip, float* y, __nv_bfloat16* ys, int N, int H, int W, int C, cudaStream_t st);
// y[N,2H,2W,C]: (x + skip) at the even positions, zero elsewhere (ConvTranspose2d stride 2 as a stride-1 convolution)
int launch_zero_insert2x_add(const float* x, const float* skip, float* y, __nv_bfloat16* ys, int N, int H, int W, int C, cudaStream_t st);
// NCHW fp32 [N,cin,H,W] (cin <= 8) -> packed split-bf16 row-window tensor [2][N][H][W+8][8], pixel x at column x + left
// (src_H x src_W = size of the source planes, sampled every `stride` pixels: nearest-neighbour reduction by an integer factor)
int launch_head_pack(const float* x_nchw, __nv_bfloat16* packed, int N, int cin, int H, int W, int left, cudaStream_t st, int src_H = 0,
                     int src_W = 0, int stride = 1, int src_planes = 0);

// ---- ET-Net token path (etnet.cu)
int launch_layernorm256(const float* x, const float* gamma, const float* beta, float* out, __nv_bfloat16* out_s, int64_t T, cudaStream_t st);
int launch_attention(const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* out, __nv_bfloat16* out_s, int N, int Lq,
                     int Lk, cudaStream_t st);
int launch_add_pos(const float* x, const float* pos, float* out, int N, int64_t per_sample, cudaStream_t st);
int launch_avg6(const float* const* six, float* out, int64_t n, cudaStream_t st);
void sine_position_table(int n, int d, std::vector<float>& out);

// ---- SPADE-E2VID glue (spade.cu)
int launch_add_split(const float* x, const float* s, float* out, __nv_bfloat16* out_s, int64_t n, cudaStream_t st);
int launch_spade_shuffle(const float* c0, const float* gb, const float* alpha, const float* shift, float* out, __nv_bfloat16* out_s, int N,
                         int h, int w, int C, cudaStream_t st);
int launch_spade_pred(const float* x, const float* head, const float* w, const float* bias_host3, float* prev, float* image, int N, int64_t HW,
                      int C, cudaStream_t st);
int launch_spade_first_frame(float* in, float* prev, int N, int bins, int64_t HW, cudaStream_t st);
// host: [kh*kw*cin][cout] (cin = T tensors of c_tensor channels) -> window K layout [kh*T*64][group*cout]
// (k = r*64*T + t*64 + slot*c_tensor + c; output pixel g of a group reads tap q from window slot g + q)
void pack_weights_window(const float* w_kc, int kh, int kw, int cin, int c_tensor, int cout, int group, std::vector<float>& out);
// fp32 [rows][W][C] -> split planes in the row-padded layout [2][rows][wp][C] (pixel x at column x + left; pads untouched)
int launch_split_padded(const float* src, __nv_bfloat16* dst, int64_t rows, int W, int C, int wp, int left, cudaStream_t st);
// host: head weights [kh*kw*cin][cout] (SIMT layout) -> row-window K layout [kh*64][group*cout] (k = r*64 + slot*8 + c,
// output pixel g of a group reads tap q from window slot g + q)
void pack_head_weights_rowwin(const float* w_kc, int kh, int kw, int cin, int cout, int group, std::vector<float>& out);

}  // namespace evk
