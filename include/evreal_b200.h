/*
 * evreal_b200 -- C ABI of the B200-native event->video hot path.
 *
 * Every entry point replaces one piece of EVREAL's per-frame loop
 * (reference paths are relative to the EVREAL tree).
 * Plain pointers and sizes only: no torch types, no C++ exceptions cross this
 * boundary.  All pointers are DEVICE pointers unless the name says "host".
 * `stream` is a cudaStream_t passed as void*.  Every function returns 0 on
 * success or a negative EVK_ERR_* code; evk_last_error() gives the message of
 * the last failure on the calling thread.
 *
 * There is no CPU implementation behind this ABI: if no sm_100 device is
 * present the calls fail with EVK_ERR_CUDA.
 */
#ifndef EVREAL_B200_H
#define EVREAL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVK_OK 0
#define EVK_ERR_ARG (-1)     /* bad argument (python adapter raises ValueError/AssertionError) */
#define EVK_ERR_CUDA (-2)    /* CUDA runtime error */
#define EVK_ERR_INDEX (-3)   /* event coordinate outside the sensor (reference: IndexError) */
#define EVK_ERR_STATE (-4)   /* call order violated (e.g. forward before finalize) */
#define EVK_ERR_KEY (-5)     /* missing / unexpected weight tensor (reference: load_state_dict error) */

int evk_version(void);
const char* evk_last_error(void);

/* ---------------------------------------------------------------- stage 1 --
 * events_to_voxel_torch(xs, ys, ts, ps, num_bins, device, sensor_size)
 *   reference: utils/event_utils.py:27-59 (+ events_to_image_torch :4-24)
 * x,y,t,p: [n] float32 (x,y integer valued, t seconds relative to the window's
 * first event, p in {-1,+1} or any weight).  grid: [num_bins,H,W] float32,
 * zeroed by the call.  Coordinates are truncated toward zero like .long();
 * coordinates in [-W,0) / [-H,0) wrap like python indexing; anything else is
 * skipped and ADDED to *oob_count (int32 device scalar owned and zeroed by the
 * caller, may be NULL; the python adapter turns a non-zero count into
 * IndexError like the reference).
 * n == 0 is rejected with EVK_ERR_ARG (the reference indexes ts[-1]).
 */
int evk_voxelize(const float* x, const float* y, const float* t, const float* p, int64_t n,
                 int num_bins, int H, int W, float* grid, int* oob_count, void* stream);

/* Raw on-disk window variant, fusing MemMapDataset.get_events / __getitem__
 * casts (dataset.py:222-228, :52-58): xy int16 [n,2] (x,y), t float64 [n]
 * absolute seconds, pol uint8 [n] in {0,1}.  ts = (t - t[0]) rounded to f32,
 * ps = pol*2-1. */
int evk_voxelize_raw(const int16_t* xy, const double* t, const uint8_t* pol, int64_t n,
                     int num_bins, int H, int W, float* grid, int* oob_count, void* stream);

/* The same for several windows in ONE launch (frame i of a lock-step batch of independent sequences: eval.py:197
 * resets state per sequence, so sequences are independent units).  windows: HOST array of device pointers / sizes;
 * grids: [n_windows, num_bins, H, W] float32, zeroed by the call.  A window with n == 0 yields the zeros grid, as
 * MemMapDataset.__getitem__ does for empty windows (dataset.py:59-71). */
typedef struct evk_event_window {
    const int16_t* xy;
    const double* t;
    const uint8_t* pol;
    int64_t n;
} evk_event_window;
int evk_voxelize_raw_batch(const evk_event_window* windows, int n_windows, int num_bins, int H, int W,
                           float* grids, int* oob_count, void* stream);

/* Host-buffer ingest (the DataLoader's role, eval.py:72 / dataset.py:222-228): copies the raw windows of a lock-step
 * batch from (pinned) HOST arrays into per-sequence device staging rows of stride_events events, asynchronously on
 * `stream`; window b lands at st_xy + b*stride_events*2, st_t + b*stride_events, st_pol + b*stride_events. */
int evk_stage_windows_h2d(const evk_event_window* host_windows, int n_windows, int16_t* st_xy, double* st_t,
                          uint8_t* st_pol, int64_t stride_events, void* stream);
/* n_frames uint8 reference frames (HOST array of host pointers, numel bytes each) -> dst [n_frames, numel] on the device */
int evk_stage_frames_h2d(const uint8_t* const* host_frames, int n_frames, int64_t numel, uint8_t* dst, void* stream);

/* normalize_event_tensor(event_tensor)   reference: eval.py:398-410
 * in/out: [n_samples, numel] float32 (may alias).  Statistics are per sample
 * over non-zero entries.  Optionally fuses CropParameters.pad
 * (utils/util.py:30-48): when Hp/Wp differ from H/W the output is
 * [n_samples, C, Hp, Wp] with the input centred per the reference's
 * ceil/floor rule and zero borders.  Pass do_normalize=0 for pad only. */
int evk_normalize_pad(const float* in, float* out, int n_samples, int C, int H, int W,
                      int Hp, int Wp, int do_normalize, void* stream);

/* ---------------------------------------------------------------- stage 2 --
 * Recurrent reconstruction networks.   reference: model/model.py:108-190,
 * model/unet.py:85-143, model/legacy.py:32-187, model/submodules.py,
 * model/hyper/hyper_dynamic.py; factory eval.py:124-158.
 */
typedef struct evk_model evk_model;

#define EVK_ARCH_UNET_RECURRENT 0   /* E2VIDRecurrent / FlowNet: E2VID, E2VID+, SSL-E2VID, HyperE2VID */
#define EVK_ARCH_FIRENET_LEGACY 1   /* FireNet_legacy  (pretrained/FireNet)  */
#define EVK_ARCH_FIRENET 2          /* FireNet         (pretrained/FireNet+) */
#define EVK_ARCH_SPADE_E2VID 3      /* Unet6           (pretrained/SPADE-E2VID; model/spade_e2v.py:113-179) */
#define EVK_ARCH_ETNET 4            /* EITR / mls_tpa  (pretrained/ET-Net; model/eitr/u_trans.py:13-123) */

typedef struct {
    int arch;                 /* EVK_ARCH_* */
    int num_bins;             /* 5 */
    int base_channels;        /* 32 (unet) / 16 (firenet) */
    int num_encoders;         /* 3 (unet); ignored for firenet */
    int num_residual_blocks;  /* 2 */
    int kernel_size;          /* 5 (unet) / 3 (firenet) */
    int num_output_channels;  /* 1, or 3 for FlowNet (image = channel 0) */
    int final_sigmoid;        /* eval.py:143 */
    int dynamic_decoder;      /* HyperE2VID (unet.py:59-64) */
    int batch;                /* independent sequences run in lock-step */
    int height, width;        /* padded input size (multiple of 2^num_encoders) */
    int precision;            /* 0 = default (tensor-core split-bf16 where applicable), 1 = force fp32 SIMT */
} evk_model_config;

int evk_model_create(const evk_model_config* cfg, evk_model** out);
/* One call per state_dict entry, reference names with the wrapper prefix
 * (unetrecurrent. / unetflow. / net.) stripped; host float32 data. */
int evk_model_load_tensor(evk_model* m, const char* name, const float* host_data,
                          const int64_t* shape, int ndim);
/* Folds eval-mode BatchNorm into the convolutions, repacks and uploads. */
int evk_model_finalize(evk_model* m, void* stream);
/* model.reset_states()  (model/model.py:128-130, legacy.py:182-183) */
int evk_model_reset_states(evk_model* m, void* stream);
/* model(voxel)['image']: voxel [batch,num_bins,height,width] -> image [batch,1,height,width] */
int evk_model_forward(evk_model* m, const float* voxel, float* image, void* stream);
/* Number of recurrent state tensors and their sizes; get/set copy NCHW float32
 * (model.states property, model/model.py:116-122). index: unet -> 2*i (h), 2*i+1 (c); firenet -> i. */
int evk_model_num_states(evk_model* m);
int evk_model_state_shape(evk_model* m, int index, int64_t shape_nchw[4]);
int evk_model_get_state(evk_model* m, int index, float* out_nchw, void* stream);
int evk_model_set_state(evk_model* m, int index, const float* in_nchw, void* stream);
int evk_model_destroy(evk_model* m);
/* Handle-owned input / output buffers of the NEXT evk_model_forward call: in [batch,num_bins,height,width], out
 * [batch,1,height,width].  They are double-buffered by forward parity (call k uses pair k & 1), so a caller may fill the
 * input of call k+1 and read the output of call k-1 on other streams while call k runs; passing the returned pointers to
 * evk_model_forward skips the device-to-device staging copies.  (HyperE2VID's previous reconstruction, model/model.py:139-143,
 * is the other output buffer: the caller must not overwrite `out` of call k before call k+1 has finished.) */
int evk_model_io_buffers(evk_model* m, float** in, float** out);
/* kernels launched by the last forward (for bench.py's gpu_launches) */
int evk_model_last_launch_count(evk_model* m);
/* algorithmic conv FLOPs of one forward (2*Cout*Cin*kh*kw*Hout*Wout summed) */
double evk_model_flops(evk_model* m);

/* Eager (graph-less) forward with a CUDA-event pair around every launch: per-launch milliseconds and
 * algorithmic FLOPs (host arrays of max_ops entries; *n_ops = launches of one forward).  Advances the
 * recurrent state exactly like evk_model_forward and synchronises the stream.  This is the live
 * measurement behind bench.py's roofline block; the reference's only counterpart is CudaTimer
 * (utils/timers.py:11-25) around the whole forward. */
int evk_model_profile(evk_model* m, const float* voxel, float* image, void* stream, int max_ops, float* host_ms,
                      double* host_flops, int* n_ops);
/* Human-readable description of launch `index` of one forward ("conv3x3 s1 128+128->512 lstm @46x60"). */
int evk_model_op_desc(evk_model* m, int index, char* buf, int cap);

/* One ConvLayer (model/submodules.py:8-35: conv2d [+ residual] + activation) on NHWC float32 device tensors,
 * the building block of every network above, exposed for layer-level parity tests and external callers.
 * x: [N,H,W,Cin]; w: HOST [Cout,Cin,k,k] (torch layout); bias: HOST [Cout] or NULL; res: device [N,Ho,Wo,Cout] or
 * NULL (added before the activation); act: 0 none, 1 relu, 2 sigmoid, 3 tanh; y: [N,Ho,Wo,Cout].
 * precision 0 = tcgen05 split-bf16 tensor-core kernel when Cin % 32 == 0, else / precision 1 = fp32 CUDA cores.
 * Synchronises the stream. */
int evk_conv2d_nhwc(const float* x, int N, int H, int W, int Cin, const float* w_oihw_host, const float* bias_host, int Cout,
                    int k, int stride, int pad, int act, const float* res, int precision, float* y, void* stream);

/* Host-only weight re-packers of the re-shaped layers, exposed so that their index maps can be checked without a GPU
 * (tests/test_weight_packers.py compares them with torch's own operators).  w: HOST torch layout [Cout,Cin,kh,kw].
 *  kind 0: UpsampleConvLayer (model/submodules.py:69-97), four stacked output phases  -> [25*Cin][4*Cout]
 *  kind 1: its border corrections, two 1x5 line convolutions, negated                 -> [2][5*Cin][4*Cout]
 *  kind 2: stride-2 5x5 ConvLayer over pixel pairs (first encoder)                    -> [kh*3*2*Cin][Cout]
 *  kind 3: 3x3 ConvLayer over 4-pixel windows of 16-channel tensors, `group` output pixels per row (FireNet;
 *          Cin = 16 or 32 = cat(x, h))                                                -> [kh*(Cin/16)*64][group*Cout]
 *  kind 4: mixed-operand form of a plain layer (fp16 product + two fp8 products, kh*kw*Cin % 64 == 0), DECODED back to
 *          floats: [3][kh*kw*Cin][Cout] = the fp16 part, the e4m3 remainder, the e4m3 copy (per-channel power-of-two scales undone)
 * Row index k = position in the GEMM's K dimension, column = packed output channel; *out_len = elements written
 * (EVK_ERR_ARG if out_cap is too small).  No CUDA call is made. */
int evk_pack_layer_weights(int kind, const float* w_oihw_host, int Cout, int Cin, int kh, int kw, int group, float* out_host,
                           int64_t out_cap, int64_t* out_len);

/* ------------------------------------------------------------ post-process --
 * post_process_normalization / normalize   reference: eval.py:380-395,
 * utils/eval_utils.py:15-35.  out = (v - P_qmin) / (P_qmax - P_qmin) with
 * numpy's linear-interpolated percentiles; apply_exp=1 first takes exp(v)
 * ('exprobust').  img/out: [n_images, numel] float32 (may alias). */
int evk_percentile_normalize(const float* img, float* out, int n_images, int numel,
                             double q_min, double q_max, int apply_exp, void* stream);

/* CropParameters.crop (utils/util.py:58-59): [n,C,Hp,Wp] -> [n,C,H,W] centre crop. */
int evk_crop(const float* in, float* out, int n, int C, int Hp, int Wp, int H, int W, void* stream);

/* ---------------------------------------------------------------- stage 3 --
 * MseMetric / SsimMetric .calculate(img, ref)   reference: utils/eval_metrics.py:77-97
 * (+ the [0,1] clip of EvalMetricsTracker.update :253-255 when clip != 0).
 * img, ref: [n_images,H,W] float32.  scores: [n_images,2] float64 = (mse, ssim).
 * SSIM = skimage.structural_similarity(gaussian_weights, sigma 1.5, data_range 1).
 * H, W must be >= 11. */
int evk_mse_ssim(const float* img, const float* ref, int n_images, int H, int W, int clip,
                 double* scores, void* stream);

/* LPIPS   reference: utils/eval_metrics.py:100-156 (PyIqaMetricFactory: `pyiqa.create_metric('lpips')` on frames
 * repeated to 3 channels by cv2torch(num_ch=3), utils/eval_utils.py:46-54, in queues of 4).  pyiqa and its weights are
 * a third-party dependency absent from the reference tree: this is the published LPIPS v0.1 algorithm on the same
 * convolution kernels, with weights supplied by the caller under the lpips / pyiqa state_dict names
 * ("net.slice{s}.{i}.weight|bias" or torchvision "features.{i}.*", "lin{k}.model.1.weight" or "lins.{k}...").
 * backbone: 0 = AlexNet (pyiqa 'lpips'), 1 = VGG16 ('lpips-vgg').  batch = pairs per call (the reference queues 4).
 * img, ref: [n, H, W] float32 grey frames in [0,1] on the device; scores: [n] float64 on the device. */
typedef struct evk_lpips evk_lpips;
int evk_lpips_create(int backbone, int batch, int H, int W, int precision, evk_lpips** out);
int evk_lpips_load_tensor(evk_lpips* l, const char* name, const float* host_data, const int64_t* shape, int ndim);
int evk_lpips_finalize(evk_lpips* l);
int evk_lpips_forward(evk_lpips* l, const float* img, const float* ref, int n, double* scores, void* stream);
double evk_lpips_flops(evk_lpips* l);
int evk_lpips_num_tc_layers(evk_lpips* l);
int evk_lpips_destroy(evk_lpips* l);

/* uint8 frame -> float32 / 255  (dataset.py:84) */
int evk_u8_to_f32(const uint8_t* in, float* out, int64_t numel, void* stream);
/* n_frames frames of numel bytes each (HOST array of device pointers) -> out [n_frames, numel] float32, one launch */
int evk_u8_to_f32_batch(const uint8_t* const* frames, int n_frames, int64_t numel, float* out, void* stream);

/* EvalMetricsTracker.histogram_equalization with hist_eq == 'global' (utils/eval_metrics.py:326-331):
 * skimage.exposure.equalize_hist (256 bins over [min, max], cdf, np.interp over the bin centres) -> float32.
 * img/out: [n_images, numel] float32 (may alias); clip != 0 first clamps to [0,1]. */
int evk_equalize_hist(const float* img, float* out, int n_images, int numel, int clip, void* stream);

/* EvalMetricsTracker.histogram_equalization with hist_eq == 'local' (utils/eval_metrics.py:332-339):
 * img_as_float32(skimage.filters.rank.equalize(img_as_ubyte(img), footprint=disk(radius))), radius = 55 in the reference:
 * per pixel, over the disk's pixels inside the image, uint8(255 * #{v <= own grey level} / population) / 255.
 * img/out: [n_images, H, W] float32, NOT aliased; clip != 0 first clamps to [0,1]. */
int evk_equalize_local(const float* img, float* out, int n_images, int H, int W, int radius, int clip, void* stream);

/* float32 frame -> uint8 for the PNG writer: uint8(round_half_even(clip(v, 0, 1) * 255))
 * (save_inferred_image, utils/eval_utils.py:80-84, after the clip of EvalMetricsTracker.update, utils/eval_metrics.py:253-255) */
int evk_quantize_u8(const float* in, uint8_t* out, int64_t numel, void* stream);

/* np.searchsorted over DEVICE-resident float64 event timestamps: the 't_seconds' window boundaries of
 * MemMapDataset.compute_timeblock_indices (dataset.py:104-117, `np.searchsorted(self.filehandle["t"], end_time)`, side
 * 'left') without a host copy of the timestamps.  t [n] ascending, values [m] (the reference's float64 expression
 * ((t - sw) * i + t0) + t, evaluated by the caller), out [m] int64, all on the device; right != 0 selects side 'right'.
 * Float64 comparisons only: the indices are numpy's, bit for bit. */
int evk_searchsorted_f64(const double* t, int64_t n, const double* values, int64_t m, int right, int64_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EVREAL_B200_H */
