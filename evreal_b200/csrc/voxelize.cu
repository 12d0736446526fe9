// Stage 1: event window -> voxel grid, plus the glue around it
// (normalize_event_tensor, CropParameters.pad/crop, uint8 frame -> float).
//
// Reference semantics: utils/event_utils.py:4-59, dataset.py:52-58,222-228,
// eval.py:398-410, utils/util.py:30-59 (paths in the EVREAL tree).
//
// HBM-bound integer/float scatter: every event is read exactly once with
// 128-bit loads (4 events per thread per array) and contributes to at most two
// ADJACENT temporal bins.  The scatter target is a pixel-interleaved scratch grid
// [H][W][G][4] (G = (bins-1)/3 + 1 groups of four slots; group g holds bins 3g..3g+3, so every adjacent pair of
// bins shares one aligned 16-byte group) and each event is ONE fire-and-forget RED.ADD.V4.F32 into it: the L2
// reduction units are request-rate bound (~83 requests/clk measured, tools/microbench/red_bench.cu) and a v4
// request costs the same as a scalar one, so this halves the scatter time of the two-scalar-reds form.  A gather
// pass then writes the reference's planar [bins,H,W] layout (bin 3g = slot 0 of group g + slot 3 of group g-1).
// Small windows skip the scratch grid: below ~S/8 events (S = scratch bytes; 170k events at 240x180, 1.2M at 640x480; measured
// crossover at 640x480 between 0.8M and 2M) clearing and gathering the 1.6x larger interleaved grid plus two extra launches
// cost more than the second reduction per event, so each event issues TWO
// scalar RED.ADD.F32 straight into the (cleared) planar grid -- one launch, no gather pass.
// Algorithmic bytes per window: 16*N (f32 SoA) or 13*N (raw int16/f64/u8) + 4*bins*H*W for the grid.
#include <algorithm>

#include "evk_common.cuh"

namespace evk {

constexpr int kVoxThreads = 256;

struct VoxGeom {
    int bins, H, W;
    int groups;     // 16-byte slot groups per pixel in the scratch grid
};
static inline int vox_groups(int bins) { return (bins - 1) / 3 + 1; }

// One event's contribution.  tn is the normalised time in [0, bins-1]; only
// floor(tn) and floor(tn)+1 can have weight max(0, 1-|tn-b|) > 0, so the
// reference's five passes collapse to two adds with identical float values.
template <bool kDirect = false>
__device__ __forceinline__ void scatter_event(float xf, float yf, float tn, float pol, const VoxGeom g,
                                              float* __restrict__ scratch, int& oob) {
    int xi = (int)xf;   // truncation toward zero == tensor.long()
    int yi = (int)yf;
    if (xi < 0) xi += g.W;   // python-style wrap of negative indices (index_put_)
    if (yi < 0) yi += g.H;
    if ((unsigned)xi >= (unsigned)g.W || (unsigned)yi >= (unsigned)g.H) {
        oob++;
        return;
    }
    const float fb = floorf(tn);
    const int b0 = (int)fb;
    if (b0 < -1 || b0 >= g.bins) return;              // no bin within distance 1 (cannot happen for sorted windows)
    const int grp = max(b0, 0) / 3;
    // the reference's per-bin weight max(0, 1 - |tn - b|) for the two bins that can be non-zero (same float values)
    const float w0 = 1.0f - fabsf(tn - (float)b0), w1 = 1.0f - fabsf(tn - (float)(b0 + 1));
    const float c0 = (b0 >= 0 && w0 > 0.0f) ? pol * w0 : 0.0f;
    const float c1 = (b0 + 1 < g.bins && w1 > 0.0f) ? pol * w1 : 0.0f;
    if (kDirect) {                                    // planar grid [bins][H][W]: one scalar reduction per non-zero weight
        float* cell = scratch + ((size_t)max(b0, 0) * g.H + yi) * g.W + xi;
        if (b0 >= 0 && w0 > 0.0f) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cell), "f"(c0) : "memory");
        if (b0 + 1 < g.bins && w1 > 0.0f)
            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(cell + (b0 >= 0 ? (size_t)g.H * g.W : 0)), "f"(c1) : "memory");
        return;
    }
    const int o = b0 - 3 * grp;                       // slot of bin b0 in its group: -1 (b0 == -1), 0, 1 or 2
    const float v0 = o == 0 ? c0 : (o == -1 ? c1 : 0.0f);
    const float v1 = o == 1 ? c0 : (o == 0 ? c1 : 0.0f);
    const float v2 = o == 2 ? c0 : (o == 1 ? c1 : 0.0f);
    const float v3 = o == 2 ? c1 : 0.0f;
    float* cell = scratch + (((size_t)yi * g.W + xi) * g.groups + grp) * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cell), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
}

// scratch [n][H*W][G][4] -> planar [n][bins][H*W]
__global__ void __launch_bounds__(256) voxel_gather_kernel(const float* __restrict__ scratch, float* __restrict__ grid, VoxGeom g,
                                                           int64_t pixels, int n_grids) {
    const int64_t total = pixels * n_grids;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = i / pixels, pix = i - w * pixels;
        const float4* src = reinterpret_cast<const float4*>(scratch) + i * g.groups;
        float* dst = grid + (size_t)w * g.bins * pixels + pix;
        float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int gi = 0; gi < g.groups; ++gi) {
            const float4 c = __ldcs(src + gi);
            const int b = 3 * gi;
            if (b < g.bins) dst[(size_t)b * pixels] = gi > 0 ? c.x + prev.w : c.x;
            if (b + 1 < g.bins) dst[(size_t)(b + 1) * pixels] = c.y;
            if (b + 2 < g.bins) dst[(size_t)(b + 2) * pixels] = c.z;
            prev = c;
        }
    }
}

// torch.linspace(0, bins-1, n)[i] in float32 (scalar ATen formula).
__device__ __forceinline__ float linspace_at(int64_t i, int64_t n, float end) {
    if (n == 1) return 0.0f;
    const float step = end / (float)(n - 1);
    return (i < n / 2) ? (0.0f + step * (float)i) : (end - step * (float)(n - 1 - i));
}

struct TimeNorm {
    float t0, dt, scale;
    bool degenerate;   // dt < 1e-9 -> linspace branch (utils/event_utils.py:48-49)
    int64_t n;
    __device__ __forceinline__ float operator()(float t, int64_t i) const {
        if (degenerate) return linspace_at(i, n, scale);
        return __fmul_rn(__fdiv_rn(__fsub_rn(t, t0), dt), scale);
    }
};

template <bool kVec, bool kDirect>
__global__ void __launch_bounds__(kVoxThreads)
voxelize_f32_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ t,
                    const float* __restrict__ p, int64_t n, int64_t head, VoxGeom g,
                    float* __restrict__ grid, int* __restrict__ oob_count) {
    TimeNorm tn;
    tn.t0 = __ldg(t);
    tn.dt = __fsub_rn(__ldg(t + n - 1), tn.t0);
    tn.scale = (float)(g.bins - 1);
    tn.degenerate = (double)tn.dt < 1e-9;
    tn.n = n;
    int oob = 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    if (kVec) {
        // unaligned head and tail handled scalar; the body is 128-bit aligned
        const int64_t nvec = (n - head) / 4;
        const float4* x4 = reinterpret_cast<const float4*>(x + head);
        const float4* y4 = reinterpret_cast<const float4*>(y + head);
        const float4* t4 = reinterpret_cast<const float4*>(t + head);
        const float4* p4 = reinterpret_cast<const float4*>(p + head);
        for (int64_t v = tid; v < nvec; v += nthreads) {
            const float4 xv = __ldcs(x4 + v), yv = __ldcs(y4 + v), tv = __ldcs(t4 + v), pv = __ldcs(p4 + v);
            const int64_t i = head + v * 4;
            scatter_event<kDirect>(xv.x, yv.x, tn(tv.x, i + 0), pv.x, g, grid, oob);
            scatter_event<kDirect>(xv.y, yv.y, tn(tv.y, i + 1), pv.y, g, grid, oob);
            scatter_event<kDirect>(xv.z, yv.z, tn(tv.z, i + 2), pv.z, g, grid, oob);
            scatter_event<kDirect>(xv.w, yv.w, tn(tv.w, i + 3), pv.w, g, grid, oob);
        }
        const int64_t tail0 = head + nvec * 4;
        const int64_t nscalar = head + (n - tail0);
        for (int64_t s = tid; s < nscalar; s += nthreads) {
            const int64_t i = s < head ? s : tail0 + (s - head);
            scatter_event<kDirect>(x[i], y[i], tn(t[i], i), p[i], g, grid, oob);
        }
    } else {
        for (int64_t i = tid; i < n; i += nthreads) scatter_event<kDirect>(x[i], y[i], tn(t[i], i), p[i], g, grid, oob);
    }
    if (oob_count != nullptr && oob > 0) atomicAdd(oob_count, oob);
}

// Raw on-disk layout: xy int16 pairs, t float64 absolute, pol uint8 {0,1}.
template <bool kDirect>
__device__ __forceinline__ void voxelize_raw_body(const int16_t* __restrict__ xy, const double* __restrict__ t,
                                                  const uint8_t* __restrict__ pol, int64_t n, VoxGeom g, float* __restrict__ grid,
                                                  int* __restrict__ oob_count, int64_t tid, int64_t nthreads) {
    const double t0d = __ldg(t);
    TimeNorm tn;
    tn.t0 = 0.0f;   // (t - t[0]).astype(f32)[0] == 0
    tn.dt = (float)(__ldg(t + n - 1) - t0d);
    tn.scale = (float)(g.bins - 1);
    tn.degenerate = (double)tn.dt < 1e-9;
    tn.n = n;
    int oob = 0;
    const int* xy32 = reinterpret_cast<const int*>(xy);   // int16 pairs are 4-byte aligned by construction
    for (int64_t i = tid; i < n; i += nthreads) {
        const int c = __ldcs(xy32 + i);
        const float xf = (float)(short)(c & 0xffff);
        const float yf = (float)(short)(c >> 16);
        const float tf = (float)(__ldcs(t + i) - t0d);
        const float pf = (float)((double)pol[i] * 2.0 - 1.0);
        scatter_event<kDirect>(xf, yf, tn(tf, i), pf, g, grid, oob);
    }
    if (oob_count != nullptr && oob > 0) atomicAdd(oob_count, oob);
}

template <bool kDirect>
__global__ void __launch_bounds__(kVoxThreads)
voxelize_raw_kernel(const int16_t* __restrict__ xy, const double* __restrict__ t, const uint8_t* __restrict__ pol,
                    int64_t n, VoxGeom g, float* __restrict__ grid, int* __restrict__ oob_count) {
    voxelize_raw_body<kDirect>(xy, t, pol, n, g, grid, oob_count, (int64_t)blockIdx.x * blockDim.x + threadIdx.x,
                               (int64_t)gridDim.x * blockDim.x);
}

// Several windows (one per sequence of a lock-step batch) in ONE launch: blockIdx.y = window, grids contiguous.
constexpr int kVoxBatch = 32;
struct VoxBatch {
    const int16_t* xy[kVoxBatch];
    const double* t[kVoxBatch];
    const uint8_t* pol[kVoxBatch];
    long long n[kVoxBatch];
};

template <bool kDirect>
__global__ void __launch_bounds__(kVoxThreads)
voxelize_raw_batch_kernel(const __grid_constant__ VoxBatch wb, VoxGeom g, float* __restrict__ grids, int* __restrict__ oob_count) {
    const int w = blockIdx.y;
    const int64_t n = wb.n[w];
    if (n <= 0) return;                      // empty window: the (pre-zeroed) grid stays zero (dataset.py:59-71)
    const size_t per_grid = kDirect ? (size_t)g.bins * g.H * g.W : (size_t)4 * g.groups * g.H * g.W;
    voxelize_raw_body<kDirect>(wb.xy[w], wb.t[w], wb.pol[w], n, g, grids + (size_t)w * per_grid, oob_count,
                               (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// Direct (two scalar reductions per event, planar grid) below this many events per window: S / 8 with S = bytes of the
// interleaved scratch grid (see the file header); EVK_VOX_DIRECT_MAX overrides (0 = never, <0 = always).
static int64_t vox_direct_max(const VoxGeom& g) {
    static const char* e = getenv("EVK_VOX_DIRECT_MAX");
    if (e && e[0]) { const long long v = atoll(e); return v < 0 ? INT64_MAX : (int64_t)v; }
    return (int64_t)((size_t)16 * g.groups * g.H * g.W / 8);
}

static int vox_grid_blocks(int64_t n, int per_thread) {
    int64_t blocks = ceil_div64(n, (int64_t)kVoxThreads * per_thread);
    const int64_t cap = (int64_t)kNumSMs * 8;   // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// stream-ordered scratch (no library-global state): allocate + zero, ... scatter ..., gather into the planar grid + free
static int vox_scratch_begin(float** scratch, const VoxGeom& g, int n_grids, cudaStream_t st) {
    const size_t bytes = sizeof(float) * 4 * g.groups * (size_t)g.H * g.W * n_grids;
    // keep freed blocks in the device's default pool (the default release threshold of 0 hands them back to the driver
    // at every synchronisation point, which turns each call into a real allocation: measured 100+ us)
    static bool pool_ready[64] = {false};
    int dev = 0;
    EVK_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !pool_ready[dev]) {
        cudaMemPool_t pool;
        EVK_CHECK_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        EVK_CHECK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pool_ready[dev] = true;
    }
    EVK_CHECK_CUDA(cudaMallocAsync((void**)scratch, bytes, st));
    EVK_CHECK_CUDA(cudaMemsetAsync(*scratch, 0, bytes, st));
    return EVK_OK;
}
static int vox_scratch_end(float* scratch, float* grid, const VoxGeom& g, int n_grids, cudaStream_t st) {
    const int64_t pixels = (int64_t)g.H * g.W;
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div64(pixels * n_grids, 256), (int64_t)kNumSMs * 8);
    voxel_gather_kernel<<<blocks, 256, 0, st>>>(scratch, grid, g, pixels, n_grids);
    EVK_CHECK_CUDA(cudaGetLastError());
    EVK_CHECK_CUDA(cudaFreeAsync(scratch, st));
    return EVK_OK;
}

int voxelize_f32(const float* x, const float* y, const float* t, const float* p, int64_t n, int bins, int H, int W,
                 float* grid, int* oob_count, cudaStream_t st) {
    EVK_REQUIRE(n > 0, EVK_ERR_ARG, "evk_voxelize: empty window (n=%lld); the reference indexes ts[-1]", (long long)n);
    EVK_REQUIRE(bins > 0 && H > 0 && W > 0, EVK_ERR_ARG, "evk_voxelize: bad geometry bins=%d H=%d W=%d", bins, H, W);
    VoxGeom g{bins, H, W, vox_groups(bins)};
    const bool direct = n <= vox_direct_max(g);
    float* scratch = nullptr;
    if (direct) {
        EVK_CHECK_CUDA(cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)bins * H * W, st));
        scratch = grid;
    } else {
        int r = vox_scratch_begin(&scratch, g, 1, st);
        if (r != EVK_OK) return r;
    }
    // all four arrays must share the same 16-byte phase for the vector body
    auto phase = [](const void* q) { return (int)(((uintptr_t)q >> 2) & 3); };
    const bool same = phase(x) == phase(y) && phase(y) == phase(t) && phase(t) == phase(p) &&
                      (((uintptr_t)x | (uintptr_t)y | (uintptr_t)t | (uintptr_t)p) & 3) == 0;
    if (same && n >= 64) {
        int64_t head = (4 - phase(x)) & 3;
        if (direct) voxelize_f32_kernel<true, true><<<vox_grid_blocks(n, 4), kVoxThreads, 0, st>>>(x, y, t, p, n, head, g, scratch, oob_count);
        else voxelize_f32_kernel<true, false><<<vox_grid_blocks(n, 4), kVoxThreads, 0, st>>>(x, y, t, p, n, head, g, scratch, oob_count);
    } else {
        if (direct) voxelize_f32_kernel<false, true><<<vox_grid_blocks(n, 1), kVoxThreads, 0, st>>>(x, y, t, p, n, 0, g, scratch, oob_count);
        else voxelize_f32_kernel<false, false><<<vox_grid_blocks(n, 1), kVoxThreads, 0, st>>>(x, y, t, p, n, 0, g, scratch, oob_count);
    }
    EVK_CHECK_CUDA(cudaGetLastError());
    return direct ? EVK_OK : vox_scratch_end(scratch, grid, g, 1, st);
}

int voxelize_raw(const int16_t* xy, const double* t, const uint8_t* pol, int64_t n, int bins, int H, int W,
                 float* grid, int* oob_count, cudaStream_t st) {
    EVK_REQUIRE(n > 0, EVK_ERR_ARG, "evk_voxelize_raw: empty window");
    EVK_REQUIRE(bins > 0 && H > 0 && W > 0, EVK_ERR_ARG, "evk_voxelize_raw: bad geometry");
    EVK_REQUIRE(((uintptr_t)xy & 3) == 0 && ((uintptr_t)t & 7) == 0, EVK_ERR_ARG, "evk_voxelize_raw: misaligned arrays");
    VoxGeom g{bins, H, W, vox_groups(bins)};
    if (n <= vox_direct_max(g)) {
        EVK_CHECK_CUDA(cudaMemsetAsync(grid, 0, sizeof(float) * (size_t)bins * H * W, st));
        voxelize_raw_kernel<true><<<vox_grid_blocks(n, 2), kVoxThreads, 0, st>>>(xy, t, pol, n, g, grid, oob_count);
        EVK_CHECK_CUDA(cudaGetLastError());
        return EVK_OK;
    }
    float* scratch = nullptr;
    int r = vox_scratch_begin(&scratch, g, 1, st);
    if (r != EVK_OK) return r;
    voxelize_raw_kernel<false><<<vox_grid_blocks(n, 2), kVoxThreads, 0, st>>>(xy, t, pol, n, g, scratch, oob_count);
    EVK_CHECK_CUDA(cudaGetLastError());
    return vox_scratch_end(scratch, grid, g, 1, st);
}

int voxelize_raw_batch(const evk_event_window* windows, int n_windows, int bins, int H, int W, float* grids, int* oob_count,
                       cudaStream_t st) {
    EVK_REQUIRE(windows && grids && n_windows > 0, EVK_ERR_ARG, "evk_voxelize_raw_batch: bad argument");
    EVK_REQUIRE(bins > 0 && H > 0 && W > 0, EVK_ERR_ARG, "evk_voxelize_raw_batch: bad geometry");
    VoxGeom g{bins, H, W, vox_groups(bins)};
    int64_t n_largest = 0;
    for (int i = 0; i < n_windows; ++i) n_largest = std::max<int64_t>(n_largest, windows[i].n);
    const bool direct = n_largest <= vox_direct_max(g);
    const size_t scratch_elems = direct ? (size_t)bins * H * W : (size_t)4 * g.groups * H * W;
    float* scratch = nullptr;
    if (direct) {
        EVK_CHECK_CUDA(cudaMemsetAsync(grids, 0, sizeof(float) * scratch_elems * n_windows, st));
        scratch = grids;
    } else {
        int r = vox_scratch_begin(&scratch, g, n_windows, st);
        if (r != EVK_OK) return r;
    }
    for (int w0 = 0; w0 < n_windows; w0 += kVoxBatch) {
        const int cnt = std::min(kVoxBatch, n_windows - w0);
        VoxBatch wb;
        int64_t nmax = 0;
        for (int i = 0; i < kVoxBatch; ++i) {
            if (i < cnt) {
                const evk_event_window& e = windows[w0 + i];
                EVK_REQUIRE(e.n >= 0 && (e.n == 0 || (e.xy && e.t && e.pol)), EVK_ERR_ARG, "evk_voxelize_raw_batch: window %d is null", w0 + i);
                EVK_REQUIRE(((uintptr_t)e.xy & 3) == 0 && ((uintptr_t)e.t & 7) == 0, EVK_ERR_ARG, "evk_voxelize_raw_batch: misaligned arrays");
                wb.xy[i] = e.xy; wb.t[i] = e.t; wb.pol[i] = e.pol; wb.n[i] = e.n;
                nmax = std::max<int64_t>(nmax, e.n);
            } else {
                wb.xy[i] = nullptr; wb.t[i] = nullptr; wb.pol[i] = nullptr; wb.n[i] = 0;
            }
        }
        if (nmax == 0) continue;
        int64_t bx = ceil_div64(nmax, (int64_t)kVoxThreads * 2);
        const int64_t cap = std::max<int64_t>(1, (int64_t)kNumSMs * 8 / cnt);
        bx = std::max<int64_t>(1, std::min(bx, cap));
        if (direct) voxelize_raw_batch_kernel<true><<<dim3((unsigned)bx, (unsigned)cnt), kVoxThreads, 0, st>>>(wb, g, scratch + scratch_elems * w0, oob_count);
        else voxelize_raw_batch_kernel<false><<<dim3((unsigned)bx, (unsigned)cnt), kVoxThreads, 0, st>>>(wb, g, scratch + scratch_elems * w0, oob_count);
        EVK_CHECK_CUDA(cudaGetLastError());
    }
    return direct ? EVK_OK : vox_scratch_end(scratch, grids, g, n_windows, st);
}

// ---------------------------------------------------------------------------
// normalize_event_tensor (eval.py:398-410) fused with CropParameters.pad.
// Pass 1: per-sample (count of non-zeros, sum, sum of squares) in float64.
// Pass 2: out = mask * (v - mean) / std written into the zero-padded frame.
// ---------------------------------------------------------------------------
struct EvStats {
    double sum, sumsq;
    unsigned long long nnz;
};

__global__ void __launch_bounds__(256) event_stats_kernel(const float* __restrict__ in, int64_t numel, EvStats* stats) {
    const int s = blockIdx.y;
    const float* v = in + (size_t)s * numel;
    double sum = 0.0, sq = 0.0;
    unsigned int nnz = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = v[i];
        if (a != 0.0f) {
            nnz++;
            sum += (double)a;
            sq += (double)a * (double)a;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        nnz += __shfl_xor_sync(0xffffffffu, nnz, o);
    }
    __shared__ double ssum[8], ssq[8];
    __shared__ unsigned int snz[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { ssum[warp] = sum; ssq[warp] = sq; snz[warp] = nnz; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { sum += ssum[w]; sq += ssq[w]; nnz += snz[w]; }
        atomicAdd(&stats[s].sum, sum);
        atomicAdd(&stats[s].sumsq, sq);
        atomicAdd(&stats[s].nnz, (unsigned long long)nnz);
    }
}

__global__ void __launch_bounds__(256)
normalize_pad_kernel(const float* __restrict__ in, float* __restrict__ out, const EvStats* __restrict__ stats,
                     int C, int H, int W, int Hp, int Wp, int top, int left, int do_norm) {
    const int s = blockIdx.y;
    const int64_t total = (int64_t)C * Hp * Wp;
    float mean = 0.0f, stddev = 1.0f;
    bool active = false;
    if (do_norm) {
        const EvStats st = stats[s];
        if (st.nnz > 0) {
            active = true;
            const float nf = (float)st.nnz;
            mean = (float)st.sum / nf;
            stddev = sqrtf((float)st.sumsq / nf - mean * mean);
            stddev = fmaxf(stddev, 1e-6f);
        }
    }
    if ((Wp & 3) == 0) {
        // four output columns per thread (one 16-byte store; the index arithmetic of the scalar form cost more than the bytes)
        const int W4 = Wp >> 2;
        const int64_t total4 = (int64_t)C * Hp * W4;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
            const int x4 = (int)(i % W4);
            const int64_t row = i / W4;
            const int yo = (int)(row % Hp);
            const int c = (int)(row / Hp);
            const int yi = yo - top;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if ((unsigned)yi < (unsigned)H) {
                const float* src = in + ((size_t)s * C + c) * H * W + (size_t)yi * W;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int xi = x4 * 4 + k - left;
                    if ((unsigned)xi < (unsigned)W) {
                        const float a = src[xi];
                        v[k] = (active && a != 0.0f) ? __fdiv_rn(__fsub_rn(a, mean), stddev) : (active ? 0.0f : a);
                    }
                }
            }
            reinterpret_cast<float4*>(out + (size_t)s * total)[i] = make_float4(v[0], v[1], v[2], v[3]);
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % Wp);
        const int yo = (int)((i / Wp) % Hp);
        const int c = (int)(i / ((int64_t)Wp * Hp));
        const int yi = yo - top, xi = xo - left;
        float v = 0.0f;
        if ((unsigned)yi < (unsigned)H && (unsigned)xi < (unsigned)W) {
            v = in[((size_t)s * C + c) * H * W + (size_t)yi * W + xi];
            if (active) v = (v != 0.0f) ? __fdiv_rn(__fsub_rn(v, mean), stddev) : 0.0f;
        }
        out[(size_t)s * total + i] = v;
    }
}

int normalize_pad(const float* in, float* out, int n_samples, int C, int H, int W, int Hp, int Wp, int do_norm,
                  cudaStream_t st) {
    EVK_REQUIRE(n_samples > 0 && C > 0 && H > 0 && W > 0 && Hp >= H && Wp >= W, EVK_ERR_ARG,
                "evk_normalize_pad: bad shape n=%d C=%d %dx%d -> %dx%d", n_samples, C, H, W, Hp, Wp);
    EVK_REQUIRE(!(in == out && (Hp != H || Wp != W)), EVK_ERR_ARG, "evk_normalize_pad: in-place padding is not possible");
    EvStats* stats = nullptr;
    if (do_norm) {
        EVK_CHECK_CUDA(cudaMallocAsync(&stats, sizeof(EvStats) * n_samples, st));
        EVK_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(EvStats) * n_samples, st));
        const int64_t numel = (int64_t)C * H * W;
        // few blocks per sample: every block ends in three contended float64 atomics on the sample's record
        dim3 grid((unsigned)std::min<int64_t>(ceil_div64(numel, 256 * 4), n_samples >= 8 ? 32 : 128), n_samples);
        event_stats_kernel<<<grid, 256, 0, st>>>(in, numel, stats);
        EVK_CHECK_CUDA(cudaGetLastError());
    }
    // CropParameters: top/left get the ceil half (utils/util.py:43-46)
    const int top = (Hp - H + 1) / 2, left = (Wp - W + 1) / 2;
    const int64_t total = (int64_t)C * Hp * Wp;
    const int64_t work = (Wp & 3) == 0 ? total / 4 : total;
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(work, 256), std::max(1, kNumSMs * 8 / n_samples))), n_samples);
    normalize_pad_kernel<<<grid, 256, 0, st>>>(in, out, stats, C, H, W, Hp, Wp, top, left, do_norm);
    EVK_CHECK_CUDA(cudaGetLastError());
    if (stats) EVK_CHECK_CUDA(cudaFreeAsync(stats, st));
    return EVK_OK;
}

// CropParameters.crop: iy0 = floor(Hp/2) - floor(H/2) (utils/util.py:50-59)
__global__ void __launch_bounds__(256)
crop_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t planes, int Hp, int Wp, int H, int W, int iy0, int ix0) {
    const int64_t total = planes * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int64_t pl = i / ((int64_t)W * H);
        out[i] = in[(pl * Hp + (y + iy0)) * Wp + (x + ix0)];
    }
}

int crop(const float* in, float* out, int n, int C, int Hp, int Wp, int H, int W, cudaStream_t st) {
    EVK_REQUIRE(n > 0 && C > 0 && Hp >= H && Wp >= W && H > 0 && W > 0, EVK_ERR_ARG, "evk_crop: bad shape");
    const int iy0 = Hp / 2 - H / 2, ix0 = Wp / 2 - W / 2;
    const int64_t total = (int64_t)n * C * H * W;
    crop_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1184), 256, 0, st>>>(in, out, (int64_t)n * C, Hp, Wp, H, W, iy0, ix0);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __fdiv_rn((float)in[i], 255.0f);
}

constexpr int kU8Batch = 64;
struct U8Batch { const uint8_t* src[kU8Batch]; };

__global__ void __launch_bounds__(256) u8_to_f32_batch_kernel(const __grid_constant__ U8Batch b, float* __restrict__ out, int64_t n) {
    const uint8_t* in = b.src[blockIdx.y];
    float* o = out + (size_t)blockIdx.y * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        o[i] = __fdiv_rn((float)in[i], 255.0f);
}

int u8_to_f32_batch(const uint8_t* const* frames, int n_frames, int64_t numel, float* out, cudaStream_t st) {
    EVK_REQUIRE(frames && out && n_frames > 0 && numel > 0, EVK_ERR_ARG, "evk_u8_to_f32_batch: bad argument");
    for (int f0 = 0; f0 < n_frames; f0 += kU8Batch) {
        const int cnt = std::min(kU8Batch, n_frames - f0);
        U8Batch b;
        for (int i = 0; i < kU8Batch; ++i) {
            b.src[i] = i < cnt ? frames[f0 + i] : nullptr;
            EVK_REQUIRE(i >= cnt || b.src[i] != nullptr, EVK_ERR_ARG, "evk_u8_to_f32_batch: frame %d is null", f0 + i);
        }
        const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(numel, 256 * 4), kNumSMs * 8 / cnt));
        u8_to_f32_batch_kernel<<<dim3(bx, (unsigned)cnt), 256, 0, st>>>(b, out + (size_t)f0 * numel, numel);
        EVK_CHECK_CUDA(cudaGetLastError());
    }
    return EVK_OK;
}

// save_inferred_image's quantisation (utils/eval_utils.py:80-84) after the tracker's clip (utils/eval_metrics.py:253-255):
// uint8(np.round(clip(v, 0, 1) * 255)); np.round rounds half to even = rintf
__global__ void __launch_bounds__(256) quantize_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)rintf(__fmul_rn(fminf(fmaxf(in[i], 0.0f), 1.0f), 255.0f));
}

int quantize_u8(const float* in, uint8_t* out, int64_t n, cudaStream_t st) {
    EVK_REQUIRE(n > 0, EVK_ERR_ARG, "evk_quantize_u8: empty input");
    quantize_u8_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 1184), 256, 0, st>>>(in, out, n);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

int u8_to_f32(const uint8_t* in, float* out, int64_t n, cudaStream_t st) {
    EVK_REQUIRE(n > 0, EVK_ERR_ARG, "evk_u8_to_f32: empty input");
    u8_to_f32_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 1184), 256, 0, st>>>(in, out, n);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// Device-side np.searchsorted over the resident float64 timestamps: the 't_seconds' window boundaries of
// MemMapDataset.compute_timeblock_indices (dataset.py:104-117) without a host copy of the event timestamps.  The query
// values are the reference's own float64 expression ((t - sw) * i + t0) + t, evaluated by the caller; the search compares
// float64 values only, so the indices are the ones numpy returns (side = 'left': first index with t[idx] >= v,
// 'right': first index with t[idx] > v; NaN sorts last as in numpy).
__global__ void __launch_bounds__(128) searchsorted_f64_kernel(const double* __restrict__ t, int64_t n, const double* __restrict__ values,
                                                               int64_t m, int right, long long* __restrict__ out) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < m; q += (int64_t)gridDim.x * blockDim.x) {
        const double v = values[q];
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = lo + ((hi - lo) >> 1);
            const double a = __ldg(t + mid);
            // numpy's ordering with NaN last: a < v  <=>  a < v || (v is NaN && a is not)
            const bool below = right ? !(v < a || (a != a && v == v)) : (a < v || (v != v && a == a));
            if (below) lo = mid + 1; else hi = mid;
        }
        out[q] = (long long)lo;
    }
}

int searchsorted_f64(const double* t, int64_t n, const double* values, int64_t m, int right, long long* out, cudaStream_t st) {
    EVK_REQUIRE(n >= 0 && m > 0, EVK_ERR_ARG, "evk_searchsorted_f64: n = %lld, m = %lld", (long long)n, (long long)m);
    searchsorted_f64_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(m, 128), kNumSMs * 8), 128, 0, st>>>(t, n, values, m, right, out);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
