// Tensor-core implicit-GEMM convolution for sm_100a: TMA-staged activation / weight tiles, tcgen05.mma with the
// fp32 accumulator in tensor memory, and the ConvLayer / ConvLSTM epilogues fused behind tcgen05.ld.
//
// Reference semantics: model/submodules.py:8-35 (ConvLayer), :152-184 (ResidualBlock), :187-245 (ConvLSTM),
// :69-97 (the 5x5 convolution of UpsampleConvLayer).  fp32 parity (1e-4, north_star) rules out single-pass
// bf16/tf32 operands (2e-3 / 3e-4), so every operand travels as TWO bf16 planes (hi = bf16(v),
// lo = bf16(v - hi)) and each K step issues three products into the same TMEM accumulator:
//     D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo          (error ~ 2^-16 relative; measured 4-6e-6 end to end)
//
// GEMM view: M = 128 output pixels of one image, N = BN output channels, K = kh*kw*(c1+c2).
//
// Operand traffic (the per-tap im2col re-load of the first version ran at the chip's ~6.3 kB/clk L2 -> SM ceiling):
//
//  * A (activations), HALO BOXES.  The M tile is 8 pixels along an "atom" axis U by 16 pixels along a "shift" axis V
//    (U = x, V = y or the transpose, whichever pads the layer less).  8 pixels x BK channels of bf16 are exactly one
//    swizzle atom (8 rows of the K-major UMMA layout), so ONE tiled TMA load of the box [BK, 8, 16 + kV - 1]
//    lands (16 + kV - 1) atoms in shared memory, and a kernel tap that shifts the window by j pixels along V is
//    the SAME data read through a shared-memory descriptor that starts j atoms later -- no re-load.  Only the kU
//    shifts along U need their own load: kU loads of (16+kV-1) atoms per channel chunk instead of kU*kV loads of
//    16 atoms (5x5: 100 instead of 400 atoms; 3x3: 54 instead of 144).  Stride 2 = element strides in the box and
//    one box per parity of the V tap.  Out-of-bounds zero fill is the convolution's zero padding.  cat(x, h) is
//    two tensor maps.
//  * B (weights), CLUSTER MULTICAST.  The CTAs of a thread-block cluster work on different M tiles of the same
//    N tile; each loads 1/cs of the weight tile and multicasts it into every CTA's shared memory.
//
// With the loads out of the way the kernel is bound by tcgen05.mma ISSUE (see the MMA issuer section below): two
// issuer warps, and shapes that keep every MMA wide and every K row 128 bytes: several output pixels per GEMM row (head,
// FireNet window mode), output phases stacked along N (upsample-conv decoders, poly.cu), row pairs for other cout = 32 layers.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over tiles): warp 0 = A producer, warp 3 = B producer,
// warps 1-2 = MMA issuers (K blocks dealt alternately; warp 2 also allocates TMEM), warps 4-11 = epilogue
// (thread = accumulator row = output pixel).  Two independent shared-memory rings (A: box stages, B: one K block per
// stage); the accumulator is double-buffered in tensor memory so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv.cuh"
#include "tc.cuh"

namespace evk {

struct TcArgs {
    int N, Hout, Wout, su, sv, pad_u, pad_v, kh, kw;       // Hout counts row PAIRS in row-pair mode; su / sv = stride along U / V
    int cw;                     // epilogue chunk width in accumulator columns: 32, or 16 for a 32-column linear tile (both halves of the epilogue warps get work)
    int rp, creal, hreal;       // row-pair mode, real channel count / output height (addressing)
    int s_pitch, s_grp, s_off;  // split OUTPUT tensors with padded rows (window-mode consumers, ConvParams::s_wp): elements per padded
                                // row, per GEMM row and of the left padding; 0 = dense [N,H,W,C]
    int fastps;                 // phase-stacked last decoder (cout 32, ReLU, fused prediction layer, no tensor output): straight-line epilogue
    int fastgru;                // ConvGRU epilogues, channel counts multiples of 16: straight-line variants (window-mode FireNet)
    int fastlin;                // EPI_LINEAR, wide, 32-column chunks, no row-pair / phase / prediction: straight-line epilogue
    int predw;                  // window mode (2 pixels x 16 channels per GEMM row, 16-column chunks): the 1x1 prediction layer of FireNet fused
                                //    into the straight-line epilogue -- a chunk is one pixel, its thread holds all 16 channels
    float act_floor;            // fastlin: lower clamp of the activation (0 for ReLU, -inf for none)
    int wide;                   // EPI_LINEAR: 256-bit stores (real and packed channel counts are multiples of 16)
    int ps, wreal;              // phase-stacked mode (ConvParams::phase4): column block a*2+b -> output pixel (2oy+a, 2ox+b); output width
    const float* ring_h; const float* ring_v;   // phase-stacked mode: border corrections (ConvParams), added before the activation
    int ux;                     // 1: U = x (tile 16 rows x 8 cols), 0: U = y (tile 8 rows x 16 cols)
    int tiles_u, tiles_v;       // tiles per image along U (8 px) and V (16 px)
    int ku, kv;                 // kernel extent along U and V
    int chunks1, chunks2;
    int bn, n_tiles, m_tiles, cout;
    int epi, act;
    int a_stages, b_stages;
    int tpb;                    // taps per K block (= per weight stage): 1, or all taps of a V group when they fit in 40 KB
    int ar;                     // atoms per A box
    int n_groups;               // V-tap groups per (chunk, U shift): 1 (stride 1) or 2 (stride 2: even / odd taps)
    int g_tap0[2], g_ntaps[2], g_step;
    int cs;                     // cluster size (CTAs sharing one weight tile)
    int issuers;                // MMA issuer warps (1 or 2)
    int own_acc;                // 1: every issuer has its own accumulator of acc_cols columns (free-running, summed by the
                                //    epilogue); 0: shared accumulator, strict block order by handshake
    int acc_cols;
    int acc_stride;             // TMEM columns per accumulator stage (= acc_cols, or 2 * acc_cols with own_acc)
    uint32_t tmem_cols;
    const float* bias;
    const float* iscale;        // MIXED kernels: 1 / S[n] per packed output column (ConvParams::w_iscale)
    int ys_mixed, hs_mixed;     // the split copy written by this epilogue is in the mixed (fp16 | fp8 pair) format
    const float* res;
    float* y;
    __nv_bfloat16* ys; long long ys_plane;
    const float* c_prev; float* c_new; float* h_new;
    __nv_bfloat16* hs_new; long long hs_plane;
    const float* h_prev; const float* u_in; float* u_out; float* hr_out; __nv_bfloat16* hrs_out;   // ConvGRU epilogues
    const float* pred_w; const float* pred_skip; float* pred_out; float pred_bias; int pred_sigmoid;
    const __nv_bfloat16* pred_skip_s; long long pred_skip_plane;
    int lean;                   // MIXED: the lean issuer loop (one tap per block, two free-running issuers)
    int deal;                   // free-running issuers: K blocks dealt singly (1) or in pairs (2)
    int poll;                   // look-ahead poll of the issuer's next weight barrier: 0 before the issue (try_wait), 1 none, 2 after it (test_wait)
    int hiprio;                 // 1: producer / MMA-issuer roles on the four HIGHEST warp ids (the scheduler favours high warp ids: the
                                //    issuers must not queue behind eight busy epilogue warps when tiles are short)
    int exp;                    // DBG kernels only (EVK_TC_EXP bit mask): 1 skip weight loads, 2 skip activation loads, 4 skip epilogue stores
    unsigned long long* dbg;    // EVK_TC_TIMING: per-CTA clock64 phase counters [grid][8], else nullptr
};

struct TcPlan {
    CUtensorMap tm_x1, tm_x2, tm_w;
    TcArgs a;
    int bk;
    bool mixed = false;
    dim3 grid;
    size_t smem;
};

constexpr int kTcThreads = 384;      // 12 warps: A producer, MMA, TMEM alloc, B producer, 8 epilogue
constexpr int kEpiWarps = 8;

// sigmoid / tanh on the SFU (ex2.approx, rcp.approx): absolute error ~2e-7, far inside the 1e-4 parity budget
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ float fast_act(float v, int act) {
    switch (act) {
        case ACT_RELU: return fmaxf(v, 0.0f);
        case ACT_SIGMOID: return fast_sigmoid(v);
        case ACT_TANH: return fast_tanh(v);
        default: return v;
    }
}

// 256-bit global stores (sm_100+): one full 32-byte sector per lane and instruction.  The epilogue thread owns one output
// pixel (= one row of channels), so its lanes never share a sector: with 8 / 16-byte stores the store path moves a
// quarter / half sector per request and small-K layers (head, first encoder) ran at the L1 store request rate.
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t* w) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" :: "l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// 16 channels held as fp32 bit patterns v[0..15]
__device__ __forceinline__ void store_mixed16(__nv_bfloat16* base, long long plane, size_t o, int c0, const uint32_t* v) {
    uint32_t h[8], x[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mixed_cvt2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]), h[i], x[i], l[i]);
    st_global_v8(base + o, h);
    uint8_t* b1 = reinterpret_cast<uint8_t*>(base + plane) + 2 * o - (size_t)(c0 & 63);
    *reinterpret_cast<uint4*>(b1) = make_uint4(x[0] | (x[1] << 16), x[2] | (x[3] << 16), x[4] | (x[5] << 16), x[6] | (x[7] << 16));
    *reinterpret_cast<uint4*>(b1 + 64) = make_uint4(l[0] | (l[1] << 16), l[2] | (l[3] << 16), l[4] | (l[5] << 16), l[6] | (l[7] << 16));
}

struct TileCoord { int nt, img, ou0, ov0; bool dummy; };

__device__ __forceinline__ TileCoord tile_coord(const TcArgs& a, int st, int crank) {
    TileCoord t;
    t.nt = st % a.n_tiles;
    int mt = (st / a.n_tiles) * a.cs + crank;
    t.dummy = mt >= a.m_tiles;            // cluster padding: runs the pipeline (peers multicast into it), stores nothing
    if (t.dummy) mt = a.m_tiles - 1;
    const int per_img = a.tiles_u * a.tiles_v;
    t.img = mt / per_img;
    const int rem = mt - t.img * per_img;
    t.ov0 = (rem / a.tiles_u) * 16;
    t.ou0 = (rem % a.tiles_u) * 8;
    return t;
}

// Phase-pair mode (a.ps == 2, U = y): N tile nt holds the two column phases of ROW phase nt, whose composite 5x5 kernel
// has an all-zero first (row phase 1) or last (row phase 0) tap row -- that U shift is skipped by all three pipeline roles.
// Single-phase mode (a.ps == 3, bn = cout): N tile nt is output phase nt = a*2+b; the zero tap COLUMN of column phase b is
// skipped as well (first / last V tap of every box), 16 of 25 taps per tile.
__device__ __forceinline__ void su_range(const TcArgs& a, int nt, int& su0, int& su1) {
    su0 = 0; su1 = a.ku;
    if (a.ps == 2) { if (nt == 0) su1 = a.ku - 1; else su0 = 1; }
    if (a.ps == 3) { if ((nt >> 1) == 0) su1 = a.ku - 1; else su0 = 1; }
}
__device__ __forceinline__ void tv_range(const TcArgs& a, int nt, int ntaps, int& tv0, int& tv1) {
    tv0 = 0; tv1 = ntaps;
    if (a.ps == 3) { if ((nt & 1) == 0) tv1 = ntaps - 1; else tv0 = 1; }
}

template <int BK, bool DBG, bool MIXED = false>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_x2,
               const __grid_constant__ CUtensorMap tm_w, const TcArgs a) {
    constexpr uint32_t ROW_BYTES = BK * 2;
    constexpr uint32_t ATOM = 8 * ROW_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_plane = (uint32_t)a.ar * ATOM, a_stage = 2 * a_plane;
    const uint32_t b_plane = (uint32_t)a.bn * ROW_BYTES, b_tap = 2 * b_plane, b_stage = (uint32_t)a.tpb * b_tap;
    const uint32_t smem_b = base + (uint32_t)a.a_stages * a_stage;
    const uint32_t bar_fa = smem_b + (uint32_t)a.b_stages * b_stage;
    const uint32_t bar_ea = bar_fa + 8u * a.a_stages;
    const uint32_t bar_fb = bar_ea + 8u * a.a_stages;
    const uint32_t bar_eb = bar_fb + 8u * a.b_stages;
    const uint32_t bar_tfull = bar_eb + 8u * a.b_stages;      // [2] accumulator ready
    const uint32_t bar_tempty = bar_tfull + 16u;              // [2] accumulator drained by the epilogue
    const uint32_t slot = bar_tempty + 16u;

    // logical warp (role) <- physical warp: by default physical warps 8..11 take the producer / issuer roles 0..3 and physical
    // warps 0..7 the epilogue roles 4..11 (an epilogue warp's tensor-memory lane quadrant is its PHYSICAL warp id % 4; role - 4
    // == physical id there, so the quadrant arithmetic below is unchanged)
    const int lane = threadIdx.x & 31;
    const int warp = a.hiprio ? (int)(((threadIdx.x >> 5) + 4) % 12) : (int)(threadIdx.x >> 5);
    const int cs = a.cs;
    const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / cs, ncl = gridDim.x / cs;
    const int chunks = a.chunks1 + a.chunks2;
    const int n_super = a.n_tiles * ((a.m_tiles + cs - 1) / cs);
    const uint16_t cmask = (uint16_t)((1u << cs) - 1u);

    // programmatic dependent launch: the next kernel of the stream may start its prologue on SMs this grid has left
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x1);
        if (a.chunks2) tma_prefetch_desc(&tm_x2);
        tma_prefetch_desc(&tm_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < a.a_stages; ++s) {
            mbar_init(bar_fa + 8u * s, 1);
            mbar_init(bar_ea + 8u * s, (uint32_t)a.issuers);      // every MMA issuer commits
        }
        for (int s = 0; s < a.b_stages; ++s) {
            mbar_init(bar_fb + 8u * s, 1);
            mbar_init(bar_eb + 8u * s, (uint32_t)cs);     // every CTA of the cluster must have consumed the slot
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8u * i, (uint32_t)a.issuers);   // every MMA issuer commits
            mbar_init(bar_tempty + 8u * i, kEpiWarps);
        }
        mbar_fence_init();
    }
    if (warp == 2) tc_alloc(slot, a.tmem_cols);
    tc_fence_before();
    if (cs > 1) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
    // ... and this grid touches global memory only after the previous kernel has completed (no-op without the launch attribute)
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // The three single-issuer roles run their loops WARP-WIDE (all 32 lanes carry identical values) and predicate only
    // the issuing instruction with elect.sync: inside an `if (lane == 0)` region ptxas cannot keep descriptors in
    // uniform registers and wraps every UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST waterfall loop (~100+ cycles
    // each, measured: the tensor pipe ran at ~55% of its floor because of it).
    if (MIXED && warp == 0 && a.issuers == 3) {
        // ===== A + B producer in ONE warp (MIXED kernels with three issuers: warp 3 issues MMAs instead).  Loads are requested in
        // consumption order -- the activation box of a (chunk, U shift, V group), then the weight stages of its K blocks -- so a wait
        // for a free weight slot can only be behind MMAs whose operands have already been requested.
        const int bnp = a.bn / cs;
        const int ctot = chunks * BK;
        uint32_t sa_i = 0, pha = 0, sb_i = 0, phb = 0;
        long long w_ea = 0, w_eb = 0;
        for (int st = cid; st < n_super; st += ncl) {
            const TileCoord t = tile_coord(a, st, crank);
            const int row0 = t.nt * a.bn + crank * bnp;
            int su0, su1;
            su_range(a, t.nt, su0, su1);
            for (int ch = 0; ch < chunks; ++ch) {
                const bool first = ch < a.chunks1;
                const CUtensorMap* m = first ? &tm_x1 : &tm_x2;
                const int c0 = (first ? ch : ch - a.chunks1) * BK;
                for (int su = su0; su < su1; ++su)
                    for (int g = 0; g < a.n_groups; ++g) {
                        long long t0 = DBG ? clock64() : 0;
                        mbar_wait(bar_ea + 8u * sa_i, pha ^ 1u);
                        if (DBG) w_ea += clock64() - t0;
                        const int iu0 = t.ou0 * a.su - a.pad_u + su;
                        const int iv0 = t.ov0 * a.sv - a.pad_v + a.g_tap0[g];
                        const uint32_t sa = base + sa_i * a_stage;
                        if (elect_one()) {
                            if (DBG && (a.exp & 2)) {
                                mbar_arrive(bar_fa + 8u * sa_i);
                            } else {
                                mbar_expect_tx(bar_fa + 8u * sa_i, a_stage);
                                tma_load_5d(sa, m, bar_fa + 8u * sa_i, c0, iu0, iv0, t.img, 0);
                                tma_load_5d(sa + a_plane, m, bar_fa + 8u * sa_i, c0, iu0, iv0, t.img, 1);
                            }
                        }
                        __syncwarp();
                        if (++sa_i == (uint32_t)a.a_stages) { sa_i = 0; pha ^= 1u; }
                        int tv0, tv1;
                        tv_range(a, t.nt, a.g_ntaps[g], tv0, tv1);
                        for (int j0 = tv0; j0 < tv1; j0 += a.tpb) {
                            const int ntb = min(a.tpb, tv1 - j0);
                            t0 = DBG ? clock64() : 0;
                            mbar_wait(bar_eb + 8u * sb_i, phb ^ 1u);
                            if (DBG) w_eb += clock64() - t0;
                            const uint32_t sb = smem_b + sb_i * b_stage + (uint32_t)(crank * bnp) * ROW_BYTES;
                            if (DBG && (a.exp & 1)) {
                                if (elect_one()) mbar_arrive(bar_fb + 8u * sb_i);
                            } else if (elect_one()) {
                                mbar_expect_tx(bar_fb + 8u * sb_i, (uint32_t)ntb * b_tap);
                                for (int jj = 0; jj < ntb; ++jj) {
                                    const int tv = a.g_tap0[g] + (j0 + jj) * a.g_step;
                                    const int tap = a.ux ? tv * a.kw + su : su * a.kw + tv;
                                    const int k0 = tap * ctot + ch * BK;
                                    const uint32_t dst = sb + (uint32_t)jj * b_tap;
                                    if (cs > 1) {
                                        tma_load_3d_mc(dst, &tm_w, bar_fb + 8u * sb_i, k0, row0, 0, cmask);
                                        tma_load_3d_mc(dst + b_plane, &tm_w, bar_fb + 8u * sb_i, k0, row0, 1, cmask);
                                    } else {
                                        tma_load_3d(dst, &tm_w, bar_fb + 8u * sb_i, k0, row0, 0);
                                        tma_load_3d(dst + b_plane, &tm_w, bar_fb + 8u * sb_i, k0, row0, 1);
                                    }
                                }
                            }
                            __syncwarp();
                            if (++sb_i == (uint32_t)a.b_stages) { sb_i = 0; phb ^= 1u; }
                        }
                    }
            }
        }
        if (DBG && lane == 0) { a.dbg[blockIdx.x * 12 + 4] = (unsigned long long)w_ea; a.dbg[blockIdx.x * 12 + 5] = (unsigned long long)w_eb; }
        if (cs > 1) {
            for (int i = 0; i < a.b_stages; ++i) {
                mbar_wait(bar_eb + 8u * sb_i, phb ^ 1u);
                if (++sb_i == (uint32_t)a.b_stages) { sb_i = 0; phb ^= 1u; }
            }
        }
    } else if (warp == 0) {
        // ===== A producer: one halo box (both planes) per (channel chunk, U shift, V-tap group)
        uint32_t s = 0, ph = 0;
        long long w_ea = 0;
        for (int st = cid; st < n_super; st += ncl) {
            const TileCoord t = tile_coord(a, st, crank);
            int su0, su1;
            su_range(a, t.nt, su0, su1);
            for (int ch = 0; ch < chunks; ++ch) {
                const bool first = ch < a.chunks1;
                const CUtensorMap* m = first ? &tm_x1 : &tm_x2;
                const int c0 = (first ? ch : ch - a.chunks1) * BK;
                for (int su = su0; su < su1; ++su)
                    for (int g = 0; g < a.n_groups; ++g) {
                        const long long t0 = DBG ? clock64() : 0;
                        mbar_wait(bar_ea + 8u * s, ph ^ 1u);
                        if (DBG) w_ea += clock64() - t0;
                        const int iu0 = t.ou0 * a.su - a.pad_u + su;
                        const int iv0 = t.ov0 * a.sv - a.pad_v + a.g_tap0[g];
                        const uint32_t sa = base + s * a_stage;
                        if (elect_one()) {
                            if (DBG && (a.exp & 2)) {
                                mbar_arrive(bar_fa + 8u * s);
                            } else {
                                mbar_expect_tx(bar_fa + 8u * s, a_stage);
                                tma_load_5d(sa, m, bar_fa + 8u * s, c0, iu0, iv0, t.img, 0);
                                tma_load_5d(sa + a_plane, m, bar_fa + 8u * s, c0, iu0, iv0, t.img, 1);
                            }
                        }
                        __syncwarp();
                        if (++s == (uint32_t)a.a_stages) { s = 0; ph ^= 1u; }
                    }
            }
        }
        if (DBG && lane == 0) a.dbg[blockIdx.x * 12 + 4] = (unsigned long long)w_ea;
    } else if (warp == 3 && !(MIXED && a.issuers == 3)) {
        // ===== B producer: one K block (tpb taps x BK channels of the weight tile) per stage; with a cluster every CTA loads
        // 1/cs of the rows and multicasts them to all CTAs (same smem offset, same barrier offset everywhere)
        const int bnp = a.bn / cs;
        const int ctot = chunks * BK;
        uint32_t s = 0, ph = 0;
        long long w_eb = 0;
        for (int st = cid; st < n_super; st += ncl) {
            const TileCoord t = tile_coord(a, st, crank);
            const int row0 = t.nt * a.bn + crank * bnp;
            int su0, su1;
            su_range(a, t.nt, su0, su1);
            for (int ch = 0; ch < chunks; ++ch)
                for (int su = su0; su < su1; ++su)
                    for (int g = 0; g < a.n_groups; ++g) {
                        int tv0, tv1;
                        tv_range(a, t.nt, a.g_ntaps[g], tv0, tv1);
                        for (int j0 = tv0; j0 < tv1; j0 += a.tpb) {
                            const int ntb = min(a.tpb, tv1 - j0);       // taps in this K block
                            const long long t0 = DBG ? clock64() : 0;
                            mbar_wait(bar_eb + 8u * s, ph ^ 1u);
                            if (DBG) w_eb += clock64() - t0;
                            const uint32_t sb = smem_b + s * b_stage + (uint32_t)(crank * bnp) * ROW_BYTES;
                            if (DBG && (a.exp & 1)) {
                                if (elect_one()) mbar_arrive(bar_fb + 8u * s);
                            } else if (elect_one()) {
                                mbar_expect_tx(bar_fb + 8u * s, (uint32_t)ntb * b_tap);
                                for (int jj = 0; jj < ntb; ++jj) {
                                    const int tv = a.g_tap0[g] + (j0 + jj) * a.g_step;
                                    const int tap = a.ux ? tv * a.kw + su : su * a.kw + tv;
                                    const int k0 = tap * ctot + ch * BK;
                                    const uint32_t dst = sb + (uint32_t)jj * b_tap;
                                    if (cs > 1) {
                                        tma_load_3d_mc(dst, &tm_w, bar_fb + 8u * s, k0, row0, 0, cmask);
                                        tma_load_3d_mc(dst + b_plane, &tm_w, bar_fb + 8u * s, k0, row0, 1, cmask);
                                    } else {
                                        tma_load_3d(dst, &tm_w, bar_fb + 8u * s, k0, row0, 0);
                                        tma_load_3d(dst + b_plane, &tm_w, bar_fb + 8u * s, k0, row0, 1);
                                    }
                                }
                            }
                            __syncwarp();
                            if (++s == (uint32_t)a.b_stages) { s = 0; ph ^= 1u; }
                        }
                    }
        }
        if (DBG && lane == 0) a.dbg[blockIdx.x * 12 + 5] = (unsigned long long)w_eb;
        // tail: every arrive the peers send to this CTA's empty barriers must land before the CTA exits
        if (cs > 1) {
            for (int i = 0; i < a.b_stages; ++i) {       // = the waits of b_stages more (virtual) uses: the previous use of
                mbar_wait(bar_eb + 8u * s, ph ^ 1u);     //   every slot has been consumed by every CTA of the cluster
                if (++s == (uint32_t)a.b_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1 || (warp == 2 && a.issuers >= 2) || (MIXED && warp == 3 && a.issuers == 3)) {
        // ===== MMA issuers (one elected thread per warp).  Per 16-deep K slice two MMAs: B_hi and B_lo are adjacent in
        // the stage, so
        //   D[:, 0:2bn]  (+)= A_hi * [B_hi; B_lo]^T      (N = 2*bn: hi*hi | hi*lo)
        //   D[:, 0:bn]    += A_lo *  B_hi^T
        // (A is read from shared memory twice instead of three times; the epilogue adds the two halves.)
        //
        // Why two issuers: tcgen05.mma issue is back-pressured -- the issuing thread runs at most ~1-2 MMAs ahead of the
        // tensor pipe -- so everything a single issuer does between two K blocks (commit, mbarrier poll ~100 cycles even
        // when complete, descriptors, elect, R2UR) is a pipe bubble: measured 283 / 198 / 178 cycles per K slice at
        // bn = 128 / 64 / 32 against a pipe floor of 192 / 112 / 94 -- with all loads and stores disabled, i.e. not a
        // memory effect (EVK_TC_EXP; tools/microbench/mma_loop_bench.cu).  The K blocks of a tile are dealt alternately
        // to two issuer warps that accumulate into the SAME tensor-memory accumulator in STRICT block order: issuer r
        // prepares block b (barrier wait, fence, descriptors), then waits on named barrier 1+r for the other issuer's
        // "block b-1 issued" signal, issues its eight MMAs and signals barrier 2-r.  Only the ~30-cycle handshake sits
        // between two blocks (192 / 143 / 132 cycles per slice in the microbenchmark), and because the pipe executes
        // in issue order the fp32 summation order -- hence every output bit -- is the same from run to run.
        // Small N tiles (bn <= 64), whose MMAs are too short to cover even the handshake, run the two issuers FREE
        // (no handshake) into one accumulator EACH, summed in fixed order by the epilogue -- equally deterministic,
        // 134 / 122 cycles per slice; bn = 128 cannot afford two accumulators next to the epilogue's double buffer.
        // Free-running issuers deal the blocks by the GLOBAL block count (continuous across tiles) over an EVEN number
        // of weight stages, so that a ring slot is always consumed by the same issuer: mbarrier waits are parity
        // based, and an issuer that runs ahead must never wait for use n of a slot whose use n-1 the OTHER issuer has
        // not seen complete yet (the parity test would pass on the stale phase).
        // Barriers that guard data read by both issuers' MMAs (A stage free, accumulator complete) count two commits.
        const uint32_t role = (uint32_t)(warp - 1);
        // MIXED (see "mixed operands" at the top of the host section): plane 0 = fp16, plane 1 = fp8 pairs; per 32-byte K step
        //   D[:, 0:bn] (+)= A16 * B16^T  (kind::f16, K = 16)   and   D[:, 0:bn] += [x8 | xl8] * [wl8 | w8]^T  (kind::f8f6f4, K = 32)
        const uint32_t idesc2 = MIXED ? umma_idesc_f16(128, (uint32_t)a.bn) : umma_idesc_bf16(128, (uint32_t)(2 * a.bn));
        const uint32_t idesc1 = MIXED ? umma_idesc_f8_e5m2_e4m3(128, (uint32_t)a.bn) : umma_idesc_bf16(128, (uint32_t)a.bn);
        const uint32_t b2 = MIXED ? (b_plane >> 4) : 0u;      // second MMA's weight operand: plane 1 (MIXED) or B_hi again
        const uint32_t desc_hi = (uint32_t)(umma_desc_kmajor(0, ROW_BYTES) >> 32);
        auto mk = [&](uint32_t lo) -> uint64_t { return ((uint64_t)desc_hi << 32) | lo; };
        auto lo_of = [](uint32_t addr) -> uint32_t { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); };
        const uint32_t atom16 = ATOM >> 4, a_plane16 = a_plane >> 4, btap16 = b_tap >> 4;
        const uint32_t nbs = (uint32_t)a.b_stages;
        const uint32_t two = a.issuers >= 2 ? 1u : 0u;
        const uint32_t ni = (uint32_t)a.issuers;            // (three only in MIXED kernels: strict order, round-robin handshake)
        const bool own = a.own_acc != 0;
        const uint32_t dsh = (own && a.deal == 2) ? 1u : 0u;
        const uint32_t kb_total = a.ps == 3 ? (uint32_t)(chunks * (a.ku - 1) * (a.kv - 1)) : (uint32_t)(chunks * (a.ps == 2 ? a.ku - 1 : a.ku) * ((a.g_ntaps[0] + a.tpb - 1) / a.tpb + (a.n_groups > 1 ? (a.g_ntaps[1] + a.tpb - 1) / a.tpb : 0)));
        uint32_t sA = 0, phA = 0, sB = 0, phB = 0, it = 0, gblk = 0;
        bool b_ready = false;
        long long w_te = 0, w_fa = 0, w_fb = 0, w_try = 0, w_issue = 0, w_body = 0, w_grp = 0, w_else = 0;
        const long long t_begin = DBG ? clock64() : 0;
        if (MIXED && !DBG && a.lean) {
            // LEAN issuer loop (MIXED, one tap per K block, two free-running issuers on their own accumulators): the same barrier
            // protocol as the general loop below with everything that is constant per launch hoisted, no mode branches and no
            // look-ahead poll.  ncu's instruction-level sampling of the general loop (profiles/r02b_issuer_stalls.md) showed the
            // issuer warps blocked on the tensor pipe only a quarter of the time: ~157 mostly serial instructions per K block
            // (kernel-parameter reloads, mode branches) cost ~1 450 cycles next to 512 cycles of MMAs.  (The same loop generalised
            // to the bf16x3 kernels -- in the same kernel, or as a separate instantiation with only this loop -- cost them registers:
            // 56-88 bytes of spills, FireNet 8-13 % slower, whose layers are bound by their epilogues: kept to MIXED.  The K block
            // as ONE asm statement with four 64-bit descriptors advanced in place changed nothing.)
            const uint32_t a_stage16 = a_stage >> 4, b_stage16 = b_stage >> 4;
            const uint32_t a_lo_base = lo_of(base), b_lo_base = lo_of(smem_b);
            const uint32_t nas = (uint32_t)a.a_stages;
            const int n_groups = a.n_groups, gn0 = a.g_ntaps[0], gn1 = a.g_ntaps[1], n_tiles = a.n_tiles;
            const uint32_t acc_stride = (uint32_t)a.acc_stride, d_role = tmem_base + role * (uint32_t)a.acc_cols;
            for (int st = cid; st < n_super; st += ncl, ++it) {
                const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
                mbar_wait(bar_tempty + 8u * as, aph ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = d_role + as * acc_stride;
                uint32_t accf = 0u;               // 0 until this issuer's first MMA of the tile
                const int nt_tile = st % n_tiles;
                int su0, su1;
                su_range(a, nt_tile, su0, su1);
                for (int ch = 0; ch < chunks; ++ch)
                    for (int su = su0; su < su1; ++su)
                        for (int g = 0; g < n_groups; ++g) {
                            int tv0, tv1;
                            tv_range(a, nt_tile, g ? gn1 : gn0, tv0, tv1);
                            mbar_wait(bar_fa + 8u * sA, phA);
                            uint32_t al = a_lo_base + sA * a_stage16 + (uint32_t)tv0 * atom16;
                            for (int j = tv0; j < tv1; ++j, ++gblk, al += atom16) {
                                if ((gblk & 1u) == role) {
                                    mbar_wait(bar_fb + 8u * sB, phB);
                                    tc_fence_after();
                                    const uint32_t bl = b_lo_base + sB * b_stage16;
                                    if (elect_one()) {
#pragma unroll
                                        for (int k = 0; k < BK / 16; ++k) {
                                            tc_mma_bf16(d_tmem, mk(al + 2 * k), mk(bl + 2 * k), idesc2, k == 0 ? accf : 1u);
                                            tc_mma_f8(d_tmem, mk(al + a_plane16 + 2 * k), mk(bl + b2 + 2 * k), idesc1, 1u);
                                        }
                                        if (cs > 1) tc_commit_mc(bar_eb + 8u * sB, cmask); else tc_commit(bar_eb + 8u * sB);
                                    }
                                    __syncwarp();
                                    accf = 1u;
                                }
                                if (++sB == nbs) { sB = 0; phB ^= 1u; }
                            }
                            if (elect_one()) tc_commit(bar_ea + 8u * sA);
                            __syncwarp();
                            if (++sA == nas) { sA = 0; phA ^= 1u; }
                        }
                if (elect_one()) tc_commit(bar_tfull + 8u * as);
                __syncwarp();
            }
        } else
        for (int st = cid; st < n_super; st += ncl, ++it) {
            const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
            long long t0 = DBG ? clock64() : 0;
            mbar_wait(bar_tempty + 8u * as, aph ^ 1u);          // epilogue drained this accumulator stage
            if (DBG) w_te += clock64() - t0;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * (uint32_t)a.acc_stride + (own ? role * (uint32_t)a.acc_cols : 0u);
            uint32_t blk = 0;
            bool fresh = true;                // the next MMA of this issuer initialises its accumulator
            if (!own) b_ready = false;        // strict mode deals blocks per tile: the look-ahead distance differs across tiles
            int su0, su1;
            const int nt_tile = st % a.n_tiles;
            su_range(a, nt_tile, su0, su1);
            for (int ch = 0; ch < chunks; ++ch)
                for (int su = su0; su < su1; ++su)
                    for (int g = 0; g < a.n_groups; ++g) {
                        const int nt_g = a.g_ntaps[g];
                        uint32_t ah_lo = lo_of(base + sA * a_stage);
                        t0 = DBG ? clock64() : 0;
                        mbar_wait(bar_fa + 8u * sA, phA);
                        if (DBG) w_fa += clock64() - t0;
                        int tv0, tv1;
                        tv_range(a, nt_tile, nt_g, tv0, tv1);
                        ah_lo += (uint32_t)tv0 * atom16;
                        for (int j0 = tv0; j0 < tv1; j0 += a.tpb, ++blk, ++gblk) {
                            const int ntb = min(a.tpb, tv1 - j0);               // taps in this K block
                            const long long t_it = DBG ? clock64() : 0;
                            // free-running issuers may take the blocks in PAIRS (a.deal == 2: blocks 4i, 4i+1 / 4i+2, 4i+3): twice the MMAs per
                            // issue phase for the same barrier round trips (a multiple of four weight stages keeps a slot with one issuer)
                            const uint32_t turn = own ? (gblk >> dsh) : blk;
                            const bool mine_dbg = two == 0u || (ni == 3u ? turn % 3u : (turn & 1u)) == role;
                            if (mine_dbg) {
                                const long long t_body = DBG ? clock64() : 0;
                                const uint32_t bh_lo = lo_of(smem_b + sB * b_stage);
                                const uint32_t bar_free = bar_eb + 8u * sB;
                                if (!b_ready) {
                                    t0 = DBG ? clock64() : 0;
                                    mbar_wait(bar_fb + 8u * sB, phB);
                                    if (DBG) w_fb += clock64() - t0;
                                }
                                t0 = DBG ? clock64() : 0;
                                tc_fence_after();
                                if (DBG) w_te += clock64() - t0;        // (DBG: the "tempty" counter also carries the per-block fence)
                                // this issuer's NEXT block: poll its barrier now, the round trip hides behind the issue
                                uint32_t s2 = sB + ni, ph2 = phB;
                                if (dsh) s2 = sB + ((gblk & 1u) ? 3u : 1u);
                                if (s2 >= nbs) { s2 -= nbs; ph2 ^= 1u; }
                                t0 = DBG ? clock64() : 0;
                                if (a.poll == 0) b_ready = mbar_try_wait(bar_fb + 8u * s2, ph2);
                                else b_ready = false;
                                if (DBG) w_try += clock64() - t0;
                                if (two && !own && blk > 0) {     // my turn: the previous issuer has issued block blk - 1
                                    if (role == 0u) asm volatile("bar.sync 1, 64;" ::: "memory");
                                    else if (role == 1u) asm volatile("bar.sync 2, 64;" ::: "memory");
                                    else asm volatile("bar.sync 3, 64;" ::: "memory");
                                }
                                t0 = DBG ? clock64() : 0;
                                if (elect_one()) {
                                    const uint32_t acc0 = (own ? fresh : blk == 0u) ? 0u : 1u;
                                    if (a.tpb == 1) {             // (kept separate: no loop-carried state on the hot single-tap path)
#pragma unroll
                                        for (int k = 0; k < BK / 16; ++k) {
                                            // +16 elements (32 B) along K inside the swizzle atom = +2 in the 16-byte address field
                                            if (!DBG || !(a.exp & 8)) tc_mma_bf16(d_tmem, mk(ah_lo + 2 * k), mk(bh_lo + 2 * k), idesc2, k == 0 ? acc0 : 1u);
                                            if (DBG && (a.exp & 16)) continue;
                                            if (MIXED) tc_mma_f8(d_tmem, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + b2 + 2 * k), idesc1, 1u);
                                            else tc_mma_bf16(d_tmem, mk(ah_lo + a_plane16 + 2 * k), mk(bh_lo + 2 * k), idesc1, 1u);
                                        }
                                    } else {
                                        uint32_t al = ah_lo, bl = bh_lo, acc = acc0;
                                        for (int jj = 0; jj < ntb; ++jj, al += atom16, bl += btap16) {
#pragma unroll
                                            for (int k = 0; k < BK / 16; ++k) {
                                                tc_mma_bf16(d_tmem, mk(al + 2 * k), mk(bl + 2 * k), idesc2, acc);
                                                acc = 1u;
                                                if (MIXED) tc_mma_f8(d_tmem, mk(al + a_plane16 + 2 * k), mk(bl + b2 + 2 * k), idesc1, 1u);
                                                else tc_mma_bf16(d_tmem, mk(al + a_plane16 + 2 * k), mk(bl + 2 * k), idesc1, 1u);
                                            }
                                        }
                                    }
                                    // frees the weight slot in every CTA of the cluster when these MMAs retire
                                    if (cs > 1) tc_commit_mc(bar_free, cmask); else tc_commit(bar_free);
                                }
                                __syncwarp();
                                if (a.poll == 2) b_ready = mbar_test_wait(bar_fb + 8u * s2, ph2);      // (after the issue: never delays it)
                                if (DBG) w_issue += clock64() - t0;
                                fresh = false;
                                if (two && !own && blk + 1 < kb_total) {  // hand the turn to the next issuer (role + 1 mod ni waits on barrier 1 + that role)
                                    const uint32_t nxt = role + 1u == ni ? 0u : role + 1u;
                                    if (nxt == 0u) asm volatile("bar.arrive 1, 64;" ::: "memory");
                                    else if (nxt == 1u) asm volatile("bar.arrive 2, 64;" ::: "memory");
                                    else asm volatile("bar.arrive 3, 64;" ::: "memory");
                                }
                                if (DBG) w_body += clock64() - t_body;
                            }
                            if (++sB == nbs) { sB = 0; phB ^= 1u; }
                            ah_lo += (uint32_t)ntb * atom16;
                            if (DBG && !mine_dbg) w_else += clock64() - t_it;
                        }
                        t0 = DBG ? clock64() : 0;
                        if (elect_one()) tc_commit(bar_ea + 8u * sA);       // `issuers` arrivals free the activation stage
                        __syncwarp();
                        if (DBG) w_grp += clock64() - t0;
                        if (++sA == (uint32_t)a.a_stages) { sA = 0; phA ^= 1u; }
                    }
            if (elect_one()) tc_commit(bar_tfull + 8u * as);         // `issuers` arrivals: accumulator complete
            __syncwarp();
        }
        if (DBG && lane == 0 && role == 0) {
            a.dbg[blockIdx.x * 12 + 0] = (unsigned long long)(clock64() - t_begin);
            a.dbg[blockIdx.x * 12 + 1] = (unsigned long long)w_te;
            a.dbg[blockIdx.x * 12 + 2] = (unsigned long long)w_fa;
            a.dbg[blockIdx.x * 12 + 3] = (unsigned long long)w_fb;
            a.dbg[blockIdx.x * 12 + 10] = (unsigned long long)(a.exp & 128 ? w_else : a.exp & 64 ? w_grp : a.exp & 32 ? w_body : w_try);
            a.dbg[blockIdx.x * 12 + 11] = (unsigned long long)w_issue;
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warp (4 + e) owns TMEM lane quadrant e % 4 (its hardware-accessible lanes) and
        // the 32-column chunks with parity e / 4.  Thread = accumulator row = output pixel.
        const int e = warp - 4;
        const int wq = e & 3, half = e >> 2;
        const int row = wq * 32 + lane;
        const int lu = row & 7, lv = row >> 3;
        uint32_t it = 0;
        long long w_tf = 0, w_ld = 0, w_rest = 0;
        const long long e_begin = DBG ? clock64() : 0;
        for (int st = cid; st < n_super; st += ncl, ++it) {
            const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
            const TileCoord t = tile_coord(a, st, crank);
            const int oy = a.ux ? t.ov0 + lv : t.ou0 + lu;
            const int ox = a.ux ? t.ou0 + lu : t.ov0 + lv;
            const int n0 = t.nt * a.bn;
            const bool valid = !t.dummy && oy < a.Hout && ox < a.Wout;
            const size_t pix = ((size_t)t.img * a.Hout + oy) * a.Wout + ox;
            // split outputs in a row-padded layout: offset of this GEMM row's record relative to its dense offset pix * s_grp
            const long long sd = a.s_pitch ? (long long)((size_t)t.img * a.Hout + oy) * (a.s_pitch - a.Wout * a.s_grp) + a.s_off : 0;
            const int cw = a.cw;
            const int nchunks = (a.bn + cw - 1) / cw;
            const long long t0 = DBG ? clock64() : 0;
            mbar_wait(bar_tfull + 8u * as, aph);
            if (DBG) w_tf += clock64() - t0;
            tc_fence_after();
            const uint32_t t_row = tmem_base + as * (uint32_t)a.acc_stride + ((uint32_t)(wq * 32) << 16);
            for (int c = half; c < nchunks; c += 2) {
                const int j0 = c * cw;
                uint32_t v[32];
                __syncwarp();                     // tcgen05.ld is .sync.aligned: reconverge after the masked stores
                const long long t_c0 = DBG ? clock64() : 0;
                {
                    // bf16x3: columns [0, bn) hold hi*hi + lo*hi, columns [bn, 2bn) hi*lo; MIXED: one accumulator of bn columns
                    uint32_t u[32];
                    if (cw == 32) {
                        tc_ld_32x32(t_row + (uint32_t)j0, v);
                        if (!MIXED) tc_ld_32x32(t_row + (uint32_t)(a.bn + j0), u);
                    } else if (cw == 16) {
#pragma unroll
                        for (int i = 16; i < 32; ++i) { v[i] = 0u; u[i] = 0u; }
                        tc_ld_32x16(t_row + (uint32_t)j0, v);
                        if (!MIXED) tc_ld_32x16(t_row + (uint32_t)(a.bn + j0), u);
                    } else {
#pragma unroll
                        for (int i = 8; i < 32; ++i) { v[i] = 0u; u[i] = 0u; }
                        tc_ld_32x8(t_row + (uint32_t)j0, v);
                        if (!MIXED) tc_ld_32x8(t_row + (uint32_t)(a.bn + j0), u);
                    }
                    if (MIXED && a.own_acc) {         // second issuer's accumulator
                        if (cw == 32) tc_ld_32x32(t_row + (uint32_t)(a.acc_cols + j0), u);
                        else if (cw == 16) tc_ld_32x16(t_row + (uint32_t)(a.acc_cols + j0), u);
                        else tc_ld_32x8(t_row + (uint32_t)(a.acc_cols + j0), u);
                    }
                    tc_wait_ld();
                    if (!MIXED || a.own_acc) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(u[i]));
                    }
                    if (!MIXED && a.own_acc) {        // second issuer's accumulator (fixed order: deterministic bits)
                        uint32_t w[32];
                        if (cw == 32) {
                            tc_ld_32x32(t_row + (uint32_t)(a.acc_cols + j0), u);
                            tc_ld_32x32(t_row + (uint32_t)(a.acc_cols + a.bn + j0), w);
                        } else if (cw == 16) {
#pragma unroll
                            for (int i = 16; i < 32; ++i) w[i] = 0u;
                            tc_ld_32x16(t_row + (uint32_t)(a.acc_cols + j0), u);
                            tc_ld_32x16(t_row + (uint32_t)(a.acc_cols + a.bn + j0), w);
                        } else {
#pragma unroll
                            for (int i = 8; i < 32; ++i) w[i] = 0u;
                            tc_ld_32x8(t_row + (uint32_t)(a.acc_cols + j0), u);
                            tc_ld_32x8(t_row + (uint32_t)(a.acc_cols + a.bn + j0), w);
                        }
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            v[i] = __float_as_uint(__uint_as_float(v[i]) + (__uint_as_float(u[i]) + __uint_as_float(w[i])));
                    }
                }
                const long long t_c1 = DBG ? clock64() : 0;
                if (DBG) w_ld += t_c1 - t_c0;
                struct RestTimer { long long& acc; long long t0; bool on; __device__ ~RestTimer() { if (on) acc += clock64() - t0; } } rest_timer{w_rest, t_c1, DBG};
                if (c + 2 >= nchunks) {           // last TMEM read of this warp for this tile: release the stage
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8u * as);
                }
                const int nb = n0 + j0;
                if (!valid || nb >= a.cout || (DBG && (a.exp & 4))) continue;
                if (MIXED) {                      // the accumulator holds S[n] * (x . w): undo the per-column weight scale (a power of two)
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (g * 4 >= cw) break;
                        const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.iscale + nb + g * 4));
                        v[g * 4 + 0] = __float_as_uint(__uint_as_float(v[g * 4 + 0]) * s4.x);
                        v[g * 4 + 1] = __float_as_uint(__uint_as_float(v[g * 4 + 1]) * s4.y);
                        v[g * 4 + 2] = __float_as_uint(__uint_as_float(v[g * 4 + 2]) * s4.z);
                        v[g * 4 + 3] = __float_as_uint(__uint_as_float(v[g * 4 + 3]) * s4.w);
                    }
                }
                if (a.epi == EPI_LINEAR && a.fastlin && cw == 32) {
                    // Straight-line form of the common case (bias [+ residual] -> ReLU / none -> fp32 and / or split-bf16 stores,
                    // channel counts multiples of 32): no per-group branches, all loads issued up front.  The generic path below
                    // costs ~760 instructions per 32x32 chunk at ~9 cycles each with two epilogue warps per scheduler -- 3x the MMA
                    // time of a 20-slice tile (head), and it is the un-overlapped tail of every layer's last tile.
                    const size_t o = pix * a.creal + nb;
                    float4 b4[8], r4[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) b4[g] = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                    if (a.res != nullptr) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) r4[g] = __ldg(reinterpret_cast<const float4*>(a.res + o + g * 4));
                    } else {
#pragma unroll
                        for (int g = 0; g < 8; ++g) r4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const float fl = a.act_floor;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        v[g * 4 + 0] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 0]) + b4[g].x + r4[g].x, fl));
                        v[g * 4 + 1] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 1]) + b4[g].y + r4[g].y, fl));
                        v[g * 4 + 2] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 2]) + b4[g].z + r4[g].z, fl));
                        v[g * 4 + 3] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 3]) + b4[g].w + r4[g].w, fl));
                    }
                    if (a.y != nullptr) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) st_global_v8(a.y + o + g * 8, &v[g * 8]);
                    }
                    if (a.ys != nullptr && a.ys_mixed) {
                        store_mixed16(a.ys, a.ys_plane, o, nb, &v[0]);
                        store_mixed16(a.ys, a.ys_plane, o + 16, nb + 16, &v[16]);
                    } else if (a.ys != nullptr) {
                        uint32_t hw[16], lw[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            __nv_bfloat16 h0, l0, h1, l1;
                            split_bf16(__uint_as_float(v[2 * i]), h0, l0);
                            split_bf16(__uint_as_float(v[2 * i + 1]), h1, l1);
                            hw[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lw[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        st_global_v8(a.ys + sd + o, hw);
                        st_global_v8(a.ys + sd + o + 16, hw + 8);
                        st_global_v8(a.ys + a.ys_plane + sd + o, lw);
                        st_global_v8(a.ys + a.ys_plane + sd + o + 16, lw + 8);
                    }
                } else if (!MIXED && a.epi == EPI_LINEAR && a.fastlin && cw == 16) {
                    // 16-column variant of the straight-line linear epilogue (32-column tiles split between the two warp halves)
                    const size_t o = pix * a.creal + nb;
                    float4 b4[4], r4[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) b4[g] = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                    if (a.res != nullptr) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) r4[g] = __ldg(reinterpret_cast<const float4*>(a.res + o + g * 4));
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) r4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const float fl = a.act_floor;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        v[g * 4 + 0] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 0]) + b4[g].x + r4[g].x, fl));
                        v[g * 4 + 1] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 1]) + b4[g].y + r4[g].y, fl));
                        v[g * 4 + 2] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 2]) + b4[g].z + r4[g].z, fl));
                        v[g * 4 + 3] = __float_as_uint(fmaxf(__uint_as_float(v[g * 4 + 3]) + b4[g].w + r4[g].w, fl));
                    }
                    if (a.predw) {                  // out[pixel] = sum_c w[c] * y[pixel, c] + b  (summation order of pred_kernel)
                        float pacc = 0.f;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.pred_w + g * 4));
                            pacc = fmaf(__uint_as_float(v[g * 4 + 0]), w4.x, pacc);
                            pacc = fmaf(__uint_as_float(v[g * 4 + 1]), w4.y, pacc);
                            pacc = fmaf(__uint_as_float(v[g * 4 + 2]), w4.z, pacc);
                            pacc = fmaf(__uint_as_float(v[g * 4 + 3]), w4.w, pacc);
                        }
                        pacc += a.pred_bias;
                        a.pred_out[pix * 2 + (size_t)(nb >> 4)] = a.pred_sigmoid ? sigmoidf_(pacc) : pacc;
                    }
                    if (a.y != nullptr) { st_global_v8(a.y + o, &v[0]); st_global_v8(a.y + o + 8, &v[8]); }
                    if (a.ys != nullptr) {
                        uint32_t hi8[8], lo8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat16 h0, l0, h1, l1;
                            split_bf16(__uint_as_float(v[2 * i]), h0, l0);
                            split_bf16(__uint_as_float(v[2 * i + 1]), h1, l1);
                            hi8[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo8[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        st_global_v8(a.ys + sd + o, hi8);
                        st_global_v8(a.ys + a.ys_plane + sd + o, lo8);
                    }
                } else if (a.epi == EPI_LINEAR && a.fastps) {
                    // Straight-line form of the last decoder's epilogue (four stacked phases of 32 channels, one phase per chunk;
                    // ReLU, then the fused 1x1 prediction layer; nothing but the image is stored): all loads up front, same
                    // summation order as the generic path below.
                    const int ph = nb >> 5;
                    const int Y = 2 * oy + (ph >> 1), X = 2 * ox + (ph & 1);
                    const size_t pixl = ((size_t)t.img * a.hreal + Y) * a.wreal + X;
                    const size_t o = pixl * 32;
                    const float* rh = nullptr;
                    const float* rv = nullptr;
                    if (Y < 2 || Y >= a.hreal - 2) {          // border corrections (poly.cu), two outermost rows / columns only
                        const int side = Y < 2 ? 0 : 1, l = Y < 2 ? Y : Y - a.hreal + 4;
                        rh = a.ring_h + (((size_t)(side * a.N + t.img) * a.wreal + X) * 4 + l) * 32;
                    }
                    if (X < 2 || X >= a.wreal - 2) {
                        const int side = X < 2 ? 0 : 1, l = X < 2 ? X : X - a.wreal + 4;
                        rv = a.ring_v + (((size_t)(side * a.N + t.img) * a.hreal + Y) * 4 + l) * 32;
                    }
                    float pacc = 0.f;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {          // 16 channels at a time (register budget)
                        float4 b4[4], w4[4], s4[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            b4[g] = __ldg(reinterpret_cast<const float4*>(a.bias + hf * 16 + g * 4));
                            w4[g] = __ldg(reinterpret_cast<const float4*>(a.pred_w + hf * 16 + g * 4));
                        }
                        if (a.pred_skip_s != nullptr) {          // hi + lo planes of the skip tensor, 8 channels per 16-byte load
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(a.pred_skip_s + o + hf * 16 + q * 8));
                                const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(a.pred_skip_s + a.pred_skip_plane + o + hf * 16 + q * 8));
                                const uint32_t hh[4] = {h4.x, h4.y, h4.z, h4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w};
                                float e[8];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    e[2 * i] = __uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16);
                                    e[2 * i + 1] = __uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u);
                                }
                                s4[2 * q] = make_float4(e[0], e[1], e[2], e[3]);
                                s4[2 * q + 1] = make_float4(e[4], e[5], e[6], e[7]);
                            }
                        } else if (a.pred_skip != nullptr) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) s4[g] = __ldg(reinterpret_cast<const float4*>(a.pred_skip + o + hf * 16 + g * 4));
                        } else {
#pragma unroll
                            for (int g = 0; g < 4; ++g) s4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        if (rh != nullptr) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rh + hf * 16 + g * 4));
                                b4[g].x += r4.x; b4[g].y += r4.y; b4[g].z += r4.z; b4[g].w += r4.w;
                            }
                        }
                        if (rv != nullptr) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rv + hf * 16 + g * 4));
                                b4[g].x += r4.x; b4[g].y += r4.y; b4[g].z += r4.z; b4[g].w += r4.w;
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int c = hf * 16 + g * 4;
                            const float f0 = fmaxf(__uint_as_float(v[c + 0]) + b4[g].x, 0.f), f1 = fmaxf(__uint_as_float(v[c + 1]) + b4[g].y, 0.f);
                            const float f2 = fmaxf(__uint_as_float(v[c + 2]) + b4[g].z, 0.f), f3 = fmaxf(__uint_as_float(v[c + 3]) + b4[g].w, 0.f);
                            pacc = fmaf(f0 + s4[g].x, w4[g].x, pacc);
                            pacc = fmaf(f1 + s4[g].y, w4[g].y, pacc);
                            pacc = fmaf(f2 + s4[g].z, w4[g].z, pacc);
                            pacc = fmaf(f3 + s4[g].w, w4[g].w, pacc);
                        }
                    }
                    pacc += a.pred_bias;
                    a.pred_out[pixl] = a.pred_sigmoid ? sigmoidf_(pacc) : pacc;
                } else if (a.epi == EPI_LINEAR) {
                    // row-pair mode: columns [0, C) are output row 2*oy, columns [C, 2C) row 2*oy + 1
                    // phase-stacked mode: column block ph = a*2+b of GEMM row (oy, ox) is output pixel (2oy+a, 2ox+b)
                    int nbr = nb;
                    size_t pixl = pix;
                    const float* ringh = nullptr;
                    const float* ringv = nullptr;
                    if (a.rp) {
                        const int rowsel = nb >= a.creal ? 1 : 0;
                        nbr = nb - rowsel * a.creal;
                        pixl = ((size_t)t.img * a.hreal + 2 * oy + rowsel) * a.Wout + ox;
                    } else if (a.ps) {
                        const int ph = nb / a.creal;
                        nbr = nb - ph * a.creal;
                        const int Y = 2 * oy + (ph >> 1), X = 2 * ox + (ph & 1);
                        pixl = ((size_t)t.img * a.hreal + Y) * a.wreal + X;
                        if (Y < 2 || Y >= a.hreal - 2) {          // border line l = {0, 1, Ho-2, Ho-1} -> column block l
                            const int side = Y < 2 ? 0 : 1, l = Y < 2 ? Y : Y - a.hreal + 4;
                            ringh = a.ring_h + (((size_t)(side * a.N + t.img) * a.wreal + X) * 4 + l) * a.creal + nbr;
                        }
                        if (X < 2 || X >= a.wreal - 2) {
                            const int side = X < 2 ? 0 : 1, l = X < 2 ? X : X - a.wreal + 4;
                            ringv = a.ring_v + (((size_t)(side * a.N + t.img) * a.hreal + Y) * 4 + l) * a.creal + nbr;
                        }
                    }
                    const size_t o = pixl * a.creal + nbr;
                    const bool wide = a.wide != 0;     // channel counts are multiples of 16: every 8 fp32 / 16 bf16 run is one aligned sector
                    float pacc = 0.f;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (nb + g * 4 >= a.cout || j0 + g * 4 >= a.bn || g * 4 >= cw) break;
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nbr + g * 4));
                        float f[4] = {__uint_as_float(v[g * 4 + 0]) + b4.x, __uint_as_float(v[g * 4 + 1]) + b4.y,
                                      __uint_as_float(v[g * 4 + 2]) + b4.z, __uint_as_float(v[g * 4 + 3]) + b4.w};
                        if (a.res != nullptr) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.res + o + g * 4));
                            f[0] += r4.x; f[1] += r4.y; f[2] += r4.z; f[3] += r4.w;
                        }
                        if (ringh != nullptr) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(ringh + g * 4));
                            f[0] += r4.x; f[1] += r4.y; f[2] += r4.z; f[3] += r4.w;
                        }
                        if (ringv != nullptr) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(ringv + g * 4));
                            f[0] += r4.x; f[1] += r4.y; f[2] += r4.z; f[3] += r4.w;
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) f[i] = fast_act(f[i], a.act);
                        if (a.pred_out != nullptr) {      // fused 1x1 prediction layer (same summation order as pred_kernel)
                            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (a.pred_skip_s != nullptr) {          // hi + lo planes of the skip tensor
                                const uint2 h2 = __ldg(reinterpret_cast<const uint2*>(a.pred_skip_s + o + g * 4));
                                const uint2 l2 = __ldg(reinterpret_cast<const uint2*>(a.pred_skip_s + a.pred_skip_plane + o + g * 4));
                                s4.x = __uint_as_float(h2.x << 16) + __uint_as_float(l2.x << 16);
                                s4.y = __uint_as_float(h2.x & 0xffff0000u) + __uint_as_float(l2.x & 0xffff0000u);
                                s4.z = __uint_as_float(h2.y << 16) + __uint_as_float(l2.y << 16);
                                s4.w = __uint_as_float(h2.y & 0xffff0000u) + __uint_as_float(l2.y & 0xffff0000u);
                            } else if (a.pred_skip != nullptr) {
                                s4 = __ldg(reinterpret_cast<const float4*>(a.pred_skip + o + g * 4));
                            }
                            const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.pred_w + nbr + g * 4));
                            pacc = fmaf(f[0] + s4.x, w4.x, pacc);
                            pacc = fmaf(f[1] + s4.y, w4.y, pacc);
                            pacc = fmaf(f[2] + s4.z, w4.z, pacc);
                            pacc = fmaf(f[3] + s4.w, w4.w, pacc);
                        }
                        if (wide) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[g * 4 + i] = __float_as_uint(f[i]);
                            if ((g & 1) && a.y != nullptr) st_global_v8(a.y + o + (g - 1) * 4, &v[(g - 1) * 4]);
                            if ((g & 3) == 3 && a.ys != nullptr && a.ys_mixed) {
                                store_mixed16(a.ys, a.ys_plane, o + (g - 3) * 4, nbr + (g - 3) * 4, &v[(g - 3) * 4]);
                            } else if ((g & 3) == 3 && a.ys != nullptr) {
                                uint32_t hw[8], lw[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    __nv_bfloat16 h0, l0, h1, l1;
                                    split_bf16(__uint_as_float(v[(g - 3) * 4 + 2 * i]), h0, l0);
                                    split_bf16(__uint_as_float(v[(g - 3) * 4 + 2 * i + 1]), h1, l1);
                                    hw[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                                    lw[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                                }
                                st_global_v8(a.ys + sd + o + (g - 3) * 4, hw);
                                st_global_v8(a.ys + a.ys_plane + sd + o + (g - 3) * 4, lw);
                            }
                            continue;
                        }
                        if (a.y != nullptr) *reinterpret_cast<float4*>(a.y + o + g * 4) = make_float4(f[0], f[1], f[2], f[3]);
                        if (a.ys != nullptr) {
                            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_bf16(f[i], hi[i], lo[i]);
                            *reinterpret_cast<uint2*>(a.ys + sd + o + g * 4) = *reinterpret_cast<uint2*>(hi);
                            *reinterpret_cast<uint2*>(a.ys + a.ys_plane + sd + o + g * 4) = *reinterpret_cast<uint2*>(lo);
                        }
                    }
                    if (a.pred_out != nullptr) {
                        pacc += a.pred_bias;
                        a.pred_out[pixl] = a.pred_sigmoid ? sigmoidf_(pacc) : pacc;
                    }
                } else if (!MIXED && a.epi == EPI_GRU_UR && a.fastgru && cw == 32) {
                    // straight-line form (window-mode FireNet): 16 channels x {update, reset} per chunk, loads up front, SFU gates
                    const size_t o = pix * (size_t)(a.cout >> 1) + (nb >> 1);
                    float4 b4[8], hp[4];
#pragma unroll
                    for (int g = 0; g < 8; ++g) b4[g] = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
#pragma unroll
                    for (int g = 0; g < 4; ++g) hp[g] = *reinterpret_cast<const float4*>(a.h_prev + o + g * 4);
                    uint32_t uw[16], hw_[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float hh[4] = {hp[g].x, hp[g].y, hp[g].z, hp[g].w};
                        const float bb[8] = {b4[2 * g].x, b4[2 * g].y, b4[2 * g].z, b4[2 * g].w, b4[2 * g + 1].x, b4[2 * g + 1].y, b4[2 * g + 1].z, b4[2 * g + 1].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float u = fast_sigmoid(__uint_as_float(v[g * 8 + 2 * i]) + bb[2 * i]);
                            const float r = fast_sigmoid(__uint_as_float(v[g * 8 + 2 * i + 1]) + bb[2 * i + 1]);
                            uw[g * 4 + i] = __float_as_uint(u);
                            hw_[g * 4 + i] = __float_as_uint(hh[i] * r);
                        }
                    }
                    st_global_v8(a.u_out + o, uw); st_global_v8(a.u_out + o + 8, uw + 8);
                    if (a.hr_out != nullptr) { st_global_v8(a.hr_out + o, hw_); st_global_v8(a.hr_out + o + 8, hw_ + 8); }   // (fp32 copy only for a CUDA-core consumer)
                    if (a.hrs_out != nullptr) {
                        uint32_t hi8[8], lo8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat16 h0, l0, h1, l1;
                            split_bf16(__uint_as_float(hw_[2 * i]), h0, l0);
                            split_bf16(__uint_as_float(hw_[2 * i + 1]), h1, l1);
                            hi8[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo8[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        st_global_v8(a.hrs_out + sd + o, hi8);
                        st_global_v8(a.hrs_out + a.hs_plane + sd + o, lo8);
                    }
                } else if (!MIXED && a.epi == EPI_GRU_OUT && a.fastgru && cw == 16) {
                    // straight-line form: 16 channels per chunk; h' = h (1 - u) + tanh(.) u in the reference's operation order
                    const size_t o = pix * (size_t)a.cout + nb;
                    float4 b4[4], u4[4], h4[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        b4[g] = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                        u4[g] = *reinterpret_cast<const float4*>(a.u_in + o + g * 4);
                        h4[g] = *reinterpret_cast<const float4*>(a.h_prev + o + g * 4);
                    }
                    uint32_t nw[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float bb[4] = {b4[g].x, b4[g].y, b4[g].z, b4[g].w}, uu[4] = {u4[g].x, u4[g].y, u4[g].z, u4[g].w};
                        const float hh[4] = {h4[g].x, h4[g].y, h4[g].z, h4[g].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float cand = fast_tanh(__uint_as_float(v[g * 4 + i]) + bb[i]);
                            nw[g * 4 + i] = __float_as_uint(__fadd_rn(__fmul_rn(hh[i], __fsub_rn(1.0f, uu[i])), __fmul_rn(cand, uu[i])));
                        }
                    }
                    st_global_v8(a.h_new + o, nw); st_global_v8(a.h_new + o + 8, nw + 8);
                    if (a.hs_new != nullptr) {
                        uint32_t hi8[8], lo8[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat16 h0, l0, h1, l1;
                            split_bf16(__uint_as_float(nw[2 * i]), h0, l0);
                            split_bf16(__uint_as_float(nw[2 * i + 1]), h1, l1);
                            hi8[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            lo8[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        st_global_v8(a.hs_new + sd + o, hi8);
                        st_global_v8(a.hs_new + a.hs_plane + sd + o, lo8);
                    }
                } else if (!MIXED && a.epi == EPI_GRU_UR) {   // packed column = channel*2 + {update, reset} (model/submodules.py:281-282)
                    const int C = a.cout >> 1;
                    const size_t o = pix * C + (nb >> 1);
                    float hr[16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (nb + g * 8 >= a.cout || j0 + g * 8 >= a.bn || g * 8 >= cw) break;
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 8));
                        const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 8 + 4));
                        const float4 hp = *reinterpret_cast<const float4*>(a.h_prev + o + g * 4);
                        const float u0 = sigmoidf_(__uint_as_float(v[g * 8 + 0]) + b0.x), r0 = sigmoidf_(__uint_as_float(v[g * 8 + 1]) + b0.y);
                        const float u1 = sigmoidf_(__uint_as_float(v[g * 8 + 2]) + b0.z), r1 = sigmoidf_(__uint_as_float(v[g * 8 + 3]) + b0.w);
                        const float u2 = sigmoidf_(__uint_as_float(v[g * 8 + 4]) + b1.x), r2 = sigmoidf_(__uint_as_float(v[g * 8 + 5]) + b1.y);
                        const float u3 = sigmoidf_(__uint_as_float(v[g * 8 + 6]) + b1.z), r3 = sigmoidf_(__uint_as_float(v[g * 8 + 7]) + b1.w);
                        *reinterpret_cast<float4*>(a.u_out + o + g * 4) = make_float4(u0, u1, u2, u3);
                        hr[g * 4 + 0] = hp.x * r0; hr[g * 4 + 1] = hp.y * r1; hr[g * 4 + 2] = hp.z * r2; hr[g * 4 + 3] = hp.w * r3;
                        if (a.hr_out != nullptr) *reinterpret_cast<float4*>(a.hr_out + o + g * 4) = make_float4(hr[g * 4 + 0], hr[g * 4 + 1], hr[g * 4 + 2], hr[g * 4 + 3]);
                        if (a.hrs_out != nullptr) {
                            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_bf16(hr[g * 4 + i], hi[i], lo[i]);
                            *reinterpret_cast<uint2*>(a.hrs_out + sd + o + g * 4) = *reinterpret_cast<uint2*>(hi);
                            *reinterpret_cast<uint2*>(a.hrs_out + a.hs_plane + sd + o + g * 4) = *reinterpret_cast<uint2*>(lo);
                        }
                    }
                } else if (!MIXED && a.epi == EPI_GRU_OUT) {  // h' = h (1 - u) + tanh(.) u (model/submodules.py:283-285)
                    const size_t o = pix * a.cout + nb;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        if (nb + g * 4 >= a.cout || j0 + g * 4 >= a.bn || g * 4 >= cw) break;
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                        const float4 u4 = *reinterpret_cast<const float4*>(a.u_in + o + g * 4);
                        const float4 h4 = *reinterpret_cast<const float4*>(a.h_prev + o + g * 4);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, uu[4] = {u4.x, u4.y, u4.z, u4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
                        float hn[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float cand = tanhf(__uint_as_float(v[g * 4 + i]) + bb[i]);
                            hn[i] = __fadd_rn(__fmul_rn(hh[i], __fsub_rn(1.0f, uu[i])), __fmul_rn(cand, uu[i]));
                        }
                        *reinterpret_cast<float4*>(a.h_new + o + g * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                        if (a.hs_new != nullptr) {
                            __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split_bf16(hn[i], hi[i], lo[i]);
                            *reinterpret_cast<uint2*>(a.hs_new + sd + o + g * 4) = *reinterpret_cast<uint2*>(hi);
                            *reinterpret_cast<uint2*>(a.hs_new + a.hs_plane + sd + o + g * 4) = *reinterpret_cast<uint2*>(lo);
                        }
                    }
                } else {   // EPI_LSTM: packed column = channel*4 + {in, remember, out, cell}
                    const int C = a.cout >> 2;
                    const size_t o = pix * C + (nb >> 2);
                    const float4 cp0 = *reinterpret_cast<const float4*>(a.c_prev + o);
                    const float4 cp1 = *reinterpret_cast<const float4*>(a.c_prev + o + 4);
                    const float cprev[8] = {cp0.x, cp0.y, cp0.z, cp0.w, cp1.x, cp1.y, cp1.z, cp1.w};
                    float cn[8], hn[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + nb + g * 4));
                        const float ig = fast_sigmoid(__uint_as_float(v[g * 4 + 0]) + b4.x);
                        const float fg = fast_sigmoid(__uint_as_float(v[g * 4 + 1]) + b4.y);
                        const float og = fast_sigmoid(__uint_as_float(v[g * 4 + 2]) + b4.z);
                        const float cg = fast_tanh(__uint_as_float(v[g * 4 + 3]) + b4.w);
                        const float cell = __fadd_rn(__fmul_rn(fg, cprev[g]), __fmul_rn(ig, cg));
                        cn[g] = cell;
                        hn[g] = og * fast_tanh(cell);
                    }
                    *reinterpret_cast<float4*>(a.c_new + o) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(a.c_new + o + 4) = make_float4(cn[4], cn[5], cn[6], cn[7]);
                    *reinterpret_cast<float4*>(a.h_new + o) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(a.h_new + o + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                    if (a.hs_new != nullptr && a.hs_mixed) {
                        store_mixed8(a.hs_new, a.hs_plane, o, nb >> 2, hn);
                    } else if (a.hs_new != nullptr) {
                        __nv_bfloat16 hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) split_bf16(hn[i], hi[i], lo[i]);
                        *reinterpret_cast<uint4*>(a.hs_new + o) = *reinterpret_cast<uint4*>(hi);
                        *reinterpret_cast<uint4*>(a.hs_new + a.hs_plane + o) = *reinterpret_cast<uint4*>(lo);
                    }
                }
            }
            // a warp with no chunk of its parity (bn <= 32 and half == 1) still has to release the stage
            if (half >= nchunks) {
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8u * as);
            }
        }
        if (DBG && e == 0 && lane == 0) {
            a.dbg[blockIdx.x * 12 + 6] = (unsigned long long)(clock64() - e_begin);
            a.dbg[blockIdx.x * 12 + 7] = (unsigned long long)w_tf;
            a.dbg[blockIdx.x * 12 + 8] = (unsigned long long)w_ld;
            a.dbg[blockIdx.x * 12 + 9] = (unsigned long long)w_rest;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tc_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------ host side
static void* g_encode = nullptr;

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
    if (!g_encode) {
        cudaDriverEntryPointQueryResult qres;
        void* fn = nullptr;
        EVK_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        EVK_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, EVK_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        g_encode = fn;
    }
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides[i]; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    const CUresult r = ((Fn)g_encode)(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs,
                                      bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    EVK_REQUIRE(r == CUDA_SUCCESS, EVK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, box %u %u %u)", (int)r,
                rank, box[0], box[1], rank > 2 ? box[2] : 0);
    return EVK_OK;
}

static int pick_bk(const ConvParams& p) {
    if (p.c1 % 64 == 0 && p.c2 % 64 == 0) return 64;
    if (p.c1 % 32 == 0 && p.c2 % 32 == 0) return 32;
    if (p.c1 % 16 == 0 && p.c2 % 16 == 0) return 16;     // FireNet's 16-channel stack: one 16-deep slice per K block
    return 0;
}

static bool g_tc_stride2 = true;   // EVK_TC_STRIDE2=0 routes stride-2 convolutions to the fp32 SIMT kernel

bool tc_eligible(const ConvParams& p) {
    static bool env_read = false;
    if (!env_read) {
        const char* e = getenv("EVK_TC_STRIDE2");
        if (e && e[0] == '0') g_tc_stride2 = false;
        env_read = true;
    }
    if (pick_bk(p) == 0 || p.c1 == 0) return false;
    if (p.stride != 1 && !(p.stride == 2 && g_tc_stride2)) return false;
    if (p.cout % 4 != 0) return false;
    if (p.epi == EPI_LSTM && p.cout % 32 != 0) return false;
    if (p.epi == EPI_GRU_UR && p.cout % 8 != 0) return false;
    return true;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return (e && e[0]) ? atoi(e) : dflt;
}

// Cost model (cycles) of one layer for a candidate (orientation, N tile, cluster size): waves of co-resident CTAs x
// per-tile cost.  Per 16-deep K slice the kernel is bound by MMA issue: measured in-kernel cycles per
// slice as a function of the N tile (two issuers; tools/tc_experiment.py with EVK_TC_TIMING=1), next to the tile's share
// of the chip-wide L2 -> SM path (~42 B/clk/SM with every SM loading), which only binds for small tiles of wide layers.
// EVK_TC_BN / EVK_TC_CS / EVK_TC_UX override the choice (experiments).
struct TcChoice { int ux, bn, cs; double cost; };

static double slice_cycles(int bn) {
    // measured: bn 128 -> 235, 64 -> 150, 32 -> 140 (the N <= 64 MMAs cost >= 50 cycles each: the A-operand read)
    if (bn >= 128) return 235.0;
    if (bn >= 64) return 150.0 + (bn - 64) * (235.0 - 150.0) / 64.0;
    return 130.0 + bn * (150.0 - 130.0) / 64.0;
}

// MIXED kernels: two MMAs of N = bn per slice, bound by the operand fetch from shared memory (A 4 kB + B bn * 32 B per MMA at
// ~85 B/clk: 190 cycles at bn = 128, 285 at bn = 256 -- per output column the wide tile is a quarter cheaper)
static double slice_cycles_mixed(int bn) { return 2.0 * (4096.0 + 32.0 * bn) / 85.0 + (bn >= 256 ? 0.0 : 0.0); }

static double tile_cost(int bk, int ku, int kv, int stride, int chunks, int bn, int cs, long ctas, int* ar_out, bool mixed = false) {
    const int ngroups = stride == 2 ? 2 : 1;
    const int max_taps = stride == 2 ? (kv + 1) / 2 : kv;
    const int ar = 16 + max_taps - 1;
    if (ar_out) *ar_out = ar;
    const double k16 = (double)chunks * ku * kv * (bk / 16);
    const double t_mma = k16 * (mixed ? slice_cycles_mixed(bn) : slice_cycles(bn));
    const double a_bytes = (double)chunks * ku * ngroups * 2.0 * ar * 8 * bk * 2;
    const double b_bytes = (double)chunks * ku * kv * 2.0 * bn * bk * 2 / cs;
    const double active = (double)std::min<long>(ctas, kNumSMs);
    const double t_l2 = (a_bytes + b_bytes) * active / 6300.0;                // chip-wide ceiling shared by the active SMs
    const double t_epi = 60.0 * bn;                                           // epilogue of the previous tile (overlapped)
    return 5000.0 + std::max(std::max(t_mma, t_l2), t_epi);
}

void pack_weights_tc(const float* w_kc, int K, int cout, int cout_pad, std::vector<__nv_bfloat16>& out) {
    out.assign((size_t)2 * cout_pad * K, __float2bfloat16(0.0f));
    __nv_bfloat16* o = out.data();
    parallel_channels(cout, [=](int n0, int n1) {
        for (int nb = n0; nb < n1; nb += 16)                  // 16 columns at a time: one cache line of the source per k
            for (int k = 0; k < K; ++k)
                for (int n = nb; n < std::min(nb + 16, n1); ++n) {
                    const float v = w_kc[(size_t)k * cout + n];
                    const __nv_bfloat16 hi = __float2bfloat16(v);
                    o[(size_t)n * K + k] = hi;
                    o[(size_t)cout_pad * K + (size_t)n * K + k] = __float2bfloat16(v - __bfloat162float(hi));
                }
    }, (size_t)K * cout);
}

// Mixed operands.  With S = S[n] (a power of two per output channel, max|w[:, n]| S in (14336, 28672]):
//   plane 0:  w16 = fp16(w S)                            x16 side: fp16(x)               product = S x16 w16
//   plane 1:  wl8 = e4m3((w - w16 / S) S)                x8  = e5m2(x)                   product = S x wl
//             w8  = e4m3(w S / 256)                      xl8 = e5m2(256 (x - x16))       product = S xl w
// so the accumulator is S[n] (x . w) and the epilogue multiplies column n by 1 / S[n] (exact).  |w16| <= 28672, |wl S| <= 8,
// |w8| <= 112: inside fp16 / e4m3; activations up to 65504 before fp16 saturates (|256 xl| <= 4096, far inside e5m2).
void pack_weights_mixed(const float* w_kc, int K, int cout, int cout_pad, std::vector<__nv_bfloat16>& out, std::vector<float>& iscale) {
    out.assign((size_t)2 * cout_pad * K, __float2bfloat16(0.0f));
    iscale.assign((size_t)cout_pad + 32, 1.0f);
    uint16_t* p0 = reinterpret_cast<uint16_t*>(out.data());
    uint8_t* p1 = reinterpret_cast<uint8_t*>(out.data() + (size_t)cout_pad * K);
    float* isc = iscale.data();
    parallel_channels(cout, [=](int n0, int n1) {
        for (int nb = n0; nb < n1; nb += 16) {
            const int ne = std::min(nb + 16, n1);
            float wmax[16] = {0.f}, S[16];
            for (int k = 0; k < K; ++k)
                for (int n = nb; n < ne; ++n) wmax[n - nb] = std::max(wmax[n - nb], std::fabs(w_kc[(size_t)k * cout + n]));
            for (int n = nb; n < ne; ++n) {
                S[n - nb] = wmax[n - nb] > 0.f ? std::exp2(std::floor(std::log2(28672.0f / wmax[n - nb]))) : 1.0f;
                isc[n] = 1.0f / S[n - nb];
            }
            for (int k = 0; k < K; ++k)
                for (int n = nb; n < ne; ++n) {
                    const float w = w_kc[(size_t)k * cout + n], s = S[n - nb];
                    const __half h = __float2half_rn(w * s);
                    p0[(size_t)n * K + k] = *reinterpret_cast<const uint16_t*>(&h);
                    const float wl = w - __half2float(h) / s;
                    uint8_t* row = p1 + ((size_t)n * K + (size_t)(k / 64) * 64) * 2;
                    row[k % 64] = (uint8_t)__nv_cvt_float_to_fp8(wl * s, __NV_SATFINITE, __NV_E4M3);
                    row[64 + k % 64] = (uint8_t)__nv_cvt_float_to_fp8(w * s * (1.0f / 256.0f), __NV_SATFINITE, __NV_E4M3);
                }
        }
    }, (size_t)K * cout);
}

bool tc_mixed_capable(const ConvParams& p) {
    return p.c1 > 0 && p.c1 % 64 == 0 && p.c2 % 64 == 0 && !p.kw_packed && !p.win_c && !p.stride_x && !p.row_pair &&
           (p.epi == EPI_LINEAR || p.epi == EPI_LSTM) && (p.pred_out == nullptr || p.phase4 == 1) && p.cout_pad % 32 == 0 && tc_eligible(p);
}

template <int BK>
static int max_clusters(int cs, size_t smem) {
    auto kern = conv_tc_kernel<BK, false>;
    if (cs <= 1) return kNumSMs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(kNumSMs / cs * cs));
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int tc_plan_create(ConvParams& p) {
    EVK_REQUIRE(tc_eligible(p), EVK_ERR_ARG, "conv_tc: shape not eligible for the tensor-core path");
    EVK_REQUIRE(p.x1s && p.w_tc && (p.c2 == 0 || p.x2s), EVK_ERR_ARG, "conv_tc: split operands missing");
    EVK_REQUIRE(!p.mixed || (tc_mixed_capable(p) && p.w_mx && p.w_iscale), EVK_ERR_ARG, "conv_tc: layer cannot run with mixed operands");
    EVK_REQUIRE(!(p.ys_mixed || p.hs_mixed) || ((p.epi == EPI_LSTM ? p.cout / 4 : p.cout) % 64 == 0 && !p.row_pair && p.s_wp == 0 &&
                                                 (p.epi == EPI_LSTM || p.epi == EPI_LINEAR)),
                EVK_ERR_ARG, "conv_tc: this epilogue cannot write a mixed-format split copy");
    const int bk = pick_bk(p);
    const int cout_pad = p.cout_pad;
    const bool rp = p.row_pair != 0;
    const bool ps = p.phase4 != 0, ps2 = p.phase4 >= 2, ps3 = p.phase4 == 3;   // (ps2: a tap row is skipped per tile -> U = y)
    EVK_REQUIRE(!ps || (!rp && p.epi == EPI_LINEAR && p.stride == 1 && p.pad == 0 && p.kh == 5 && p.kw == 5 && p.cout % 32 == 0 &&
                        cout_pad == 4 * p.cout && p.res == nullptr && p.ring_h != nullptr && p.ring_v != nullptr && !p.kw_packed), EVK_ERR_ARG,
                "conv_tc: phase-stacked mode needs a 5x5 stride-1 unpadded linear layer with cout %% 32 == 0 and a ring buffer");
    EVK_REQUIRE(!rp || (p.epi == EPI_LINEAR && p.stride == 1 && p.cout == 32 && cout_pad == 64 && p.Hout % 2 == 0 && p.res == nullptr &&
                        !p.kw_packed), EVK_ERR_ARG, "conv_tc: row-pair mode needs a stride-1 linear layer with cout 32 and an even height");
    const int e_kh = rp ? p.kh + 1 : p.kh, e_hout = rp ? p.Hout / 2 : p.Hout, e_cout = rp ? 2 * p.cout : ps ? 4 * p.cout : p.cout;
    const int s_y = rp ? 2 : p.stride, s_x = p.stride_x ? p.stride_x : p.stride;
    const int pad_x = p.pad_x >= 0 ? p.pad_x : p.pad;
    EVK_REQUIRE(cout_pad >= e_cout && cout_pad % 16 == 0, EVK_ERR_ARG, "conv_tc: cout_pad=%d must be a multiple of 16 >= cout", cout_pad);
    static bool attr_set_dev[64] = {false};          // function attributes are per device
    int dev_id = 0;
    EVK_CHECK_CUDA(cudaGetDevice(&dev_id));
    const bool attr_set = dev_id >= 0 && dev_id < 64 && attr_set_dev[dev_id];
    if (!attr_set) {
        const void* kerns[8] = {(const void*)conv_tc_kernel<64, false>, (const void*)conv_tc_kernel<32, false>,
                                (const void*)conv_tc_kernel<64, true>, (const void*)conv_tc_kernel<32, true>,
                                (const void*)conv_tc_kernel<16, false>, (const void*)conv_tc_kernel<16, true>,
                                (const void*)conv_tc_kernel<64, false, true>, (const void*)conv_tc_kernel<64, true, true>};
        for (const void* k : kerns) {
            EVK_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            EVK_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        }
        if (dev_id >= 0 && dev_id < 64) attr_set_dev[dev_id] = true;
    }
    const int chunks = (p.c1 + p.c2) / bk;
    const int granule = p.epi == EPI_LSTM ? 32 : 16;
    const int f_bn = env_int("EVK_TC_BN", 0), f_cs = env_int("EVK_TC_CS", 0), f_ux = env_int("EVK_TC_UX", -1);
    TcChoice best = {1, 0, 1, 0.0};
    for (int ux = 0; ux < 2; ++ux) {
        if (p.kw_packed ? ux != 1 : (rp || ps2) ? ux != 0 : (f_ux >= 0 && ux != f_ux)) continue;   // row-window input: atoms along x; row pairs / phase pairs: along y
        const int hu = ux ? p.Wout : e_hout, hv = ux ? e_hout : p.Wout;
        const long m_tiles = (long)ceil_div(hu, 8) * ceil_div(hv, 16) * p.N;
        const int ku = ux ? p.kw : e_kh, kv = ux ? e_kh : p.kw;
        // (MIXED: one accumulator of bn columns per stage, so a 256-column tile fits tensor memory)
        static const int mixed_bn_max = env_int("EVK_TC_MIXED_BN", 128);      // (256 fits tensor memory but leaves two weight stages: measured 50 % slower)
        for (int bn = (p.mixed && !ps) ? mixed_bn_max : 128; bn >= 16; bn -= 16) {
            if (cout_pad % bn != 0 || bn % granule != 0) continue;
            if ((p.pred_out != nullptr || rp) && bn != cout_pad) continue;
            if (ps && bn % 32 != 0) continue;             // a 32-column epilogue chunk must not straddle two phases
            if (ps2 && bn != (ps3 ? 1 : 2) * p.cout) continue;        // phase pairs: one row phase per N tile; single phases: one per tile
            if (f_bn > 0 && bn != f_bn && cout_pad % f_bn == 0 && f_bn % granule == 0) continue;
            if ((p.mixed || p.ys_mixed || p.hs_mixed) && bn % 32 != 0) continue;      // 32-column chunks only
            for (int cs = 1; cs <= 4; cs *= 2) {
                if ((bn / cs) % 8 != 0 || bn % cs != 0) continue;
                if (f_cs > 0 && cs != f_cs && (bn / f_cs) % 8 == 0 && f_cs <= 4) continue;
                const long n_super = (long)(cout_pad / bn) * ((m_tiles + cs - 1) / cs);
                const long ctas = n_super * cs;
                const long slots = (long)(kNumSMs / cs) * cs;
                const long waves = (ctas + slots - 1) / slots;
                const double cost = (double)waves * tile_cost(bk, ps2 ? ku - 1 : ku, ps3 ? kv - 1 : kv, ux ? s_y : s_x, chunks, bn, cs, ctas, nullptr, p.mixed != 0);
                if (best.bn == 0 || cost < best.cost) best = {ux, bn, cs, cost};
            }
        }
    }
    EVK_REQUIRE(best.bn >= 16, EVK_ERR_ARG, "conv_tc: no N tile for cout_pad=%d", cout_pad);
    const int ux = best.ux, bn = best.bn;
    int cs = best.cs;
    TcPlan* pl = new TcPlan();
    pl->bk = bk;
    pl->mixed = p.mixed != 0;
    TcArgs& a = pl->a;
    a.N = p.N; a.Hout = e_hout; a.Wout = p.Wout; a.pad_u = p.kw_packed ? 0 : (ux ? pad_x : p.pad); a.pad_v = ux ? p.pad : pad_x; a.kh = e_kh; a.kw = p.kw;
    a.su = ux ? s_x : s_y; a.sv = ux ? s_y : s_x;
    a.rp = rp ? 1 : 0; a.creal = p.cout; a.hreal = ps ? 2 * p.Hout : p.Hout;
    a.ps = p.phase4; a.wreal = ps ? 2 * p.Wout : p.Wout;
    a.ring_h = p.ring_h; a.ring_v = p.ring_v;
    // 16-column chunks for 32-column tiles: both halves of the epilogue warps get a chunk (linear, ConvGRU)
    const bool predw = p.pred_out != nullptr && p.win_c == 16 && p.kw_group == 2 && p.cout == 32 && bn == 32 && p.epi == EPI_LINEAR && p.pred_skip == nullptr &&
                       p.pred_skip_s == nullptr && !rp && !ps;
    a.predw = predw ? 1 : 0;
    a.cw = ((p.epi == EPI_LINEAR || p.epi == EPI_GRU_UR || p.epi == EPI_GRU_OUT) && bn == 32 && (p.pred_out == nullptr || predw) && env_int("EVK_TC_CW16", 1) &&
            !p.mixed && !p.ys_mixed) ? 16 : 32;
    // ... and 8-column chunks for 16-column tiles (FireNet's 16-channel layers)
    if ((p.epi == EPI_LINEAR || p.epi == EPI_GRU_OUT) && bn == 16 && p.pred_out == nullptr && env_int("EVK_TC_CW8", 1)) a.cw = 8;
    a.wide = (p.epi == EPI_LINEAR && p.cout % 16 == 0 && a.cw >= 16 && env_int("EVK_TC_WIDE_ST", 1)) ? 1 : 0;
    a.fastlin = (a.wide && !rp && !ps && (p.pred_out == nullptr || predw) && p.cout % 32 == 0 && bn % 32 == 0 && (p.act == ACT_RELU || p.act == ACT_NONE) &&
                 env_int("EVK_TC_FASTLIN", 1)) ? 1 : 0;
    a.act_floor = p.act == ACT_RELU ? 0.0f : -INFINITY;
    a.fastgru = (((p.epi == EPI_GRU_UR && p.cout % 32 == 0 && bn % 32 == 0) || (p.epi == EPI_GRU_OUT && p.cout % 16 == 0 && bn % 16 == 0)) &&
                 p.win_c > 0 && env_int("EVK_TC_FASTGRU", 1)) ? 1 : 0;
    a.fastps = (p.phase4 == 1 && p.cout == 32 && bn == 128 && p.pred_out != nullptr && p.y == nullptr && p.ys == nullptr && p.act == ACT_RELU &&
                env_int("EVK_TC_FASTPS", 1)) ? 1 : 0;
    a.ux = ux;
    a.tiles_u = ceil_div(ux ? p.Wout : e_hout, 8);
    a.tiles_v = ceil_div(ux ? e_hout : p.Wout, 16);
    a.ku = ux ? p.kw : e_kh; a.kv = ux ? e_kh : p.kw;
    a.chunks1 = p.c1 / bk; a.chunks2 = p.c2 / bk;
    a.bn = bn; a.n_tiles = cout_pad / bn; a.m_tiles = a.tiles_u * a.tiles_v * p.N; a.cout = e_cout; a.epi = p.epi; a.act = p.act;
    if (a.sv == 2) {
        a.n_groups = a.kv > 1 ? 2 : 1; a.g_step = 2;
        a.g_tap0[0] = 0; a.g_ntaps[0] = (a.kv + 1) / 2;
        a.g_tap0[1] = 1; a.g_ntaps[1] = a.kv / 2;
    } else {
        a.n_groups = 1; a.g_step = 1;
        a.g_tap0[0] = 0; a.g_ntaps[0] = a.kv;
        a.g_tap0[1] = 0; a.g_ntaps[1] = 0;
    }
    a.ar = 16 + a.g_ntaps[0] - 1;
    a.bias = p.bias; a.res = p.res; a.y = p.y; a.ys = p.ys;
    a.iscale = p.w_iscale; a.ys_mixed = p.ys_mixed; a.hs_mixed = p.hs_mixed;
    a.pred_skip_s = p.pred_skip_s; a.pred_skip_plane = p.pred_skip_plane;
    a.pred_w = p.pred_w; a.pred_skip = p.pred_skip; a.pred_out = p.pred_out; a.pred_bias = p.pred_bias; a.pred_sigmoid = p.pred_sigmoid;
    if (p.pred_out != nullptr && p.win_c > 0 && !(a.predw && a.fastlin && a.cw == 16)) {
        delete pl;
        EVK_REQUIRE(false, EVK_ERR_ARG, "conv_tc: the window-mode fused prediction layer needs the 16-column straight-line epilogue");
    }
    if (p.pred_out != nullptr && (p.epi != EPI_LINEAR || bn < e_cout || p.cout > 32)) {
        delete pl;
        EVK_REQUIRE(false, EVK_ERR_ARG, "conv_tc: the fused prediction layer needs all %d channels in one 32-column chunk (bn=%d)", p.cout, bn);
    }
    a.ys_plane = (long long)p.N * p.Hout * p.Wout * p.cout * (ps ? 4 : 1);
    a.s_pitch = 0; a.s_grp = 0; a.s_off = 0;
    a.c_prev = p.c_prev; a.c_new = p.c_new; a.h_new = p.h_new; a.hs_new = p.hs_new;
    a.h_prev = p.h_prev; a.u_in = p.u_in; a.u_out = p.u_out; a.hr_out = p.hr_out; a.hrs_out = p.hrs_out;
    // plane stride of the split copy of the recurrent output: hidden channels = cout/4 (LSTM), cout/2 (GRU u,r), cout (GRU out)
    a.hs_plane = (long long)p.N * p.Hout * p.Wout * (p.epi == EPI_LSTM ? p.cout / 4 : p.epi == EPI_GRU_UR ? p.cout / 2 : p.cout);
    if (p.s_wp > 0) {          // row-padded split outputs: s_c channels per pixel, pixel x at padded column x + s_left
        if (rp || ps || p.epi == EPI_LSTM) { delete pl; EVK_REQUIRE(false, EVK_ERR_ARG, "conv_tc: padded split outputs are not built for this epilogue"); }
        a.s_grp = p.epi == EPI_GRU_UR ? p.cout / 2 : p.cout;                  // split elements per GEMM row
        a.s_pitch = p.s_wp * p.s_c; a.s_off = p.s_left * p.s_c;
        a.ys_plane = a.hs_plane = (long long)p.N * p.Hout * a.s_pitch;
    }
    a.dbg = nullptr;
    a.hiprio = env_int("EVK_TC_HIPRIO", 1) ? 1 : 0;
    a.poll = p.mixed ? env_int("EVK_TC_POLL", 2) : 0;
    a.exp = env_int("EVK_TC_EXP", 0);
    const uint32_t row_bytes = bk * 2;
    const size_t a_stage = 2 * (size_t)a.ar * 8 * row_bytes;
    const size_t b_tap = 2 * (size_t)bn * row_bytes;
    // several taps per K block when the weights are small: fewer barrier round trips per MMA (the per-block issue
    // overhead is ~100+ cycles, a 16-deep slice of a small tile ~50-100)
    a.tpb = 1;
    {
        const int blocks_all = (a.chunks1 + a.chunks2) * a.ku * a.n_groups;      // K blocks per tile if a block takes a whole group
        // measured: pays for the one-slice blocks of 16-channel layers (FireNet +13%); with BK >= 32 the larger stages
        // leave too few of them in flight (the last decoder of E2VID lost 40%), so those keep one tap per block
        // (32-channel layers -- the first encoder: two-slice blocks cost 327 cycles per slice of issue time, -13 % with a tap group per block)
        const size_t limit = env_int("EVK_TC_TPB", 0) > 1 ? 40 * 1024 : bk <= 32 ? 24 * 1024 : 8 * 1024;
        if (a.g_ntaps[0] * b_tap <= limit && blocks_all >= 3 && env_int("EVK_TC_TPB", 0) != 1) a.tpb = a.g_ntaps[0];
    }
    if (ps3) a.tpb = 1;             // the per-tile tap range is applied tap by tap
    const size_t b_stage = (size_t)a.tpb * b_tap;
    const size_t budget = 227 * 1024 - 1024 - 512;
    int as = 3, bs = (int)((budget - std::min(budget, as * a_stage)) / b_stage);
    if (bs < 4) { as = 2; bs = (int)((budget - as * a_stage) / b_stage); }
    bs = std::min(bs, 8);
    if (bs < 2) { delete pl; EVK_REQUIRE(false, EVK_ERR_ARG, "conv_tc: tile does not fit shared memory (bn=%d)", bn); }
    a.a_stages = as; a.b_stages = bs;
    // per accumulator stage: [hi*hi + lo*hi | hi*lo] = 2*bn columns, read in 32-column windows (bn%32 tail -> pad)
    a.acc_cols = p.mixed ? (bn + 31) / 32 * 32 : (bn + (bn + 31) / 32 * 32 + 31) / 32 * 32;      // (MIXED: one accumulator of bn columns)
    const int kb = ps3 ? (a.chunks1 + a.chunks2) * (a.ku - 1) * (a.kv - 1) : (a.chunks1 + a.chunks2) * (ps2 ? a.ku - 1 : a.ku) * ((a.g_ntaps[0] + a.tpb - 1) / a.tpb + (a.n_groups > 1 ? (a.g_ntaps[1] + a.tpb - 1) / a.tpb : 0));
    a.issuers = (kb >= 2 && env_int("EVK_TC_ISSUERS", 2) >= 2) ? 2 : 1;
    // MIXED: three issuers in strict order (one warp requests both operands) -- an issuer's per-block preparation (barrier round
    // trips, ~230 instructions of descriptor set-up) then has two other issuers' blocks to hide behind instead of one
    if (p.mixed && kb >= 3 && bs >= 4 && env_int("EVK_TC_ISSUERS", 2) == 3) a.issuers = 3;      // (measured: +3 % at most; off)
    a.own_acc = (a.issuers == 2 && 4 * a.acc_cols <= 512 && env_int("EVK_TC_OWN_ACC", 1)) ? 1 : 0;
    a.acc_stride = a.own_acc ? 2 * a.acc_cols : a.acc_cols;
    if (a.own_acc && (a.b_stages & 1)) {      // free-running issuers need a slot to belong to one issuer (see the kernel)
        a.b_stages -= 1; bs -= 1;
    }
    a.deal = (p.mixed && a.own_acc && a.b_stages % 4 == 0 && env_int("EVK_TC_DEAL", 1) == 2) ? 2 : 1;      // (pairs measured 3 % slower: kept for experiments)
    a.lean = (p.mixed && a.tpb == 1 && a.issuers == 2 && a.own_acc && a.deal == 1 && env_int("EVK_TC_LEAN", 1)) ? 1 : 0;
    uint32_t cols = 32;
    while ((int)cols < 2 * a.acc_stride) cols <<= 1;
    if (cols > 512) { delete pl; EVK_REQUIRE(false, EVK_ERR_ARG, "conv_tc: accumulator does not fit tensor memory (bn=%d)", bn); }
    a.tmem_cols = cols;
    pl->smem = as * a_stage + bs * b_stage + 1024 + 16 * (as + bs) + 64;
    int ncl = bk == 64 ? max_clusters<64>(cs, pl->smem) : bk == 32 ? max_clusters<32>(cs, pl->smem) : max_clusters<16>(cs, pl->smem);
    if (ncl == 0 && cs > 1) {          // clusters of this size cannot be scheduled: fall back to independent CTAs
        cs = 1; ncl = kNumSMs;
    }
    a.cs = cs;
    const long n_super = (long)a.n_tiles * ((a.m_tiles + cs - 1) / cs);
    pl->grid = dim3((unsigned)(std::min<long>(n_super, ncl) * cs));
    // activations: [plane, n, V, U, c] (U = atom axis, V = shift axis)
    const int su = a.su, sv = a.sv;
    auto act_map = [&](CUtensorMap* m, const __nv_bfloat16* base, int C) -> int {
        if (p.kw_packed) {
            // row-window view of the packed head input [2][N][H][W+8][8]: "channel" dim = the 64 values starting at a
            // pixel, pixel stride 16 B (rows overlap), so one box row holds the kw taps x 8 channel slots of a kernel row
            // (kw_group = G > 1: GEMM row = G consecutive output pixels, so rows are G pixels = 16*G bytes apart)
            // (window mode, win_c > 0: the same view of a row-padded [N][H][win_wp][win_c] activation tensor, 2 * win_c bytes per pixel)
            const uint64_t G = (uint64_t)(p.kw_group > 1 ? p.kw_group : 1);
            const uint64_t pxb = p.win_c > 0 ? (uint64_t)p.win_c * 2 : 16;
            const uint64_t wp = p.win_c > 0 ? (uint64_t)p.win_wp : (uint64_t)p.Win * G + 8;
            const uint64_t dims[5] = {64, (uint64_t)p.Win, (uint64_t)p.Hin, (uint64_t)p.N, 2};
            const uint64_t str[4] = {pxb * G, wp * pxb, (uint64_t)p.Hin * wp * pxb, (uint64_t)p.N * p.Hin * wp * pxb};
            const uint32_t box[5] = {64, 8, (uint32_t)a.ar, 1, 1};
            const uint32_t es[5] = {1, 1, 1, 1, 1};
            return encode_tmap_bf16(m, base, 5, dims, str, box, es, 128);
        }
        const uint64_t sx = (uint64_t)C * 2, sy = (uint64_t)p.Win * C * 2;
        const uint64_t dims[5] = {(uint64_t)C, (uint64_t)(ux ? p.Win : p.Hin), (uint64_t)(ux ? p.Hin : p.Win), (uint64_t)p.N, 2};
        const uint64_t str[4] = {ux ? sx : sy, ux ? sy : sx, (uint64_t)p.Hin * p.Win * C * 2, (uint64_t)p.N * p.Hin * p.Win * C * 2};
        const uint32_t box[5] = {(uint32_t)bk, (uint32_t)(7 * su + 1), (uint32_t)((a.ar - 1) * sv + 1), 1, 1};
        const uint32_t es[5] = {1, (uint32_t)su, (uint32_t)sv, 1, 1};
        return encode_tmap_bf16(m, base, 5, dims, str, box, es, (int)row_bytes);
    };
    int r = act_map(&pl->tm_x1, p.x1s, p.c1);
    if (r == EVK_OK) r = p.c2 ? act_map(&pl->tm_x2, p.x2s, p.c2) : act_map(&pl->tm_x2, p.x1s, p.c1);
    if (r == EVK_OK) {
        const uint64_t K = (uint64_t)e_kh * p.kw * (p.c1 + p.c2);
        const uint64_t dims[3] = {K, (uint64_t)cout_pad, 2};
        const uint64_t str[2] = {K * 2, (uint64_t)cout_pad * K * 2};
        const uint32_t box[3] = {(uint32_t)bk, (uint32_t)(bn / cs), 1};
        const uint32_t es[3] = {1, 1, 1};
        r = encode_tmap_bf16(&pl->tm_w, p.mixed ? p.w_mx : p.w_tc, 3, dims, str, box, es, (int)row_bytes);
    }
    if (r != EVK_OK) { delete pl; return r; }
    if (env_int("EVK_TC_VERBOSE", 0))
        fprintf(stderr, "conv_tc plan: %dx%d s%d %d+%d->%d @%dx%dx%d  ux=%d bn=%d cs=%d issuers=%d own_acc=%d tpb=%d stages A%d/B%d ar=%d grid=%u smem=%zu%s\n", p.kh, p.kw,
                p.stride, p.c1, p.c2, p.cout, p.N, p.Hout, p.Wout, ux, bn, cs, a.issuers, a.own_acc, a.tpb, as, bs, a.ar, pl->grid.x, pl->smem, p.mixed ? " MIXED" : "");
    p.tc = pl;
    return EVK_OK;
}

void tc_plan_destroy(TcPlan* plan) { delete plan; }

// EVK_TC_TIMING=1 (eager launches only, e.g. evk_model_profile): per-CTA clock64 phase counters, averaged and printed
static int launch_conv_tc_timed(const ConvParams& p, cudaStream_t st);

int launch_conv_tc(const ConvParams& p, cudaStream_t st) {
    EVK_REQUIRE(p.tc != nullptr, EVK_ERR_STATE, "conv_tc: no plan");
    static const int timing = env_int("EVK_TC_TIMING", 0);
    if (timing && p.tc->a.dbg == nullptr) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cap);
        if (cap == cudaStreamCaptureStatusNone) return launch_conv_tc_timed(p, st);
    }
    const TcPlan& pl = *p.tc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = pl.grid;
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pl.a.cs > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = (unsigned)pl.a.cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    static const int pdl = env_int("EVK_TC_PDL", 1);
    if (pdl) {        // prologue (barriers, tensor memory, descriptor prefetch) overlaps the tail of the previous kernel
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = (unsigned)na;
    const bool dbg = pl.a.dbg != nullptr;
    if (pl.mixed) {
        if (dbg) EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, true, true>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
        else EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, false, true>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
    } else if (pl.bk == 64) {
        if (dbg) EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, true>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
        else EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<64, false>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
    } else if (pl.bk == 32) {
        if (dbg) EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<32, true>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
        else EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<32, false>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
    } else {
        if (dbg) EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, true>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
        else EVK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<16, false>, pl.tm_x1, pl.tm_x2, pl.tm_w, pl.a));
    }
    return EVK_OK;
}

static int launch_conv_tc_timed(const ConvParams& p, cudaStream_t st) {
    TcPlan& pl = *p.tc;
    const size_t n = (size_t)pl.grid.x * 12;
    unsigned long long* d = nullptr;
    EVK_CHECK_CUDA(cudaMalloc(&d, n * sizeof(unsigned long long)));
    EVK_CHECK_CUDA(cudaMemsetAsync(d, 0, n * sizeof(unsigned long long), st));
    pl.a.dbg = d;
    int r = launch_conv_tc(p, st);
    pl.a.dbg = nullptr;
    if (r != EVK_OK) { cudaFree(d); return r; }
    EVK_CHECK_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> h(n);
    EVK_CHECK_CUDA(cudaMemcpy(h.data(), d, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(d);
    double avg[12] = {0}, mx[12] = {0};
    for (size_t i = 0; i < n; ++i) { avg[i % 12] += (double)h[i] / pl.grid.x; mx[i % 12] = std::max(mx[i % 12], (double)h[i]); }
    const TcArgs& a = pl.a;
    const long n_super = (long)a.n_tiles * ((a.m_tiles + a.cs - 1) / a.cs);
    const double tiles_per_cta = (double)n_super * a.cs / pl.grid.x;
    const double k16 = (double)(a.chunks1 + a.chunks2) * a.kh * a.kw * (pl.bk / 16);
    fprintf(stderr, "TIMING %dx%d s%d c%d->%d @%dx%dx%d bn=%d cs=%d ux=%d grid=%u tiles/cta=%.1f k16/tile=%.0f | mma total %.0f (max %.0f) cyc = %.1f cyc/k16 | "
            "mma waits: tempty %.0f fullA %.0f fullB %.0f try-next %.0f issue+commit %.0f | producers wait: emptyA %.0f emptyB %.0f | epi total %.0f wait tfull %.0f tmem-read %.0f rest-of-chunk %.0f\n",
            a.kh, a.kw, a.su > a.sv ? a.su : a.sv, (a.chunks1 + a.chunks2) * pl.bk, a.cout, a.N, a.Hout, a.Wout, a.bn, a.cs, a.ux, pl.grid.x, tiles_per_cta, k16,
            avg[0], mx[0], avg[0] / (tiles_per_cta * k16), avg[1], avg[2], avg[3], avg[10], avg[11], avg[4], avg[5], avg[6], avg[7], avg[8], avg[9]);
    return EVK_OK;
}

// ------------------------------------------------------------------ head input: NCHW fp32 -> packed row-window split bf16
__global__ void __launch_bounds__(256) head_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int cin,
                                                        int H, int W, int left, int src_H, int src_W, int stride, int src_planes) {
    const int64_t total = (int64_t)N * H * W;
    const int64_t plane = (int64_t)N * H * (W + 8) * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W);
        const int yy = (int)((i / W) % H);
        const int n = (int)(i / ((int64_t)W * H));
        __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float v = c < cin ? __ldg(x + (((int64_t)n * src_planes + c) * src_H + (int64_t)yy * stride) * src_W + (int64_t)xx * stride) : 0.f;
            split_bf16(v, hi[c], lo[c]);
        }
        const int64_t o = (((int64_t)n * H + yy) * (W + 8) + xx + left) * 8;
        *reinterpret_cast<uint4*>(out + o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(out + plane + o) = *reinterpret_cast<const uint4*>(lo);
    }
}

int launch_head_pack(const float* x_nchw, __nv_bfloat16* packed, int N, int cin, int H, int W, int left, cudaStream_t st, int src_H, int src_W,
                     int stride, int src_planes) {
    EVK_REQUIRE(x_nchw && packed && cin >= 1 && cin <= 8 && left >= 0 && left <= 7 && stride >= 1, EVK_ERR_ARG, "head_pack: bad argument");
    if (src_H <= 0) src_H = H * stride;
    if (src_W <= 0) src_W = W * stride;
    if (src_planes <= 0) src_planes = cin;
    const int64_t total = (int64_t)N * H * W;
    head_pack_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 2368), 256, 0, st>>>(x_nchw, packed, N, cin, H, W, left, src_H, src_W, stride,
                                                                                             src_planes);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

void pack_weights_row_pair(const float* w_kc, int kh, int kw, int cin, int cout, std::vector<float>& out) {
    out.assign((size_t)(kh + 1) * kw * cin * 2 * cout, 0.f);
    for (int r = 0; r <= kh; ++r)
        for (int q = 0; q < kw; ++q)
            for (int c = 0; c < cin; ++c)
                for (int n = 0; n < cout; ++n) {
                    const size_t dst = ((size_t)(r * kw + q) * cin + c) * (2 * cout);
                    if (r < kh) out[dst + n] = w_kc[((size_t)(r * kw + q) * cin + c) * cout + n];                  // row 2y: tap r
                    if (r >= 1) out[dst + cout + n] = w_kc[((size_t)((r - 1) * kw + q) * cin + c) * cout + n];     // row 2y+1: tap r-1
                }
}

void pack_weights_pixel_pair(const float* w_kc, int kh, int kw, int cin, int cout, std::vector<float>& out) {
    const int kw2 = (kw + 1) / 2;            // pixel pairs per kernel row (5 taps -> 3 pairs, the last slot of the last pair unused)
    out.assign((size_t)kh * kw2 * 2 * cin * cout, 0.f);
    for (int r = 0; r < kh; ++r)
        for (int q = 0; q < kw; ++q)         // tap q = pair q/2, slot q%2
            for (int c = 0; c < cin; ++c)
                for (int n = 0; n < cout; ++n)
                    out[((size_t)(r * kw2 + q / 2) * (2 * cin) + (q % 2) * cin + c) * cout + n] = w_kc[((size_t)(r * kw + q) * cin + c) * cout + n];
}

void pack_weights_window(const float* w_kc, int kh, int kw, int cin, int c_tensor, int cout, int group, std::vector<float>& out) {
    const int T = cin / c_tensor, slots = 64 / c_tensor;
    out.assign((size_t)kh * 64 * T * group * cout, 0.f);
    for (int r = 0; r < kh; ++r)
        for (int g = 0; g < group; ++g)              // output pixel g of the group reads tap q from window slot g + q
            for (int q = 0; q < kw; ++q) {
                if (g + q >= slots) continue;
                for (int c = 0; c < cin; ++c) {
                    const int t = c / c_tensor, cc = c % c_tensor;
                    for (int n = 0; n < cout; ++n)
                        out[((size_t)r * 64 * T + t * 64 + (g + q) * c_tensor + cc) * (group * cout) + g * cout + n] =
                            w_kc[((size_t)(r * kw + q) * cin + c) * cout + n];
                }
            }
}

void pack_head_weights_rowwin(const float* w_kc, int kh, int kw, int cin, int cout, int group, std::vector<float>& out) {
    out.assign((size_t)kh * 64 * group * cout, 0.f);
    for (int r = 0; r < kh; ++r)
        for (int g = 0; g < group; ++g)              // output pixel g of the group sees tap q in window slot g + q
            for (int q = 0; q < kw; ++q)
                for (int c = 0; c < cin; ++c)
                    for (int n = 0; n < cout; ++n)
                        out[((size_t)r * 64 + (g + q) * 8 + c) * (group * cout) + g * cout + n] = w_kc[((size_t)(r * kw + q) * cin + c) * cout + n];
}

// ------------------------------------------------------------------ fp32 -> split planes
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        __nv_bfloat16 hi, lo;
        split_bf16(src[i], hi, lo);
        dst[i] = hi;
        dst[n + i] = lo;
    }
}

int launch_split(const float* src, __nv_bfloat16* dst, int64_t n, cudaStream_t st) {
    EVK_REQUIRE(src && dst && n > 0, EVK_ERR_ARG, "split: bad argument");
    split_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(n, 256), 2368), 256, 0, st>>>(src, dst, n);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

// fp32 NHWC -> mixed-format companion (evk_model_set_state), 8 channels per thread
__global__ void __launch_bounds__(256) split_mixed_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t groups, int C,
                                                          long long plane) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < groups; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = i * 8;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(src + o)), a1 = __ldg(reinterpret_cast<const float4*>(src + o + 4));
        const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        store_mixed8(dst, plane, (size_t)o, (int)(o % C), f);
    }
}

int launch_split_mixed(const float* src, __nv_bfloat16* dst, int64_t pixels, int C, cudaStream_t st) {
    EVK_REQUIRE(src && dst && pixels > 0 && C % 64 == 0, EVK_ERR_ARG, "split_mixed: bad argument");
    const int64_t groups = pixels * C / 8;
    split_mixed_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(groups, 256), 2368), 256, 0, st>>>(src, dst, groups, C, (long long)pixels * C);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

__global__ void __launch_bounds__(256) split_padded_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t rows,
                                                           int W, int C, int wp, int left) {
    const int64_t n = rows * W * C, plane = rows * (int64_t)wp * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int x = (int)((i / C) % W);
        const int64_t r = i / ((int64_t)C * W);
        __nv_bfloat16 hi, lo;
        split_bf16(src[i], hi, lo);
        const int64_t o = (r * wp + x + left) * C + c;
        dst[o] = hi;
        dst[plane + o] = lo;
    }
}

int launch_split_padded(const float* src, __nv_bfloat16* dst, int64_t rows, int W, int C, int wp, int left, cudaStream_t st) {
    EVK_REQUIRE(src && dst && rows > 0 && wp >= W + left, EVK_ERR_ARG, "split_padded: bad argument");
    split_padded_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(rows * W * C, 256), 2368), 256, 0, st>>>(src, dst, rows, W, C, wp, left);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

}  // namespace evk
