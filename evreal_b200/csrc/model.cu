// Stage 2 runner: builds a per-(architecture, shape) launch program from a
// reference state_dict and replays it as a CUDA graph, one frame per call.
//
// Reference semantics: eval.py:124-158 (factory), model/model.py:108-190,
// model/unet.py:85-143, model/legacy.py:32-187, model/submodules.py.
//
// HBM layout: activations NHWC fp32 (channel-contiguous so that the implicit
// GEMM K dimension is contiguous); recurrent state lives in the handle: ConvLSTM
// hidden state is ping-ponged between two buffers (neighbouring CTAs still read
// h(t-1) through the 3x3 window while h(t) is written), the cell state and the
// ConvGRU state are updated in place (pointwise).  Weights are repacked once at
// finalize(): eval-mode BatchNorm folded, K-major [kh*kw*Cin][Cout], LSTM/GRU
// gate channels interleaved so one thread owns all gates of a hidden channel.
#include <chrono>
#include <map>
#include <memory>
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdlib>

#include "conv.cuh"
#include "hyper.cuh"

namespace evk {

struct HostTensor {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

enum OpKind { OP_HEAD, OP_CONV, OP_UPSAMPLE_ADD, OP_PRED, OP_HYPER_CONTEXT, OP_HYPER_ATOMS, OP_HYPER_APPLY, OP_HYPER_APPLY_U, OP_HEAD_PACK, OP_NOP, OP_ZERO_INSERT_ADD, OP_ADD_PAD, OP_RING_LINES, OP_ADD_SPLIT, OP_SPADE_SHUFFLE, OP_SPADE_PRED, OP_LAYERNORM, OP_ATTENTION, OP_ADD_POS, OP_AVG6 };

struct Op {
    OpKind kind;
    ConvParams cp;              // OP_CONV
    // OP_HEAD
    const float* in = nullptr; const float* w = nullptr; const float* b = nullptr; float* out = nullptr;
    int N = 0, cin = 0, H = 0, W = 0, k = 0, cout = 0;
    // OP_UPSAMPLE_ADD / OP_PRED
    const float* skip = nullptr;
    __nv_bfloat16* out_s = nullptr;   // OP_HEAD / OP_UPSAMPLE_ADD: split-bf16 copy of `out` for a tensor-core consumer
    __nv_bfloat16* lines_h = nullptr; __nv_bfloat16* lines_v = nullptr;   // OP_RING_LINES outputs
    int mixed = 0;                    // OP_ADD_PAD / OP_RING_LINES: the padded split tensor is in the mixed format (conv.cuh, ConvParams::mixed)
    int ring_line = 0;                // OP_CONV: 1 / 2 = horizontal / vertical border-line convolution of a phase-stacked decoder
    float bias0 = 0.f;
    int sigmoid = 0;
    // OP_HEAD_PACK: source planes [N, src_planes, srcH, srcW] sampled every `stride` pixels (0 = the op's own H x W, stride 1)
    int srcH = 0, srcW = 0, stride = 1, src_planes = 0;
    // OP_SPADE_SHUFFLE: in = conv output [N,H,W,4*cout], skip = (gamma | beta) [N,2H,2W,2*cout], w = alpha, b = shift -> out / out_s
    // OP_SPADE_PRED: in = last hidden state, skip = head, w = [cin][3], bias3 -> prev3 (NCHW, 3 planes) and out (image)
    float bias3[3] = {0.f, 0.f, 0.f};
    float* prev3 = nullptr;
    // ET-Net token path.  OP_LAYERNORM: in, w = gamma, b = beta -> out / out_s over N*H*W tokens of 256 channels.
    // OP_ATTENTION: in = q, skip = k, w = v (row-strided views of the projection outputs: q_stride / kv_stride floats per token),
    // H*W query tokens, Lk key tokens -> out / out_s.  OP_ADD_POS: out = in + w (sine table [H*W][256]).  OP_AVG6: out = mean(six).
    int q_stride = 0, kv_stride = 0, Lk = 0;
    const float* six[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    HyperParams hp;             // OP_HYPER_*
    double flops = 0.0;
};

struct StateBuf {
    float* buf[2] = {nullptr, nullptr};   // buf[1] only for ping-ponged LSTM hidden state
    int C = 0, H = 0, W = 0;
    bool pingpong = false;
};

}  // namespace evk

using namespace evk;

struct evk_model {
    evk_model_config cfg;
    std::map<std::string, HostTensor> sd;
    bool finalized = false;
    std::vector<void*> allocs;
    std::vector<Op> ops[2];          // per hidden-state parity
    std::vector<StateBuf> states;
    // Network input [N,bins,H,W] and output [N,1,H,W], DOUBLE-BUFFERED by forward parity: forward k reads in_bufs[k & 1] and
    // writes out_bufs[k & 1], so a caller may fill the next input and consume the previous output on other streams while a
    // forward runs (pipeline.py), and HyperE2VID's "previous padded reconstruction" (model/model.py:139-143) is simply the
    // other output buffer -- no copy.  The builders wire parity 0 (in_buf / out_buf / prev_rec); parity 1 is patched after.
    float* in_bufs[2] = {nullptr, nullptr};
    float* out_bufs[2] = {nullptr, nullptr};
    float* in_buf = nullptr;         // = in_bufs[0]
    float* out_buf = nullptr;        // = out_bufs[0]
    float* prev_rec = nullptr;       // = out_bufs[1]
    int parity = 0;
    float* prev3 = nullptr;          // SPADE-E2VID: previous 3-channel reconstruction [N,3,H,W] (the SPADE layers' conditioning input)
    bool spade_first = true;         // SPADE-E2VID: no previous reconstruction yet (model/spade_e2v.py:140-147)
    int last_launches = 0;
    double flops = 0.0;
    cudaStream_t cap_stream = nullptr;
    cudaStream_t cap_stream2 = nullptr;      // second branch of the captured graph (the two border-line convolutions of a decoder run side by side)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    bool use_graph = true;
    std::map<const float*, __nv_bfloat16*> split_of;   // fp32 activation buffer -> its split-bf16 companion
    std::map<const float*, size_t> buf_elems;          // element count of every activation / state buffer
    std::vector<TcPlan*> plans;
    int tc_convs = 0;
    // window mode (FireNet, 16-channel tensors at full resolution; conv.cuh ConvParams::win_c): every split companion is
    // row-padded [2][N][H][W + 2][C] (pixel x at column x + 1)
    int win_wp = 0;
    std::map<const float*, size_t> split_bytes;        // allocation size of every split companion
    // mixed operands (conv.cuh, ConvParams::mixed): on unless EVK_MIXED=0; the fp32 buffers whose companion is in the mixed format
    bool mixed_enabled = true;
    std::map<const float*, int> mixed_bufs;            // fp32 buffer -> channels per pixel
    int mixed_convs = 0;

    DeviceArena arena;               // every device buffer of the program (evk_common.cuh)
    float* dalloc(size_t nfloat) {
        void* p = arena.alloc(nfloat * sizeof(float));
        allocs.push_back(p);             // (nullptr = out of memory: checked at the end of finalize)
        if (p) buf_elems[(const float*)p] = nfloat;
        return (float*)p;
    }
    void* dalloc_bytes(size_t bytes) {
        void* p = arena.alloc(bytes);
        allocs.push_back(p);
        return p;
    }
    float* upload(const std::vector<float>& v) {
        float* p = dalloc(v.size());
        if (p) cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
        return p;
    }
    const HostTensor* find(const std::string& name) const {
        auto it = sd.find(name);
        return it == sd.end() ? nullptr : &it->second;
    }
};

namespace evk {

// ---------------------------------------------------------------- packing
struct Packed {
    std::vector<float> w;   // [kh*kw*cin][cout_packed]
    std::vector<float> b;   // [cout_packed]
    int cout = 0, cin = 0, kh = 0, kw = 0;
};

// Folds an optional eval-mode BatchNorm (prefix bn) and an optional bias into
// one conv (float64 arithmetic), and permutes output channels: packed channel
// n takes reference channel perm[n].
// transposed: the tensor is a ConvTranspose2d weight [Cin, Cout, kh, kw]; the packed result is the equivalent stride-1
// convolution on the zero-inserted input: Wc[co][ci][r][q] = Wt[ci][co][kh-1-r][kw-1-q].
static int pack_conv(const evk_model* m, const std::string& wname, const std::string& bname, const std::string& bn,
                     const std::vector<int>* perm, Packed& out, bool transposed = false) {
    const HostTensor* w = m->find(wname);
    EVK_REQUIRE(w && w->shape.size() == 4, EVK_ERR_KEY, "missing weight tensor '%s'", wname.c_str());
    const int co = (int)w->shape[transposed ? 1 : 0], ci = (int)w->shape[transposed ? 0 : 1], kh = (int)w->shape[2], kw = (int)w->shape[3];
    const HostTensor* bias = bname.empty() ? nullptr : m->find(bname);
    const HostTensor *g = nullptr, *beta = nullptr, *mean = nullptr, *var = nullptr;
    if (!bn.empty() && m->find(bn + ".running_mean")) {
        g = m->find(bn + ".weight"); beta = m->find(bn + ".bias");
        mean = m->find(bn + ".running_mean"); var = m->find(bn + ".running_var");
        EVK_REQUIRE(g && beta && mean && var, EVK_ERR_KEY, "incomplete BatchNorm '%s'", bn.c_str());
    }
    const int cop = perm ? (int)perm->size() : co;
    out.cout = cop; out.cin = ci; out.kh = kh; out.kw = kw;
    out.w.assign((size_t)kh * kw * ci * cop, 0.f);
    out.b.assign(cop, 0.f);
    // (output channels in parallel, 16 at a time so that the K-major destination rows are written a cache line at once: the
    //  serial form of this loop was most of the 160 ms a HyperE2VID / E2VID handle took to build)
    float* ow = out.w.data();
    float* ob = out.b.data();
    parallel_channels(cop, [=](int n0, int n1) {
        for (int nb = n0; nb < n1; nb += 16) {
            const int ne = std::min(nb + 16, n1);
            double scale[16];
            int src[16];
            for (int n = nb; n < ne; ++n) {
                const int r = perm ? (*perm)[n] : n;
                src[n - nb] = r;
                scale[n - nb] = 1.0;
                if (r < 0) continue;   // padding channel
                double shift = bias ? (double)bias->data[r] : 0.0;
                if (g) {
                    scale[n - nb] = (double)g->data[r] / std::sqrt((double)var->data[r] + 1e-5);
                    shift = (shift - (double)mean->data[r]) * scale[n - nb] + (double)beta->data[r];
                }
                ob[n] = (float)shift;
            }
            for (int y = 0; y < kh; ++y)
                for (int x = 0; x < kw; ++x)
                    for (int c = 0; c < ci; ++c) {
                        float* dst = ow + ((size_t)(y * kw + x) * ci + c) * cop;
                        for (int n = nb; n < ne; ++n) {
                            const int r = src[n - nb];
                            if (r < 0) continue;
                            const double wv = transposed ? (double)w->data[(((size_t)c * co + r) * kh + (kh - 1 - y)) * kw + (kw - 1 - x)]
                                                         : (double)w->data[(((size_t)r * ci + c) * kh + y) * kw + x];
                            dst[n] = (float)(wv * scale[n - nb]);
                        }
                    }
        }
    }, (size_t)kh * kw * ci * cop);
    return EVK_OK;
}

static double conv_flops(const ConvParams& p, int real_cout) {
    return 2.0 * real_cout * (p.c1 + p.c2) * p.kh * p.kw * (double)p.N * p.Hout * p.Wout;
}

struct Builder {
    evk_model* m;
    int N;
    int rc = EVK_OK;

    float* act(int H, int W, int C) { return m->dalloc((size_t)N * H * W * C); }

    // generic ConvLayer-like op appended to both parities
    int conv(const std::string& wname, const std::string& bname, const std::string& bn, const float* x, int cin, int Hin,
             int Win, int stride, int pad, int act, const float* res, float* y, int* cout_out, bool transposed = false) {
        Packed pk;
        int r = pack_conv(m, wname, bname, bn, nullptr, pk, transposed);
        if (r != EVK_OK) return r;
        EVK_REQUIRE(pk.cin == cin, EVK_ERR_KEY, "'%s': expected %d input channels, checkpoint has %d", wname.c_str(), cin, pk.cin);
        return conv_packed(pk, x, cin, Hin, Win, stride, pad, act, res, y, cout_out);
    }

    // split-bf16 K-major weights for the tensor-core path (only when the layer shape qualifies)
    void attach_tc_weights(ConvParams& p, const Packed& pk) {
        if (m->cfg.precision != 0 || !tc_eligible(p)) return;
        std::vector<__nv_bfloat16> wt;
        if (p.epi == EPI_LINEAR && pk.cout == 32 && p.stride == 1 && p.Hout > 0 && p.Hout % 2 == 0 && p.res == nullptr && !p.kw_packed &&
            getenv("EVK_NO_ROW_PAIR") == nullptr) {
            // last decoder: two output rows per GEMM row (conv.cuh, ConvParams::row_pair)
            std::vector<float> w2;
            pack_weights_row_pair(pk.w.data(), pk.kh, pk.kw, pk.cin, pk.cout, w2);
            p.row_pair = 1;
            p.cout_pad = 2 * pk.cout;
            pack_weights_tc(w2.data(), (pk.kh + 1) * pk.kw * pk.cin, 2 * pk.cout, p.cout_pad, wt);
        } else {
            p.cout_pad = (pk.cout + 15) / 16 * 16;
            pack_weights_tc(pk.w.data(), pk.kh * pk.kw * pk.cin, pk.cout, p.cout_pad, wt);
        }
        void* d = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
        if (d) cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
        p.w_tc = (const __nv_bfloat16*)d;
        // the mixed-operand form of the same weights (conv.cuh, ConvParams::mixed); wire_tc decides which one the layer runs
        if (m->mixed_enabled && tc_mixed_capable(p)) {
            std::vector<float> isc;
            pack_weights_mixed(pk.w.data(), pk.kh * pk.kw * pk.cin, pk.cout, p.cout_pad, wt, isc);
            void* dm = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
            if (dm) cudaMemcpy(dm, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
            p.w_mx = (const __nv_bfloat16*)dm;
            p.w_iscale = m->upload(isc);
        }
    }

    // Window form of a stride-1 3x3 layer whose input tensors have 16 channels (conv.cuh, ConvParams::win_c): 4 pixels of
    // 16 channels are one 128-byte K row, 2 output pixels share a window (N = 2 * cout).  `pk` holds the packed weights /
    // bias of the plain layer ([kh*kw*(c1+c2)][cout]); the outputs are the same memory seen as [N, H, W/2, 2*cout].
    bool window_ok(const ConvParams& p, const Packed& pk) const {
        return m->win_wp > 0 && m->cfg.precision == 0 && p.stride == 1 && pk.kh == 3 && pk.kw == 3 && p.c1 == 16 && (p.c2 == 0 || p.c2 == 16) &&
               p.Wout % 2 == 0 && p.Win == p.Wout && pk.cout % 8 == 0 && 2 * pk.cout <= 128 && !p.kw_packed;
    }
    int apply_window(ConvParams& p, const Packed& pk) {
        const int G = 2;
        std::vector<float> ww, bw((size_t)G * pk.cout);
        pack_weights_window(pk.w.data(), pk.kh, pk.kw, pk.cin, 16, pk.cout, G, ww);
        for (int g = 0; g < G; ++g) std::copy(pk.b.begin(), pk.b.end(), bw.begin() + (size_t)g * pk.cout);
        const int T = pk.cin / 16;
        p.kw_packed = pk.kw; p.kw_group = G; p.win_c = 16; p.win_wp = m->win_wp;
        p.kh = pk.kh; p.kw = 1; p.c1 = 64; p.c2 = T > 1 ? 64 : 0;
        p.Win = p.Wout = p.Wout / G;
        p.cout = G * pk.cout; p.cout_pad = (G * pk.cout + 15) / 16 * 16;
        p.bias = m->upload(bw); p.w = nullptr;                     // (tensor-core only)
        std::vector<__nv_bfloat16> wt;
        pack_weights_tc(ww.data(), pk.kh * 64 * T, G * pk.cout, p.cout_pad, wt);
        void* d = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
        EVK_REQUIRE(d != nullptr && p.bias != nullptr, EVK_ERR_CUDA, "out of device memory for the window weights");
        cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
        p.w_tc = (const __nv_bfloat16*)d;
        return EVK_OK;
    }

    int conv_packed(const Packed& pk, const float* x, int cin, int Hin, int Win, int stride, int pad, int act,
                    const float* res, float* y, int* cout_out) {
        Op op; op.kind = OP_CONV;
        ConvParams& p = op.cp;
        p.x1 = x; p.c1 = cin; p.N = N; p.Hin = Hin; p.Win = Win;
        p.kh = pk.kh; p.kw = pk.kw; p.stride = stride; p.pad = pad;
        p.Hout = (Hin + 2 * pad - pk.kh) / stride + 1;
        p.Wout = (Win + 2 * pad - pk.kw) / stride + 1;
        p.bias = m->upload(pk.b); p.cout = pk.cout;
        p.epi = EPI_LINEAR; p.act = act; p.res = res; p.y = y;
        op.flops = conv_flops(p, pk.cout);
        if (window_ok(p, pk)) {
            int r = apply_window(p, pk);
            if (r != EVK_OK) return r;
        } else if (m->cfg.precision == 0 && stride == 2 && pad == 2 && pk.kh == 5 && pk.kw == 5 && cin == 32 && Win % 2 == 0 && res == nullptr &&
                   tc_eligible(p) && getenv("EVK_NO_PIXEL_PAIR") == nullptr) {
            // first encoder: pixel pairs as 64-channel K rows (conv.cuh, ConvParams::stride_x)
            std::vector<float> w2;
            pack_weights_pixel_pair(pk.w.data(), pk.kh, pk.kw, pk.cin, pk.cout, w2);
            p.c1 = 2 * cin; p.Win = Win / 2; p.kw = 3; p.stride_x = 1; p.pad_x = 1; p.w = nullptr;        // (tensor-core only)
            std::vector<__nv_bfloat16> wt;
            p.cout_pad = (pk.cout + 15) / 16 * 16;
            pack_weights_tc(w2.data(), pk.kh * 3 * 2 * cin, pk.cout, p.cout_pad, wt);
            void* d = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
            EVK_REQUIRE(d != nullptr, EVK_ERR_CUDA, "out of device memory for the pixel-pair weights");
            cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
            p.w_tc = (const __nv_bfloat16*)d;
        } else {
            attach_tc_weights(p, pk);
            if (p.w_tc == nullptr) p.w = m->upload(pk.w);      // fp32 K-major weights: only the CUDA-core kernel reads them
        }
        m->ops[0].push_back(op); m->ops[1].push_back(op);
        if (cout_out) *cout_out = pk.cout;
        return EVK_OK;
    }
};

static int make_state(evk_model* m, int C, int H, int W, bool pingpong) {
    StateBuf s; s.C = C; s.H = H; s.W = W; s.pingpong = pingpong;
    const size_t n = (size_t)m->cfg.batch * H * W * C;
    s.buf[0] = m->dalloc(n);
    s.buf[1] = pingpong ? m->dalloc(n) : s.buf[0];
    EVK_REQUIRE(s.buf[0] && s.buf[1], EVK_ERR_CUDA, "out of device memory for recurrent state");
    m->states.push_back(s);
    return (int)m->states.size() - 1;
}

// ---- ConvLSTM: gates = conv3x3(cat(x,h)); packed channel ch*4+g <- reference g*C+ch
static int add_lstm(Builder& B, const std::string& pfx, const float* x, int C, int H, int W, int hs, int cs) {
    evk_model* m = B.m;
    std::vector<int> perm(4 * C);
    for (int ch = 0; ch < C; ++ch)
        for (int g = 0; g < 4; ++g) perm[ch * 4 + g] = g * C + ch;
    Packed pk;
    int r = pack_conv(m, pfx + ".Gates.weight", pfx + ".Gates.bias", "", &perm, pk);
    if (r != EVK_OK) return r;
    EVK_REQUIRE(pk.cin == 2 * C && pk.kh == 3, EVK_ERR_KEY, "'%s.Gates': unexpected shape", pfx.c_str());
    const float* b = m->upload(pk.b);
    ConvParams wt;   // carries the tensor-core weights shared by both parities
    wt.c1 = C; wt.c2 = C; wt.stride = 1; wt.cout = 4 * C; wt.epi = EPI_LSTM; wt.kh = wt.kw = 3;
    B.attach_tc_weights(wt, pk);
    const float* w = wt.w_tc != nullptr ? nullptr : m->upload(pk.w);      // fp32 K-major weights: only the CUDA-core kernel reads them
    for (int par = 0; par < 2; ++par) {
        Op op; op.kind = OP_CONV;
        ConvParams& p = op.cp;
        p.x1 = x; p.c1 = C; p.x2 = m->states[hs].buf[par]; p.c2 = C;
        p.N = B.N; p.Hin = p.Hout = H; p.Win = p.Wout = W; p.kh = p.kw = 3; p.stride = 1; p.pad = 1;
        p.w = w; p.bias = b; p.cout = 4 * C; p.epi = EPI_LSTM;
        p.w_tc = wt.w_tc; p.cout_pad = wt.cout_pad; p.w_mx = wt.w_mx; p.w_iscale = wt.w_iscale;
        p.c_prev = m->states[cs].buf[0]; p.c_new = m->states[cs].buf[0];
        p.h_new = m->states[hs].buf[par ^ 1];
        op.flops = conv_flops(p, 4 * C);
        m->ops[par].push_back(op);
    }
    return EVK_OK;
}

// ---- ConvGRU: (update, reset) fused into one conv, then the candidate conv on cat(x, h*r)
static int add_gru(Builder& B, const std::string& pfx, const float* x, int C, int H, int W, int hs, float* u_buf, float* hr_buf) {
    evk_model* m = B.m;
    const HostTensor* wu = m->find(pfx + ".update_gate.weight");
    const HostTensor* wr = m->find(pfx + ".reset_gate.weight");
    const HostTensor* bu = m->find(pfx + ".update_gate.bias");
    const HostTensor* br = m->find(pfx + ".reset_gate.bias");
    EVK_REQUIRE(wu && wr && bu && br, EVK_ERR_KEY, "missing ConvGRU gates under '%s'", pfx.c_str());
    const int cin = (int)wu->shape[1], k = (int)wu->shape[2];
    EVK_REQUIRE(cin == 2 * C && (int)wu->shape[0] == C, EVK_ERR_KEY, "'%s': unexpected ConvGRU shape", pfx.c_str());
    std::vector<float> w((size_t)k * k * cin * 2 * C), b(2 * C);
    for (int ch = 0; ch < C; ++ch) {
        b[ch * 2 + 0] = bu->data[ch];
        b[ch * 2 + 1] = br->data[ch];
        for (int c = 0; c < cin; ++c)
            for (int y = 0; y < k; ++y)
                for (int xx = 0; xx < k; ++xx) {
                    const size_t src = (((size_t)ch * cin + c) * k + y) * k + xx;
                    const size_t dst = ((size_t)(y * k + xx) * cin + c) * (2 * C) + ch * 2;
                    w[dst + 0] = wu->data[src];
                    w[dst + 1] = wr->data[src];
                }
    }
    float* h = m->states[hs].buf[0];
    {
        Op op; op.kind = OP_CONV;
        ConvParams& p = op.cp;
        p.x1 = x; p.c1 = C; p.x2 = h; p.c2 = C; p.N = B.N; p.Hin = p.Hout = H; p.Win = p.Wout = W;
        p.kh = p.kw = k; p.stride = 1; p.pad = k / 2;
        p.w = m->upload(w); p.bias = m->upload(b); p.cout = 2 * C; p.epi = EPI_GRU_UR;
        p.h_prev = h; p.u_out = u_buf; p.hr_out = hr_buf;
        op.flops = conv_flops(p, 2 * C);
        {
            Packed pk2; pk2.w = w; pk2.b = b; pk2.cout = 2 * C; pk2.cin = cin; pk2.kh = k; pk2.kw = k;
            if (B.window_ok(p, pk2)) {
                int r = B.apply_window(p, pk2);
                if (r != EVK_OK) return r;
            } else {
                B.attach_tc_weights(p, pk2);
            }
        }
        m->ops[0].push_back(op); m->ops[1].push_back(op);
    }
    {
        Packed pk;
        int r = pack_conv(m, pfx + ".out_gate.weight", pfx + ".out_gate.bias", "", nullptr, pk);
        if (r != EVK_OK) return r;
        Op op; op.kind = OP_CONV;
        ConvParams& p = op.cp;
        p.x1 = x; p.c1 = C; p.x2 = hr_buf; p.c2 = C; p.N = B.N; p.Hin = p.Hout = H; p.Win = p.Wout = W;
        p.kh = p.kw = k; p.stride = 1; p.pad = k / 2;
        p.w = m->upload(pk.w); p.bias = m->upload(pk.b); p.cout = C; p.epi = EPI_GRU_OUT;
        p.h_prev = h; p.u_in = u_buf; p.h_new = h;
        op.flops = conv_flops(p, C);
        if (B.window_ok(p, pk)) {
            r = B.apply_window(p, pk);
            if (r != EVK_OK) return r;
        } else {
            B.attach_tc_weights(p, pk);
        }
        m->ops[0].push_back(op); m->ops[1].push_back(op);
    }
    return EVK_OK;
}

static int add_resblock(Builder& B, const std::string& pfx, const float* x, float* tmp, float* y, int C, int H, int W) {
    int r = B.conv(pfx + ".conv1.weight", pfx + ".conv1.bias", pfx + ".bn1", x, C, H, W, 1, 1, ACT_RELU, nullptr, tmp, nullptr);
    if (r != EVK_OK) return r;
    return B.conv(pfx + ".conv2.weight", pfx + ".conv2.bias", pfx + ".bn2", tmp, C, H, W, 1, 1, ACT_RELU, x, y, nullptr);
}

// ConvLayer whose input is an NCHW tensor with <= 8 channels (the event tensor; SPADE's 3-channel conditioning image): packed
// into the row-window layout and run as a (kh x 1) implicit GEMM on the tensor cores, or on CUDA cores for other shapes.
// src: [N, cin, srcH, srcW] sampled every `stride` pixels -> H x W (nearest reduction by an integer factor; stride 1 for the head).
static int add_rowwin_conv(Builder& B, const std::string& conv_pfx, const std::string& bn, const float* src, int srcH, int srcW, int stride,
                           int H, int W, int act, float* y, int cout_expected, int cin_expected) {
    evk_model* m = B.m;
    Packed pk;
    int r = pack_conv(m, conv_pfx + ".weight", conv_pfx + ".bias", bn, nullptr, pk);
    if (r != EVK_OK) return r;
    EVK_REQUIRE(pk.cin == cin_expected && pk.cout == cout_expected, EVK_ERR_KEY,
                "'%s': expected [%d,%d,k,k], checkpoint has [%d,%d,%d,%d]", conv_pfx.c_str(), cout_expected,
                cin_expected, pk.cout, pk.cin, pk.kh, pk.kw);
    const double flops = 2.0 * pk.cout * pk.cin * pk.kh * pk.kw * (double)B.N * H * W;
    if (m->cfg.precision == 0 && pk.cin <= 8 && pk.kw <= 8 && pk.kh == pk.kw && pk.cout % 16 == 0 && getenv("EVK_HEAD_SIMT") == nullptr) {
        // tensor-core head: pack the NCHW event tensor into the row-window layout (conv.cuh) and run the layer as a
        // (kh x 1) implicit GEMM with 64 "channels" per kernel row
        const size_t plane = (size_t)B.N * H * (W + 8) * 8;
        __nv_bfloat16* packed = (__nv_bfloat16*)m->dalloc_bytes(2 * plane * sizeof(__nv_bfloat16));
        EVK_REQUIRE(packed != nullptr, EVK_ERR_CUDA, "out of device memory for the packed head input");
        {
            Op op; op.kind = OP_HEAD_PACK;
            op.in = src; op.out_s = packed; op.N = B.N; op.cin = pk.cin; op.H = H; op.W = W; op.k = pk.kw;
            op.srcH = srcH; op.srcW = srcW; op.stride = stride; op.src_planes = pk.cin;
            m->ops[0].push_back(op); m->ops[1].push_back(op);
        }
        // G consecutive output pixels per GEMM row (conv.cuh, ConvParams::kw_group): N = G * cout columns per MMA
        int G = 1;
        if (getenv("EVK_HEAD_GROUP") == nullptr || atoi(getenv("EVK_HEAD_GROUP")) > 1)
            for (int g = 4; g >= 2; g /= 2)
                if (W % g == 0 && g + pk.kw - 1 <= 8 && g * pk.cout <= 128) { G = g; break; }
        Op op; op.kind = OP_CONV;
        ConvParams& p = op.cp;
        p.x1 = nullptr; p.c1 = 64; p.x1s = packed; p.kw_packed = pk.kw; p.kw_group = G;
        p.N = B.N; p.Hin = p.Hout = H; p.Win = p.Wout = W / G; p.kh = pk.kh; p.kw = 1; p.stride = 1; p.pad = pk.kh / 2;
        std::vector<float> bg((size_t)G * pk.cout);
        for (int g = 0; g < G; ++g) std::copy(pk.b.begin(), pk.b.end(), bg.begin() + (size_t)g * pk.cout);
        p.bias = m->upload(bg); p.cout = G * pk.cout; p.epi = EPI_LINEAR; p.act = act; p.y = y;
        std::vector<float> wr;
        pack_head_weights_rowwin(pk.w.data(), pk.kh, pk.kw, pk.cin, pk.cout, G, wr);
        std::vector<__nv_bfloat16> wt;
        p.cout_pad = G * pk.cout;
        pack_weights_tc(wr.data(), pk.kh * 64, G * pk.cout, p.cout_pad, wt);
        void* d = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
        EVK_REQUIRE(d != nullptr, EVK_ERR_CUDA, "out of device memory for the head weights");
        cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
        p.w_tc = (const __nv_bfloat16*)d;
        op.flops = flops;
        op.cin = pk.cin; op.k = pk.kw;       // real shape, for the description
        m->ops[0].push_back(op); m->ops[1].push_back(op);
        return EVK_OK;
    }
    EVK_REQUIRE(stride == 1 && act == ACT_RELU, EVK_ERR_ARG, "'%s': the CUDA-core head kernel covers the plain ReLU head only", conv_pfx.c_str());
    Op op; op.kind = OP_HEAD;
    op.in = src; op.w = m->upload(pk.w); op.b = m->upload(pk.b); op.out = y;
    op.N = B.N; op.cin = pk.cin; op.H = H; op.W = W; op.k = pk.kh; op.cout = pk.cout;
    op.flops = flops;
    m->ops[0].push_back(op); m->ops[1].push_back(op);
    return EVK_OK;
}

static int add_head(Builder& B, const std::string& conv_pfx, const std::string& bn, float* y, int cout_expected) {
    evk_model* m = B.m;
    return add_rowwin_conv(B, conv_pfx, bn, m->in_buf, m->cfg.height, m->cfg.width, 1, m->cfg.height, m->cfg.width, ACT_RELU, y, cout_expected,
                           m->cfg.num_bins);
}

static int add_pred(Builder& B, const std::string& pfx, const float* x, const float* skip, int cin, int H, int W) {
    evk_model* m = B.m;
    Packed pk;
    int r = pack_conv(m, pfx + ".conv2d.weight", pfx + ".conv2d.bias", pfx + ".norm_layer", nullptr, pk);
    if (r != EVK_OK) return r;
    EVK_REQUIRE(pk.cin == cin && pk.kh == 1, EVK_ERR_KEY, "'%s': unexpected prediction layer shape", pfx.c_str());
    std::vector<float> w0(cin);
    for (int c = 0; c < cin; ++c) w0[c] = pk.w[(size_t)c * pk.cout + 0];   // image = output channel 0
    Op op; op.kind = OP_PRED;
    op.in = x; op.skip = skip; op.w = m->upload(w0); op.bias0 = pk.b[0]; op.out = m->out_buf;
    op.N = B.N; op.cin = cin; op.H = H; op.W = W; op.sigmoid = m->cfg.final_sigmoid;
    op.flops = 2.0 * pk.cout * cin * (double)B.N * H * W;
    m->ops[0].push_back(op); m->ops[1].push_back(op);
    return EVK_OK;
}

// UpsampleConvLayer in phase-stacked form (poly.cu): (x + skip) -> replicate-padded split planes, border-ring
// correction, then ONE 5x5 convolution on the low-resolution map with N = 4 * cout whose epilogue scatters the phases.
static int add_poly_decoder(Builder& B, const std::string& pfx, const float* x, const float* skip, int C, int H, int W, float* y) {
    evk_model* m = B.m;
    Packed pk;
    int r = pack_conv(m, pfx + ".conv2d.weight", pfx + ".conv2d.bias", pfx + ".norm_layer", nullptr, pk);
    if (r != EVK_OK) return r;
    EVK_REQUIRE(pk.cin == C && pk.kh == 5 && pk.kw == 5, EVK_ERR_KEY, "'%s': unexpected decoder shape", pfx.c_str());
    const int Co = pk.cout, N = B.N;
    const char* mp = getenv("EVK_MIXED_POLY");
    const int mixed = (m->mixed_enabled && C % 64 == 0 && !(mp && mp[0] == '0')) ? 1 : 0;      // mixed operands (conv.cuh, ConvParams::mixed)
    const size_t plane = (size_t)N * (H + 4) * (W + 4) * C;
    __nv_bfloat16* xp = (__nv_bfloat16*)m->dalloc_bytes(2 * plane * sizeof(__nv_bfloat16));
    // border corrections: u_ext lines just outside the four borders -> two 1x5 tensor-core convolutions (poly.cu)
    const int Ho = 2 * H, Wo = 2 * W, R = 2 * N;
    __nv_bfloat16* lines[2] = {(__nv_bfloat16*)m->dalloc_bytes((size_t)2 * R * (Wo + 4) * C * sizeof(__nv_bfloat16)),
                               (__nv_bfloat16*)m->dalloc_bytes((size_t)2 * R * (Ho + 4) * C * sizeof(__nv_bfloat16))};
    float* ring[2] = {m->dalloc((size_t)R * Wo * 4 * Co), m->dalloc((size_t)R * Ho * 4 * Co)};
    EVK_REQUIRE(xp && lines[0] && lines[1] && ring[0] && ring[1], EVK_ERR_CUDA, "out of device memory for the phase-stacked decoder");
    {
        Op op; op.kind = OP_ADD_PAD;
        op.in = x; op.skip = skip; op.out_s = xp; op.N = N; op.H = H; op.W = W; op.cin = C; op.mixed = mixed;
        m->ops[0].push_back(op); m->ops[1].push_back(op);
    }
    {
        Op op; op.kind = OP_RING_LINES;
        op.out_s = xp; op.lines_h = lines[0]; op.lines_v = lines[1]; op.N = N; op.H = H; op.W = W; op.cin = C; op.mixed = mixed;
        m->ops[0].push_back(op); m->ops[1].push_back(op);
    }
    {
        std::vector<float> wr;
        pack_weights_ring(pk.w.data(), C, Co, wr);
        const float* zero_bias = m->upload(std::vector<float>((size_t)4 * Co, 0.f));
        EVK_REQUIRE(zero_bias != nullptr, EVK_ERR_CUDA, "out of device memory for the ring bias");
        for (int v = 0; v < 2; ++v) {
            const int L = v ? Ho : Wo;
            Op op; op.kind = OP_CONV;
            ConvParams& p = op.cp;
            p.x1 = nullptr; p.c1 = C; p.x1s = lines[v];
            p.N = 1; p.Hin = p.Hout = R; p.Win = L + 4; p.Wout = L; p.kh = 1; p.kw = 5; p.stride = 1; p.pad = 0;
            p.bias = zero_bias; p.cout = 4 * Co; p.epi = EPI_LINEAR; p.act = ACT_NONE; p.y = ring[v];
            std::vector<__nv_bfloat16> wt;
            p.cout_pad = 4 * Co;
            pack_weights_tc(wr.data() + (size_t)v * 5 * C * 4 * Co, 5 * C, 4 * Co, p.cout_pad, wt);
            void* d = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
            EVK_REQUIRE(d != nullptr, EVK_ERR_CUDA, "out of device memory for the ring weights");
            cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
            p.w_tc = (const __nv_bfloat16*)d;
            EVK_REQUIRE(tc_eligible(p), EVK_ERR_ARG, "'%s': border-line convolution is not eligible for the tensor-core path", pfx.c_str());
            op.flops = 0.0;                       // overhead of this formulation, not the reference's arithmetic
            op.ring_line = 1 + v;
            m->ops[0].push_back(op); m->ops[1].push_back(op);
        }
    }
    Op op; op.kind = OP_CONV;
    ConvParams& p = op.cp;
    p.x1 = nullptr; p.c1 = C; p.x1s = xp; p.phase4 = Co == 32 ? 1 : Co == 64 ? 2 : 3; p.ring_h = ring[0]; p.ring_v = ring[1];    // 4 phases in one N tile, one row phase per tile, or one phase per tile
    p.N = N; p.Hin = H + 4; p.Win = W + 4; p.Hout = H; p.Wout = W; p.kh = 5; p.kw = 5; p.stride = 1; p.pad = 0;
    p.bias = m->upload(pk.b); p.cout = Co; p.epi = EPI_LINEAR; p.act = ACT_RELU; p.y = y;
    std::vector<float> wc;
    pack_weights_phase4(pk.w.data(), C, Co, wc);
    std::vector<__nv_bfloat16> wt;
    p.cout_pad = 4 * Co;
    pack_weights_tc(wc.data(), 25 * C, 4 * Co, p.cout_pad, wt);
    void* d = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
    EVK_REQUIRE(d != nullptr, EVK_ERR_CUDA, "out of device memory for the phase-stacked weights");
    cudaMemcpy(d, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
    p.w_tc = (const __nv_bfloat16*)d;
    EVK_REQUIRE(tc_eligible(p), EVK_ERR_ARG, "'%s': phase-stacked decoder is not eligible for the tensor-core path", pfx.c_str());
    if (mixed) {
        std::vector<float> isc;
        pack_weights_mixed(wc.data(), 25 * C, 4 * Co, p.cout_pad, wt, isc);
        void* dm = m->dalloc_bytes(wt.size() * sizeof(__nv_bfloat16));
        EVK_REQUIRE(dm != nullptr, EVK_ERR_CUDA, "out of device memory for the phase-stacked weights");
        cudaMemcpy(dm, wt.data(), wt.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
        p.w_mx = (const __nv_bfloat16*)dm;
        p.w_iscale = m->upload(isc);
        p.mixed = 1;
        m->mixed_convs++;
    }
    op.flops = 2.0 * Co * C * 25.0 * (double)N * (2 * H) * (2 * W);      // the reference's arithmetic (no credit for the zeros)
    op.cin = C; op.H = 2 * H; op.W = 2 * W;
    m->ops[0].push_back(op); m->ops[1].push_back(op);
    return EVK_OK;
}

// DynamicUpsampleLayer after the shared x2 upsample (model/submodules.py:120-127):
// context fusion -> atom generation -> per-pixel dynamic conv -> 1x1 compositional conv -> ReLU.
static int add_hyper_decoder(Builder& B, const std::string& pfx, const float* xu, int C, int h, int w, float* y) {
    evk_model* m = B.m;
    const evk_model_config& c = m->cfg;
    EVK_REQUIRE(h * 4 == c.height && w * 4 == c.width, EVK_ERR_ARG,
                "dynamic decoder: context (%dx%d)/4 does not match decoder resolution %dx%d", c.height, c.width, h, w);
    const std::string gen = pfx + ".dynamic_atom_generation";
    const HostTensor* bases = m->find(gen + ".bases");
    EVK_REQUIRE(bases && bases->shape.size() == 2, EVK_ERR_KEY, "missing '%s.bases'", gen.c_str());
    const int K = (int)bases->shape[0], L = (int)bases->shape[1], ks = c.kernel_size;
    EVK_REQUIRE(L == ks * ks, EVK_ERR_KEY, "'%s.bases' has %d taps, kernel_size is %d", gen.c_str(), L, ks);

    float* ctx = B.act(h, w, 8);
    {
        Op op; op.kind = OP_HYPER_CONTEXT;
        op.hp.N = B.N; op.hp.ev_nchw = m->in_buf; op.hp.prev = m->prev_rec; op.hp.ctx = ctx;
        op.hp.bins = c.num_bins; op.hp.H = c.height; op.hp.W = c.width;
        m->ops[0].push_back(op); m->ops[1].push_back(op);
    }
    // context fusion conv: 6 input channels padded to 8 (zero weight rows)
    Packed pk;
    int r = pack_conv(m, pfx + ".context_fusion.conv.weight", pfx + ".context_fusion.conv.bias", "", nullptr, pk);
    if (r != EVK_OK) return r;
    EVK_REQUIRE(pk.cin == c.num_bins + 1, EVK_ERR_KEY, "context fusion expects %d input channels", c.num_bins + 1);
    Packed pk8 = pk;
    pk8.cin = 8;
    pk8.w.assign((size_t)pk.kh * pk.kw * 8 * pk.cout, 0.f);
    for (int t = 0; t < pk.kh * pk.kw; ++t)
        for (int ci = 0; ci < pk.cin; ++ci)
            for (int n = 0; n < pk.cout; ++n)
                pk8.w[((size_t)t * 8 + ci) * pk.cout + n] = pk.w[((size_t)t * pk.cin + ci) * pk.cout + n];
    int c1 = 0, c2 = 0, c3 = 0;
    float* f1 = B.act(h, w, pk.cout);
    r = B.conv_packed(pk8, ctx, 8, h, w, 1, pk.kh / 2, ACT_NONE, nullptr, f1, &c1);
    if (r != EVK_OK) return r;
    const HostTensor* w0 = m->find(gen + ".bases_net.0.weight");
    const HostTensor* w3 = m->find(gen + ".bases_net.3.weight");
    EVK_REQUIRE(w0 && w3, EVK_ERR_KEY, "missing '%s.bases_net' weights", gen.c_str());
    float* f2 = B.act(h, w, (int)w0->shape[0]);
    r = B.conv(gen + ".bases_net.0.weight", gen + ".bases_net.0.bias", gen + ".bases_net.1", f1, c1, h, w, 1, 1, ACT_TANH, nullptr, f2, &c2);
    if (r != EVK_OK) return r;
    float* f3 = B.act(h, w, (int)w3->shape[0]);
    r = B.conv(gen + ".bases_net.3.weight", gen + ".bases_net.3.bias", gen + ".bases_net.4", f2, c2, h, w, 1, 1, ACT_TANH, nullptr, f3, &c3);
    if (r != EVK_OK) return r;
    EVK_REQUIRE(c3 % K == 0, EVK_ERR_KEY, "basis coefficient channels (%d) not a multiple of the %d bases", c3, K);
    const int A = c3 / K;
    const HostTensor* cc = m->find(pfx + ".dynamic_conv.compositional_coefficients");
    const HostTensor* cb = m->find(pfx + ".dynamic_conv.bias");
    EVK_REQUIRE(cc && cc->shape.size() == 4 && (int)cc->shape[1] == C * A && cc->shape[2] == 1 && cc->shape[3] == 1, EVK_ERR_KEY,
                "'%s.dynamic_conv.compositional_coefficients': expected [Cout, %d, 1, 1]", pfx.c_str(), C * A);
    const int Co = (int)cc->shape[0];
    // Re-associated form (hyper.cuh (3)): the static 1x1 convolution first, on the tensor cores, with the weights regrouped
    // per atom -- U[a*Co + o] = sum_c W[o][c*A + a] * xu[c] -- then the per-pixel 5x5 atoms applied to U.  Taken when the
    // shapes are the shipped ones and the tensor-core path is on; the literal order (atoms -> dynamic conv -> 1x1) otherwise.
    if (c.precision == 0 && A == 6 && K == 12 && ks == 5 && Co == 128 && C % 64 == 0 && getenv("EVK_HYPER_LITERAL") == nullptr) {
        Packed pu;
        pu.cout = A * Co; pu.cin = C; pu.kh = pu.kw = 1;
        pu.w.assign((size_t)C * A * Co, 0.f);
        pu.b.assign((size_t)A * Co, 0.f);
        for (int o = 0; o < Co; ++o)
            for (int ci = 0; ci < C; ++ci)
                for (int a = 0; a < A; ++a)
                    pu.w[(size_t)ci * (A * Co) + (size_t)a * Co + o] = cc->data[(size_t)o * (C * A) + (size_t)ci * A + a];
        float* u = B.act(h, w, A * Co);
        int cu = 0;
        r = B.conv_packed(pu, xu, C, h, w, 1, 0, ACT_NONE, nullptr, u, &cu);
        if (r != EVK_OK) return r;
        m->ops[0].back().flops = 2.0 * Co * (double)(C * A) * (double)B.N * h * w;      // the reference's 1x1 (1536 -> 128): same multiply-adds
        m->ops[1].back().flops = m->ops[0].back().flops;
        std::vector<float> bias(Co, 0.f);
        if (cb) bias = cb->data;
        Op op; op.kind = OP_HYPER_APPLY_U;
        op.hp.N = B.N; op.hp.coef = f3; op.hp.bases = m->upload(bases->data); op.hp.u = u; op.hp.out_bias = m->upload(bias); op.hp.y = y;
        op.hp.h = h; op.hp.w = w; op.hp.A = A; op.hp.K = K; op.hp.L = L; op.hp.ks = ks; op.hp.C = C; op.hp.CO = Co;
        op.flops = 2.0 * A * K * L * (double)B.N * h * w + 2.0 * A * L * C * (double)B.N * h * w;   // atoms + the reference's atom application
        m->ops[0].push_back(op); m->ops[1].push_back(op);
        return EVK_OK;
    }
    float* atoms = B.act(h, w, A * L);
    float* inter = B.act(h, w, C * A);
    {
        Op op; op.kind = OP_HYPER_ATOMS;
        op.hp.N = B.N; op.hp.coef = f3; op.hp.bases = m->upload(bases->data); op.hp.atoms = atoms;
        op.hp.h = h; op.hp.w = w; op.hp.A = A; op.hp.K = K; op.hp.L = L; op.hp.ks = ks;
        op.flops = 2.0 * A * K * L * (double)B.N * h * w;
        m->ops[0].push_back(op); m->ops[1].push_back(op);
        Op ap = op; ap.kind = OP_HYPER_APPLY;
        ap.hp.xu = xu; ap.hp.inter = inter; ap.hp.C = C;
        ap.flops = 2.0 * A * L * C * (double)B.N * h * w;
        m->ops[0].push_back(ap); m->ops[1].push_back(ap);
    }
    return B.conv(pfx + ".dynamic_conv.compositional_coefficients", pfx + ".dynamic_conv.bias", "", inter, C * A, h, w, 1, 0,
                  ACT_RELU, nullptr, y, nullptr);
}

// Parity fix-up: the builders wire every consumer of a ConvLSTM's new hidden state against the parity-0 target (buf[1]); in the
// parity-1 program the same pointers must read buf[0] (that is where parity 1 writes h_new).
static void fix_hidden_parity(evk_model* m, const std::vector<int>& hstate) {
    for (Op& op : m->ops[1]) {
        for (size_t i = 0; i < hstate.size(); ++i) {
            const StateBuf& s = m->states[hstate[i]];
            auto fix = [&](const float*& p) { if (p == s.buf[1]) p = s.buf[0]; };
            if (op.kind == OP_CONV && op.cp.epi == EPI_LSTM) { fix(op.cp.x1); continue; }   // x2/h_new already per parity
            if (op.kind == OP_CONV) { fix(op.cp.x1); fix(op.cp.res); }
            if (op.kind == OP_UPSAMPLE_ADD || op.kind == OP_ZERO_INSERT_ADD || op.kind == OP_PRED || op.kind == OP_ADD_PAD || op.kind == OP_ADD_SPLIT ||
                op.kind == OP_SPADE_PRED || op.kind == OP_LAYERNORM || op.kind == OP_ADD_POS) { fix(op.in); fix(op.skip); }
                if (op.kind == OP_AVG6) for (int k = 0; k < 6; ++k) fix(op.six[k]);
        }
    }
}

// E2VID / E2VID+ / SSL-E2VID / HyperE2VID (model/unet.py:107-143)
static int build_unet(evk_model* m) {
    const evk_model_config& c = m->cfg;
    Builder B{m, c.batch};
    const int E = c.num_encoders, k = c.kernel_size;
    EVK_REQUIRE(E >= 1 && E <= 5, EVK_ERR_ARG, "num_encoders=%d unsupported", E);
    EVK_REQUIRE(c.height % (1 << E) == 0 && c.width % (1 << E) == 0, EVK_ERR_ARG,
                "input %dx%d must be a multiple of 2^num_encoders (CropParameters.pad)", c.height, c.width);
    int H = c.height, W = c.width, C = c.base_channels;
    float* head = B.act(H, W, C);
    int r = add_head(B, "head.conv2d", "head.norm_layer", head, C);
    if (r != EVK_OK) return r;
    const float* x = head;
    std::vector<int> hstate(E);
    for (int i = 0; i < E; ++i) {
        const std::string pfx = "encoders." + std::to_string(i);
        const int Co = 2 * C, Ho = (H + 2 * (k / 2) - k) / 2 + 1, Wo = (W + 2 * (k / 2) - k) / 2 + 1;
        float* xe = B.act(Ho, Wo, Co);
        r = B.conv(pfx + ".conv.conv2d.weight", pfx + ".conv.conv2d.bias", pfx + ".conv.norm_layer", x, C, H, W, 2, k / 2,
                   ACT_RELU, nullptr, xe, nullptr);
        if (r != EVK_OK) return r;
        const int hs = make_state(m, Co, Ho, Wo, true);
        if (hs < 0) return hs;
        const int cs = make_state(m, Co, Ho, Wo, false);
        if (cs < 0) return cs;
        r = add_lstm(B, pfx + ".recurrent_block", xe, Co, Ho, Wo, hs, cs);
        if (r != EVK_OK) return r;
        hstate[i] = hs;
        C = Co; H = Ho; W = Wo;
        // Consumers of the new hidden state are built against the parity-0 target (buf[1]);
        // the parity-1 program is patched after the build (see the fix-up loop below).
        x = m->states[hs].buf[1];
    }
    // residual blocks
    float* tmp = B.act(H, W, C);
    for (int j = 0; j < c.num_residual_blocks; ++j) {
        float* y = B.act(H, W, C);
        r = add_resblock(B, "resblocks." + std::to_string(j), x, tmp, y, C, H, W);
        if (r != EVK_OK) return r;
        x = y;
    }
    // decoders
    for (int i = 0; i < E; ++i) {
        const int e = E - 1 - i;
        const std::string pfx = "decoders." + std::to_string(i);
        const float* skip = m->states[hstate[e]].buf[1];   // placeholder, fixed per parity below
        const bool tconv = m->find(pfx + ".transposed_conv2d.weight") != nullptr;     // use_upsample_conv=False
        // last decoder(s) with cout = 32: four output phases stacked along N on the low-resolution map (poly.cu)
        static const int poly_max_c = getenv("EVK_POLY_MAX_C") ? atoi(getenv("EVK_POLY_MAX_C")) : 256;
        if (!tconv && !(i == 0 && c.dynamic_decoder) && c.precision == 0 && k == 5 && (C == 64 || C == 128 || C == 256) && C <= poly_max_c &&
            H >= 2 && W >= 2 && getenv("EVK_NO_POLY") == nullptr) {
            float* y = B.act(2 * H, 2 * W, C / 2);
            r = add_poly_decoder(B, pfx, x, skip, C, H, W, y);
            if (r != EVK_OK) return r;
            x = y; C /= 2; H *= 2; W *= 2;
            continue;
        }
        float* up = B.act(2 * H, 2 * W, C);
        {
            Op op; op.kind = tconv ? OP_ZERO_INSERT_ADD : OP_UPSAMPLE_ADD;
            op.in = x; op.skip = skip; op.out = up; op.N = B.N; op.H = H; op.W = W; op.cin = C;
            m->ops[0].push_back(op); m->ops[1].push_back(op);
        }
        if (i == 0 && c.dynamic_decoder) {
            float* y = B.act(2 * H, 2 * W, C / 2);
            r = add_hyper_decoder(B, pfx, up, C, 2 * H, 2 * W, y);
            if (r != EVK_OK) return r;
            x = y;
        } else {
            float* y = B.act(2 * H, 2 * W, C / 2);
            if (tconv) {
                // TransposedConvLayer (model/submodules.py:38-66): ConvTranspose2d(k, stride 2, padding k/2, output_padding 1)
                // == stride-1 convolution of the zero-inserted map with the flipped kernel, padding k - 1 - k/2
                r = B.conv(pfx + ".transposed_conv2d.weight", pfx + ".transposed_conv2d.bias", pfx + ".norm_layer", up, C, 2 * H, 2 * W, 1,
                           k - 1 - k / 2, ACT_RELU, nullptr, y, nullptr, true);
            } else {
                EVK_REQUIRE(m->find(pfx + ".conv2d.weight") != nullptr, EVK_ERR_KEY, "'%s.conv2d.weight' / '.transposed_conv2d.weight' missing", pfx.c_str());
                r = B.conv(pfx + ".conv2d.weight", pfx + ".conv2d.bias", pfx + ".norm_layer", up, C, 2 * H, 2 * W, 1, k / 2,
                           ACT_RELU, nullptr, y, nullptr);
            }
            if (r != EVK_OK) return r;
            x = y;
        }
        C /= 2; H *= 2; W *= 2;
    }
    r = add_pred(B, "pred", x, head, C, H, W);
    if (r != EVK_OK) return r;
    fix_hidden_parity(m, hstate);
    return EVK_OK;
}

// SPADE-E2VID (model/spade_e2v.py:113-179, Unet6): fixed widths 32 / 64 / 128 / 256, recurrent encoders rec0 (stride 1, full
// resolution) / rec1 / rec2, two residual blocks, pixel-shuffle decoders up0 / up1 with SPADE normalisation conditioned on the
// previous 3-channel reconstruction, a recurrent last decoder up2, and a 3-channel sigmoid prediction whose mean is the image.
static int build_spade(evk_model* m) {
    const evk_model_config& c = m->cfg;
    Builder B{m, c.batch};
    EVK_REQUIRE(c.num_bins == 5 && c.height % 4 == 0 && c.width % 4 == 0, EVK_ERR_ARG,
                "SPADE-E2VID: 5 bins and an input that is a multiple of 4 (CropParameters pads to 8) are required (got %d bins, %dx%d)", c.num_bins,
                c.height, c.width);
    EVK_REQUIRE(c.precision == 0, EVK_ERR_ARG, "SPADE-E2VID is built on the tensor-core path only (precision 0)");
    const int H = c.height, W = c.width;
    m->prev3 = m->dalloc((size_t)c.batch * 3 * H * W);
    EVK_REQUIRE(m->prev3 != nullptr, EVK_ERR_CUDA, "out of device memory");
    float* head = B.act(H, W, 32);
    int r = add_head(B, "fc", "", head, 32);
    if (r != EVK_OK) return r;
    std::vector<int> hstate;
    // RecurrentConvLayer: conv5x5 (no bias) + BN + ReLU + ConvLSTM
    auto rec = [&](const std::string& pfx, const float* x, int cin, int cout, int h, int w, int stride, const float** out) -> int {
        const int ho = (h + 4 - 5) / stride + 1, wo = (w + 4 - 5) / stride + 1;
        float* y = B.act(ho, wo, cout);
        int rr = B.conv(pfx + ".conv0.weight", "", pfx + ".bn", x, cin, h, w, stride, 2, ACT_RELU, nullptr, y, nullptr);
        if (rr != EVK_OK) return rr;
        const int hs = make_state(m, cout, ho, wo, true);
        if (hs < 0) return hs;
        const int cs = make_state(m, cout, ho, wo, false);
        if (cs < 0) return cs;
        rr = add_lstm(B, pfx + ".recurrent_block", y, cout, ho, wo, hs, cs);
        if (rr != EVK_OK) return rr;
        hstate.push_back(hs);
        *out = m->states[hs].buf[1];
        return EVK_OK;
    };
    const float *x0 = nullptr, *x1 = nullptr, *x2 = nullptr, *x3 = nullptr;
    if ((r = rec("rec0", head, 32, 64, H, W, 1, &x0)) != EVK_OK) return r;
    if ((r = rec("rec1", x0, 64, 128, H, W, 2, &x1)) != EVK_OK) return r;
    if ((r = rec("rec2", x1, 128, 256, H / 2, W / 2, 2, &x2)) != EVK_OK) return r;
    const int H4 = H / 4, W4 = W / 4;
    float* tmp = B.act(H4, W4, 256);
    float* y0 = B.act(H4, W4, 256);
    float* y1 = B.act(H4, W4, 256);
    if ((r = add_resblock(B, "res0", x2, tmp, y0, 256, H4, W4)) != EVK_OK) return r;
    if ((r = add_resblock(B, "res1", y0, tmp, y1, 256, H4, W4)) != EVK_OK) return r;
    auto add_sum = [&](const float* a, const float* b, int C, int h, int w) -> float* {
        float* o = B.act(h, w, C);
        Op op; op.kind = OP_ADD_SPLIT;
        op.in = a; op.skip = b; op.out = o; op.N = B.N; op.H = h; op.W = w; op.cin = C;
        m->ops[0].push_back(op); m->ops[1].push_back(op);
        return o;
    };
    // UpConvLayer3: conv3x3 C -> 4*Co (no bias), PixelShuffle(2), SPADE(Co, 3), ReLU   (h x w = input size)
    auto up = [&](const std::string& pfx, const float* x, int C, int Co, int h, int w, float** out) -> int {
        float* c0 = B.act(h, w, 4 * Co);
        int rr = B.conv(pfx + ".conv0.weight", "", "", x, C, h, w, 1, 1, ACT_NONE, nullptr, c0, nullptr);
        if (rr != EVK_OK) return rr;
        const int ho = 2 * h, wo = 2 * w;
        EVK_REQUIRE(H % ho == 0 && H / ho == W / wo, EVK_ERR_ARG, "SPADE-E2VID: decoder resolution %dx%d does not divide the input", ho, wo);
        const std::string n = pfx + ".norm";
        // segmap = nearest(prev3 -> ho x wo) -> conv3x3 3 -> 64 + ReLU (row-window tensor-core form of a <= 8-channel NCHW input)
        float* actv = B.act(ho, wo, 64);
        rr = add_rowwin_conv(B, n + ".mlp_shared.0", "", m->prev3, H, W, H / ho, ho, wo, ACT_RELU, actv, 64, 3);
        if (rr != EVK_OK) return rr;
        // gamma | beta as ONE convolution 64 -> 2*Co
        Packed pg, pb, pk;
        if ((rr = pack_conv(m, n + ".mlp_gamma.weight", n + ".mlp_gamma.bias", "", nullptr, pg)) != EVK_OK) return rr;
        if ((rr = pack_conv(m, n + ".mlp_beta.weight", n + ".mlp_beta.bias", "", nullptr, pb)) != EVK_OK) return rr;
        EVK_REQUIRE(pg.cout == Co && pb.cout == Co && pg.cin == 64 && pb.cin == 64 && pg.kh == 3 && pb.kh == 3, EVK_ERR_KEY,
                    "'%s': unexpected SPADE gamma / beta shapes", n.c_str());
        pk.cout = 2 * Co; pk.cin = 64; pk.kh = pk.kw = 3;
        pk.w.resize((size_t)9 * 64 * 2 * Co);
        pk.b.resize((size_t)2 * Co);
        for (int k = 0; k < 9 * 64; ++k)
            for (int o = 0; o < Co; ++o) {
                pk.w[(size_t)k * 2 * Co + o] = pg.w[(size_t)k * Co + o];
                pk.w[(size_t)k * 2 * Co + Co + o] = pb.w[(size_t)k * Co + o];
            }
        for (int o = 0; o < Co; ++o) { pk.b[o] = pg.b[o]; pk.b[Co + o] = pb.b[o]; }
        float* gb = B.act(ho, wo, 2 * Co);
        if ((rr = B.conv_packed(pk, actv, 64, ho, wo, 1, 1, ACT_NONE, nullptr, gb, nullptr)) != EVK_OK) return rr;
        const HostTensor* mean = m->find(n + ".param_free_norm.running_mean");
        const HostTensor* var = m->find(n + ".param_free_norm.running_var");
        EVK_REQUIRE(mean && var && (int)mean->data.size() == Co && (int)var->data.size() == Co, EVK_ERR_KEY, "missing '%s.param_free_norm' statistics", n.c_str());
        std::vector<float> alpha(Co), shift(Co);
        for (int o = 0; o < Co; ++o) {
            const double a = 1.0 / std::sqrt((double)var->data[o] + 1e-5);
            alpha[o] = (float)a;
            shift[o] = (float)(-(double)mean->data[o] * a);
        }
        float* o = B.act(ho, wo, Co);
        Op op; op.kind = OP_SPADE_SHUFFLE;
        op.in = c0; op.skip = gb; op.w = m->upload(alpha); op.b = m->upload(shift); op.out = o; op.N = B.N; op.H = h; op.W = w; op.cout = Co;
        EVK_REQUIRE(op.w && op.b, EVK_ERR_CUDA, "out of device memory");
        m->ops[0].push_back(op); m->ops[1].push_back(op);
        *out = o;
        return EVK_OK;
    };
    float *u0 = nullptr, *u1 = nullptr;
    if ((r = up("up0", add_sum(y1, x2, 256, H4, W4), 256, 128, H4, W4, &u0)) != EVK_OK) return r;
    if ((r = up("up1", add_sum(u0, x1, 128, H / 2, W / 2), 128, 64, H / 2, W / 2, &u1)) != EVK_OK) return r;
    if ((r = rec("up2", add_sum(u1, x0, 64, H, W), 64, 32, H, W, 1, &x3)) != EVK_OK) return r;
    {
        Packed pk;
        if ((r = pack_conv(m, "conv_img.weight", "conv_img.bias", "bn_img", nullptr, pk)) != EVK_OK) return r;
        EVK_REQUIRE(pk.cin == 32 && pk.cout == 3 && pk.kh == 1, EVK_ERR_KEY, "'conv_img': expected [3,32,1,1]");
        Op op; op.kind = OP_SPADE_PRED;
        op.in = x3; op.skip = head; op.w = m->upload(pk.w); op.out = m->out_buf; op.prev3 = m->prev3;
        for (int k = 0; k < 3; ++k) op.bias3[k] = pk.b[k];
        op.N = B.N; op.cin = 32; op.H = H; op.W = W;
        op.flops = 2.0 * 3 * 32 * (double)B.N * H * W;
        m->ops[0].push_back(op); m->ops[1].push_back(op);
    }
    fix_hidden_parity(m, hstate);
    return EVK_OK;
}

// ET-Net (model/eitr/u_trans.py:13-123 mls_tpa): E2VID's head, three recurrent stride-2 encoders and upsample-conv decoders around a
// multi-scale token path.  A token tensor [N, L, 256] (L = h*w tokens of the 1/8-resolution map, row-major) IS the NHWC activation
// [N, h, w, 256], so every nn.Linear (attention in / out projections, feed-forward) is a 1x1 convolution on the tensor-core kernel
// with the residual add in its epilogue; LayerNorm, attention, the sine table and the final average are etnet.cu.
static int pack_linear(const evk_model* m, const std::string& wname, const std::string& bname, int row0, int rows, int cin_expected, Packed& out) {
    const HostTensor* w = m->find(wname);
    const HostTensor* b = m->find(bname);
    EVK_REQUIRE(w && b && w->shape.size() == 2 && (int)w->shape[1] == cin_expected && row0 + rows <= (int)w->shape[0] &&
                    (int)b->data.size() == (int)w->shape[0],
                EVK_ERR_KEY, "'%s': missing or unexpected linear layer shape", wname.c_str());
    const int ci = cin_expected;
    out.cout = rows; out.cin = ci; out.kh = out.kw = 1;
    out.w.resize((size_t)ci * rows);
    out.b.resize(rows);
    for (int n = 0; n < rows; ++n) {
        out.b[n] = b->data[row0 + n];
        for (int c = 0; c < ci; ++c) out.w[(size_t)c * rows + n] = w->data[(size_t)(row0 + n) * ci + c];
    }
    return EVK_OK;
}

static int build_etnet(evk_model* m) {
    const evk_model_config& c = m->cfg;
    Builder B{m, c.batch};
    EVK_REQUIRE(c.height % 8 == 0 && c.width % 8 == 0, EVK_ERR_ARG, "ET-Net: input %dx%d must be a multiple of 8 (CropParameters.pad)", c.height, c.width);
    EVK_REQUIRE(m->find("head.norm_layer.weight") == nullptr && m->find("head.norm_layer.running_mean") == nullptr, EVK_ERR_KEY,
                "ET-Net with a normalisation layer is not built (the shipped checkpoint has norm = None)");
    const int D = 256, FF = 1024;
    int H = c.height, W = c.width, C = 32;
    float* head = B.act(H, W, C);
    int r = add_head(B, "head.conv2d", "", head, C);
    if (r != EVK_OK) return r;
    const float* x = head;
    std::vector<int> hstate(3);
    const float* blocks[3];
    int bh[3], bw[3];
    for (int i = 0; i < 3; ++i) {
        const std::string pfx = "DownsampleConv." + std::to_string(i);
        const int Co = 2 * C, Ho = (H + 4 - 5) / 2 + 1, Wo = (W + 4 - 5) / 2 + 1;
        float* xe = B.act(Ho, Wo, Co);
        r = B.conv(pfx + ".conv.conv2d.weight", pfx + ".conv.conv2d.bias", "", x, C, H, W, 2, 2, ACT_RELU, nullptr, xe, nullptr);
        if (r != EVK_OK) return r;
        const int hs = make_state(m, Co, Ho, Wo, true);
        if (hs < 0) return hs;
        const int cs = make_state(m, Co, Ho, Wo, false);
        if (cs < 0) return cs;
        r = add_lstm(B, pfx + ".recurrent_block", xe, Co, Ho, Wo, hs, cs);
        if (r != EVK_OK) return r;
        hstate[i] = hs;
        C = Co; H = Ho; W = Wo;
        x = m->states[hs].buf[1];
        blocks[i] = x; bh[i] = H; bw[i] = W;
    }
    const int h = H, w = W, L = h * w;          // token grid = the 1/8-resolution map
    EVK_REQUIRE(C == D && bh[1] == 2 * h && bw[1] == 2 * w && bh[0] == 4 * h && bw[0] == 4 * w, EVK_ERR_ARG, "ET-Net: unexpected encoder pyramid");
    std::vector<float> pos_host;
    sine_position_table(L, D, pos_host);
    const float* pos = m->upload(pos_host);
    EVK_REQUIRE(pos != nullptr, EVK_ERR_CUDA, "out of device memory");
    auto push = [&](const Op& op) { m->ops[0].push_back(op); m->ops[1].push_back(op); };
    auto tokens = [&](int ch) { return B.act(h, w, ch); };
    auto layernorm = [&](const std::string& pfx, const float* in) -> float* {
        const HostTensor* g = m->find(pfx + ".weight");
        const HostTensor* b = m->find(pfx + ".bias");
        if (!(g && b && (int)g->data.size() == D && (int)b->data.size() == D)) { set_error("'%s': missing LayerNorm(256) parameters", pfx.c_str()); return nullptr; }
        float* o = tokens(D);
        Op op; op.kind = OP_LAYERNORM;
        op.in = in; op.w = m->upload(g->data); op.b = m->upload(b->data); op.out = o; op.N = B.N; op.H = h; op.W = w; op.cin = D;
        push(op);
        return o;
    };
    // y = act(x W^T + b) [+ res] over the tokens: rows [row0, row0 + rows) of an nn.Linear weight
    auto linear = [&](const std::string& wname, const std::string& bname, int row0, int rows, const float* in, int cin, int act, const float* res,
                      float** out) -> int {
        Packed pk;
        int rr = pack_linear(m, wname, bname, row0, rows, cin, pk);
        if (rr != EVK_OK) return rr;
        *out = tokens(rows);
        return B.conv_packed(pk, in, cin, h, w, 1, 0, act, res, *out, nullptr);
    };
    auto attention = [&](const float* q, int qs, const float* k, const float* v, int kvs) -> float* {
        float* o = tokens(D);
        Op op; op.kind = OP_ATTENTION;
        op.in = q; op.skip = k; op.w = v; op.q_stride = qs; op.kv_stride = kvs; op.Lk = L; op.out = o; op.N = B.N; op.H = h; op.W = w; op.cin = D;
        op.flops = 4.0 * (double)B.N * L * L * D;
        push(op);
        return o;
    };
    // x + MultiheadAttention(q_in, kv_in, kv_in): fused q|k|v projection when the inputs coincide (self attention)
    auto mha = [&](const std::string& pfx, const float* xres, const float* q_in, const float* kv_in, float** out) -> int {
        const std::string wi = pfx + ".in_proj_weight", bi = pfx + ".in_proj_bias";
        float* a = nullptr;
        int rr;
        if (q_in == kv_in) {
            float* qkv = nullptr;
            if ((rr = linear(wi, bi, 0, 3 * D, q_in, D, ACT_NONE, nullptr, &qkv)) != EVK_OK) return rr;
            a = attention(qkv, 3 * D, qkv + D, qkv + 2 * D, 3 * D);
        } else {
            float *q = nullptr, *kv = nullptr;
            if ((rr = linear(wi, bi, 0, D, q_in, D, ACT_NONE, nullptr, &q)) != EVK_OK) return rr;
            if ((rr = linear(wi, bi, D, 2 * D, kv_in, D, ACT_NONE, nullptr, &kv)) != EVK_OK) return rr;
            a = attention(q, D, kv, kv + D, 2 * D);
        }
        return linear(pfx + ".out_proj.weight", pfx + ".out_proj.bias", 0, D, a, D, ACT_NONE, xres, out);
    };
    auto ffn = [&](const std::string& pfx, const float* xres, const float* n, float** out) -> int {
        float* f = nullptr;
        int rr = linear(pfx + ".linear1.weight", pfx + ".linear1.bias", 0, FF, n, D, ACT_RELU, nullptr, &f);
        if (rr != EVK_OK) return rr;
        return linear(pfx + ".linear2.weight", pfx + ".linear2.bias", 0, D, f, FF, ACT_NONE, xres, out);
    };
    // words of the three scales: the 1/8 map itself, 2x2 patches of the 1/4 map, 4x4 patches of the 1/2 map (fp32 CUDA-core
    // convolutions on the hidden states: stride = kernel, no padding)
    const float* words[3] = {blocks[2], nullptr, nullptr};
    for (int s = 1; s < 3; ++s) {
        const std::string nm = "split" + std::to_string(s);
        Packed pk;
        if ((r = pack_conv(m, nm + ".weight", nm + ".bias", "", nullptr, pk)) != EVK_OK) return r;
        const int kk = 1 << s, ci = D >> s;
        EVK_REQUIRE(pk.cin == ci && pk.cout == D && pk.kh == kk && pk.kw == kk, EVK_ERR_KEY, "'%s': expected [256,%d,%d,%d]", nm.c_str(), ci, kk, kk);
        float* o = tokens(D);
        Op op; op.kind = OP_CONV;
        ConvParams& p = op.cp;
        p.x1 = blocks[2 - s]; p.c1 = ci; p.N = B.N; p.Hin = bh[2 - s]; p.Win = bw[2 - s]; p.Hout = h; p.Wout = w;
        p.kh = p.kw = kk; p.stride = kk; p.pad = 0; p.w = m->upload(pk.w); p.bias = m->upload(pk.b); p.cout = D;
        p.epi = EPI_LINEAR; p.act = ACT_NONE; p.y = o;
        op.flops = conv_flops(p, D);
        push(op);
        words[s] = o;
    }
    const float* hs[3];
    const float* hc[3];
    for (int s = 0; s < 3; ++s) {              // three pre-norm encoder layers per scale; the sine table is added once
        const std::string enc = "trans_encoder" + std::to_string(s) + ".encoder.layers.";
        float* xs = tokens(D);
        {
            Op op; op.kind = OP_ADD_POS;
            op.in = words[s]; op.w = pos; op.out = xs; op.N = B.N; op.H = h; op.W = w; op.cin = D;
            push(op);
        }
        const float* xcur = xs;
        for (int i = 0; i < 3; ++i) {
            const std::string p = enc + std::to_string(i);
            float* n1 = layernorm(p + ".norm1", xcur);
            if (!n1) return EVK_ERR_KEY;
            float* x1 = nullptr;
            if ((r = mha(p + ".self_attn", xcur, n1, n1, &x1)) != EVK_OK) return r;
            float* n2 = layernorm(p + ".norm2", x1);
            if (!n2) return EVK_ERR_KEY;
            float* x2 = nullptr;
            if ((r = ffn(p, x1, n2, &x2)) != EVK_OK) return r;
            xcur = x2;
        }
        hs[s] = xcur;
    }
    for (int s = 0; s < 3; ++s) {              // two decoder layers per scale, cross attention to the coarser scale's tokens
        const std::string dec = "trans_decoder" + std::to_string(s) + ".decoder.layers.";
        const float* memory = hs[s == 0 ? 0 : s - 1];
        const float* xcur = hs[s];
        for (int i = 0; i < 2; ++i) {
            const std::string p = dec + std::to_string(i);
            float* n1 = layernorm(p + ".norm1", xcur);
            if (!n1) return EVK_ERR_KEY;
            float* x1 = nullptr;
            if ((r = mha(p + ".self_attn", xcur, n1, n1, &x1)) != EVK_OK) return r;
            float* nq = layernorm(p + ".norm21", x1);
            float* nm = layernorm(p + ".norm22", memory);
            if (!nq || !nm) return EVK_ERR_KEY;
            float* x2 = nullptr;
            if ((r = mha(p + ".cross_attn", x1, nq, nm, &x2)) != EVK_OK) return r;
            float* n3 = layernorm(p + ".norm3", x2);
            if (!n3) return EVK_ERR_KEY;
            float* x3 = nullptr;
            if ((r = ffn(p, x2, n3, &x3)) != EVK_OK) return r;
            xcur = x3;
        }
        hc[s] = xcur;
    }
    float* t = tokens(D);
    {
        Op op; op.kind = OP_AVG6;
        for (int s = 0; s < 3; ++s) { op.six[s] = hs[s]; op.six[3 + s] = hc[s]; }
        op.out = t; op.N = B.N; op.H = h; op.W = w; op.cin = D;
        push(op);
    }
    // decoders (UpsampleConvLayer on x + skip) and the prediction layer, as in E2VID
    x = t;
    for (int i = 0; i < 3; ++i) {
        const std::string pfx = "UpsampleConv." + std::to_string(i);
        const float* skip = blocks[2 - i];
        float* y = B.act(2 * H, 2 * W, C / 2);
        if (c.precision == 0 && H >= 2 && W >= 2 && getenv("EVK_NO_POLY") == nullptr) {
            r = add_poly_decoder(B, pfx, x, skip, C, H, W, y);
        } else {
            float* up = B.act(2 * H, 2 * W, C);
            Op op; op.kind = OP_UPSAMPLE_ADD;
            op.in = x; op.skip = skip; op.out = up; op.N = B.N; op.H = H; op.W = W; op.cin = C;
            push(op);
            r = B.conv(pfx + ".conv2d.weight", pfx + ".conv2d.bias", "", up, C, 2 * H, 2 * W, 1, 2, ACT_RELU, nullptr, y, nullptr);
        }
        if (r != EVK_OK) return r;
        x = y; C /= 2; H *= 2; W *= 2;
    }
    r = add_pred(B, "pred", x, head, C, H, W);
    if (r != EVK_OK) return r;
    fix_hidden_parity(m, hstate);
    return EVK_OK;
}

// FireNet_legacy (model/legacy.py:79-111) and FireNet (model/model.py:178-190)
static int build_firenet(evk_model* m, bool legacy) {
    const evk_model_config& c = m->cfg;
    Builder B{m, c.batch};
    const int H = c.height, W = c.width, C = c.base_channels;
    const char* n_head = legacy ? "head.conv.conv2d" : "head.conv2d";
    const char* n_g1 = legacy ? "head.recurrent_block" : "G1";
    const char* n_r1 = legacy ? "resblocks.0.conv" : "R1";
    const char* n_g2 = legacy ? "resblocks.0.recurrent_block" : "G2";
    const char* n_r2 = legacy ? "resblocks.1" : "R2";
    if (c.precision == 0 && C == 16 && W % 2 == 0 && c.kernel_size == 3 && getenv("EVK_NO_WINDOW") == nullptr) m->win_wp = W + 2;
    float* xh = B.act(H, W, C);
    int r = add_head(B, n_head, legacy ? "head.conv.norm_layer" : "head.norm_layer", xh, C);   // (folded when the checkpoint has one: norm = 'BN')
    if (r != EVK_OK) return r;
    float* u = B.act(H, W, C);
    float* hr = B.act(H, W, C);
    float* tmp = B.act(H, W, C);
    const int s1 = make_state(m, C, H, W, false);
    if (s1 < 0) return s1;
    r = add_gru(B, n_g1, xh, C, H, W, s1, u, hr);
    if (r != EVK_OK) return r;
    float* r1 = B.act(H, W, C);
    r = add_resblock(B, n_r1, m->states[s1].buf[0], tmp, r1, C, H, W);
    if (r != EVK_OK) return r;
    const int s2 = make_state(m, C, H, W, false);
    if (s2 < 0) return s2;
    r = add_gru(B, n_g2, r1, C, H, W, s2, u, hr);
    if (r != EVK_OK) return r;
    float* r2 = B.act(H, W, C);
    r = add_resblock(B, n_r2, m->states[s2].buf[0], tmp, r2, C, H, W);
    if (r != EVK_OK) return r;
    return add_pred(B, "pred", r2, nullptr, C, H, W);
}

// Tensor-core wiring: every qualifying convolution gets split-bf16 companions of its inputs (allocated once per
// fp32 buffer), every producer of such a buffer is told to emit the split copy, and the TMA tensor maps are built.
static int wire_tc(evk_model* m) {
    if (m->cfg.precision != 0) return EVK_OK;
    auto need_split = [&](const float* ptr) -> __nv_bfloat16* {
        auto it = m->split_of.find(ptr);
        if (it != m->split_of.end()) return it->second;
        auto sz = m->buf_elems.find(ptr);
        if (sz == m->buf_elems.end()) return nullptr;
        // (window mode: rows padded from W to win_wp pixels, pads stay zero)
        const size_t elems = m->win_wp > 0 ? sz->second / m->cfg.width * m->win_wp : sz->second;
        __nv_bfloat16* d = (__nv_bfloat16*)m->dalloc_bytes(elems * 2 * sizeof(__nv_bfloat16));
        m->split_of[ptr] = d;
        m->split_bytes[ptr] = elems * 2 * sizeof(__nv_bfloat16);
        return d;
    };
    for (int par = 0; par < 2; ++par)
        for (Op& op : m->ops[par]) {
            if (op.kind != OP_CONV || op.cp.w_tc == nullptr || !tc_eligible(op.cp)) continue;
            if (op.cp.x1 != nullptr) op.cp.x1s = need_split(op.cp.x1);   // (the row-window head and the phase-stacked decoder ops bring their own packed input)
            if (op.cp.c2) op.cp.x2s = need_split(op.cp.x2);
            EVK_REQUIRE(op.cp.x1s && (!op.cp.c2 || op.cp.x2s), EVK_ERR_CUDA, "wire_tc: cannot allocate split activations");
        }
    // prediction layer -> epilogue of the tensor-core convolution that produces its input (the epilogue thread of a
    // pixel holds all <= 32 channels): saves the launch and the round trip of the full-resolution decoder output
    if (getenv("EVK_NO_PRED_FUSION") == nullptr)
        for (int par = 0; par < 2; ++par)
            for (size_t i = 0; i + 1 < m->ops[par].size(); ++i) {
                Op& cv = m->ops[par][i];
                Op& pr = m->ops[par][i + 1];
                if (cv.kind != OP_CONV || cv.cp.x1s == nullptr || cv.cp.epi != EPI_LINEAR || cv.cp.cout > 32 || cv.cp.cout % 16 != 0) continue;
                if (pr.kind != OP_PRED || pr.in != cv.cp.y) continue;
                if (cv.cp.win_c > 0) {
                    // window mode (FireNet): a GEMM row is two pixels of 16 channels; fused only in that exact shape, no skip operand
                    if (!(cv.cp.win_c == 16 && cv.cp.kw_group == 2 && cv.cp.cout == 32 && pr.cin == 16 && pr.skip == nullptr) ||
                        getenv("EVK_NO_PRED_FUSION_WIN") != nullptr) continue;
                } else if (pr.cin != cv.cp.cout) continue;
                cv.cp.pred_w = pr.w; cv.cp.pred_skip = pr.skip; cv.cp.pred_out = pr.out;
                cv.cp.pred_bias = pr.bias0; cv.cp.pred_sigmoid = pr.sigmoid;
                cv.flops += pr.flops;
                pr.kind = OP_NOP; pr.flops = 0.0;
            }
    // Mixed operands: a layer runs MIXED when its shape allows it and every companion it reads is in the mixed format; a companion
    // is in the mixed format when every producer of its buffer can write it (tensor-core linear / ConvLSTM epilogues, channel
    // counts multiples of 64) and every tensor-core consumer runs MIXED.  Greatest fixed point, both parities at once.
    if (m->mixed_enabled) {
        std::map<const float*, int> chan;                  // buffers tensor-core convolutions read: channels per pixel
        std::map<const float*, bool> ok;
        auto is_tc = [&](const Op& op) { return op.kind == OP_CONV && op.cp.w_tc != nullptr && tc_eligible(op.cp); };
        for (int par = 0; par < 2; ++par)
            for (const Op& op : m->ops[par]) {
                if (!is_tc(op)) continue;
                if (op.cp.x1 != nullptr) { chan[op.cp.x1] = op.cp.c1; ok[op.cp.x1] = true; }
                if (op.cp.c2) { chan[op.cp.x2] = op.cp.c2; ok[op.cp.x2] = true; }
            }
        // producers
        std::map<const float*, int> produced;
        for (int par = 0; par < 2; ++par)
            for (const Op& op : m->ops[par]) {
                auto out_of = [&](const float* q, bool capable) { if (q && ok.count(q)) { produced[q]++; if (!capable) ok[q] = false; } };
                if (op.kind == OP_CONV) {
                    const ConvParams& p = op.cp;
                    const bool tc = is_tc(op) && (p.x1 == nullptr || true);
                    if (p.epi == EPI_LINEAR) out_of(p.y, tc && p.cout % 64 == 0 && !p.row_pair && p.pred_out == nullptr && !p.kw_group && !p.win_c);
                    else if (p.epi == EPI_LSTM) out_of(p.h_new, tc && (p.cout / 4) % 64 == 0);
                    else { out_of(p.h_new, false); out_of(p.hr_out, false); }
                } else {
                    out_of(op.out, false);
                    out_of(op.hp.inter, false);
                }
            }
        for (auto& kv : ok) if (!produced.count(kv.first)) kv.second = false;       // (a buffer nobody here writes: keep the plain format)
        bool changed = true;
        while (changed) {
            changed = false;
            for (int par = 0; par < 2; ++par)
                for (const Op& op : m->ops[par]) {
                    if (!is_tc(op)) continue;
                    const ConvParams& p = op.cp;
                    const bool in_ok = p.x1 != nullptr && ok[p.x1] && (!p.c2 || ok[p.x2]);
                    if (p.w_mx != nullptr && tc_mixed_capable(p) && in_ok) continue;      // runs MIXED
                    if (p.x1 != nullptr && ok[p.x1]) { ok[p.x1] = false; changed = true; }
                    if (p.c2 && ok[p.x2]) { ok[p.x2] = false; changed = true; }
                }
        }
        for (auto& kv : ok) if (kv.second) m->mixed_bufs[kv.first] = chan[kv.first];
        for (int par = 0; par < 2; ++par)
            for (Op& op : m->ops[par]) {
                if (op.kind != OP_CONV) continue;
                ConvParams& p = op.cp;
                if (is_tc(op) && p.x1 != nullptr && m->mixed_bufs.count(p.x1)) { p.mixed = 1; if (par == 0) m->mixed_convs++; }
                if (p.epi == EPI_LINEAR && p.y && m->mixed_bufs.count(p.y)) p.ys_mixed = 1;
                if (p.epi == EPI_LSTM && p.h_new && m->mixed_bufs.count(p.h_new)) p.hs_mixed = 1;
            }
    }
    auto lookup = [&](const float* ptr) -> __nv_bfloat16* {
        auto it = m->split_of.find(ptr);
        return it == m->split_of.end() ? nullptr : it->second;
    };
    for (int par = 0; par < 2; ++par)
        for (Op& op : m->ops[par]) {
            switch (op.kind) {
                case OP_HEAD: case OP_UPSAMPLE_ADD: case OP_ZERO_INSERT_ADD: case OP_ADD_SPLIT: case OP_SPADE_SHUFFLE: case OP_LAYERNORM: case OP_ATTENTION:
                    op.out_s = lookup(op.out); break;
                case OP_CONV:
                    if (op.cp.epi == EPI_LINEAR) op.cp.ys = lookup(op.cp.y);
                    if (op.cp.epi == EPI_LSTM) op.cp.hs_new = lookup(op.cp.h_new);
                    if (op.cp.epi == EPI_GRU_OUT) op.cp.hs_new = lookup(op.cp.h_new);
                    if (op.cp.epi == EPI_GRU_UR) op.cp.hrs_out = lookup(op.cp.hr_out);
                    if ((op.cp.hs_new || op.cp.hrs_out) && op.cp.x1s == nullptr) {
                        // the fp32 CUDA-core kernel does not write split copies: a tensor-core consumer needs a tensor-core producer
                        set_error("wire_tc: ConvGRU on the CUDA-core path feeds a tensor-core convolution (unsupported channel mix)");
                        return EVK_ERR_ARG;
                    }
                    break;
                case OP_HYPER_APPLY: op.hp.inter_s = lookup(op.hp.inter); break;
                default: break;
            }
            if (op.kind == OP_CONV && m->win_wp > 0 && (op.cp.ys || op.cp.hs_new || op.cp.hrs_out)) {
                op.cp.s_wp = m->win_wp; op.cp.s_c = m->cfg.base_channels; op.cp.s_left = 1;     // row-padded split outputs
            }
        }
    // the fused prediction layer reads its skip operand (the head output) from the split planes when they exist: the
    // head is bound by its output stores (HBM writes), and this drops its fp32 copy
    if (getenv("EVK_PRED_SKIP_FP32") == nullptr)
        for (int par = 0; par < 2; ++par)
            for (Op& op : m->ops[par]) {
                if (op.kind != OP_CONV || op.cp.pred_out == nullptr || op.cp.pred_skip == nullptr || op.cp.x1s == nullptr) continue;
                __nv_bfloat16* sp = lookup(op.cp.pred_skip);
                auto sz = m->buf_elems.find(op.cp.pred_skip);
                if (sp == nullptr || sz == m->buf_elems.end()) continue;
                op.cp.pred_skip_s = sp; op.cp.pred_skip_plane = (long long)sz->second;
                op.cp.pred_skip = nullptr;
            }
    // fp32 copies that no kernel reads (every consumer takes the split-bf16 companion) are not written at all
    {
        std::map<const float*, int> fp32_read;
        for (const StateBuf& sb : m->states) { fp32_read[sb.buf[0]] = 1; fp32_read[sb.buf[1]] = 1; }
        for (int k = 0; k < 2; ++k) { fp32_read[m->out_bufs[k]] = 1; fp32_read[m->in_bufs[k]] = 1; }
        for (int par = 0; par < 2; ++par)
            for (const Op& op : m->ops[par]) {
                auto rd = [&](const float* q) { if (q) fp32_read[q] = 1; };
                switch (op.kind) {
                    case OP_CONV:
                        if (op.cp.x1s == nullptr) { rd(op.cp.x1); rd(op.cp.x2); }
                        rd(op.cp.res); rd(op.cp.c_prev); rd(op.cp.h_prev); rd(op.cp.u_in); rd(op.cp.pred_skip);
                        break;
                    case OP_UPSAMPLE_ADD: case OP_ZERO_INSERT_ADD: case OP_PRED: case OP_ADD_PAD: case OP_ADD_SPLIT: case OP_SPADE_SHUFFLE:
                    case OP_SPADE_PRED: rd(op.in); rd(op.skip); break;
                    case OP_LAYERNORM: case OP_ADD_POS: rd(op.in); break;
                    case OP_ATTENTION: rd(op.in); rd(op.skip); break;        // (v lives in the same projection output as k)
                    case OP_AVG6: for (int k = 0; k < 6; ++k) rd(op.six[k]); break;
                    case OP_HYPER_CONTEXT: case OP_HYPER_ATOMS: case OP_HYPER_APPLY: case OP_HYPER_APPLY_U:
                        rd(op.hp.ev_nchw); rd(op.hp.prev); rd(op.hp.coef); rd(op.hp.atoms); rd(op.hp.xu); rd(op.hp.u);   // (ctx / inter / y are outputs here)
                        break;
                    default: break;
                }
            }
        if (getenv("EVK_KEEP_FP32") == nullptr)
            for (int par = 0; par < 2; ++par)
                for (Op& op : m->ops[par]) {
                    if (op.kind == OP_CONV && op.cp.x1s != nullptr && op.cp.epi == EPI_LINEAR && op.cp.y != nullptr &&
                        !fp32_read.count(op.cp.y) && (op.cp.ys != nullptr || op.cp.pred_out != nullptr)) {
                        op.cp.y = nullptr;
                    }
                    // ConvGRU: h * reset is consumed by the candidate convolution only -- through its split planes on the tensor-core path
                    if (op.kind == OP_CONV && op.cp.x1s != nullptr && op.cp.epi == EPI_GRU_UR && op.cp.hrs_out != nullptr && op.cp.hr_out != nullptr &&
                        !fp32_read.count(op.cp.hr_out)) op.cp.hr_out = nullptr;
                    if ((op.kind == OP_UPSAMPLE_ADD || op.kind == OP_ZERO_INSERT_ADD || op.kind == OP_ADD_SPLIT || op.kind == OP_SPADE_SHUFFLE ||
                         op.kind == OP_LAYERNORM || op.kind == OP_ATTENTION) &&
                        op.out_s != nullptr && !fp32_read.count(op.out)) op.out = nullptr;
                    if (op.kind == OP_HYPER_APPLY && op.hp.inter_s != nullptr && !fp32_read.count(op.hp.inter)) op.hp.inter = nullptr;
                }
    }
    for (int par = 0; par < 2; ++par)
        for (Op& op : m->ops[par]) {
            if (op.kind != OP_CONV || op.cp.x1s == nullptr) continue;
            int r = tc_plan_create(op.cp);
            if (r != EVK_OK) return r;
            m->plans.push_back(op.cp.tc);
            if (par == 0) m->tc_convs++;
        }
    return EVK_OK;
}

// `side` != nullptr (graph capture only): the horizontal border-line convolution of every phase-stacked decoder is issued
// on `side` (forked after the border-line kernel, joined before the decoder's main convolution), so the two small
// correction launches -- ~one tile per SM each, latency bound -- overlap instead of running back to back.
static int run_ops(evk_model* m, int par, cudaStream_t st, std::vector<cudaEvent_t>* ev = nullptr, cudaStream_t side = nullptr) {
    bool forked = false;
    for (const Op& op : m->ops[par]) {
        int r = EVK_OK;
        if (ev) {
            cudaEvent_t e; EVK_CHECK_CUDA(cudaEventCreate(&e)); EVK_CHECK_CUDA(cudaEventRecord(e, st)); ev->push_back(e);
        }
        if (side != nullptr && op.kind == OP_CONV && op.ring_line == 1) {
            EVK_CHECK_CUDA(cudaEventRecord(m->ev_fork, st));
            EVK_CHECK_CUDA(cudaStreamWaitEvent(side, m->ev_fork, 0));
            r = launch_conv(op.cp, m->cfg.precision, side);
            if (r != EVK_OK) return r;
            EVK_CHECK_CUDA(cudaEventRecord(m->ev_join, side));
            forked = true;
            continue;
        }
        if (forked && op.kind == OP_CONV && op.ring_line == 0) {          // the decoder's main convolution reads both corrections
            EVK_CHECK_CUDA(cudaStreamWaitEvent(st, m->ev_join, 0));
            forked = false;
        }
        switch (op.kind) {
            case OP_HEAD: r = launch_head_conv(op.in, op.w, op.b, op.out, op.out_s, op.N, op.cin, op.H, op.W, op.k, op.cout, st); break;
            case OP_CONV: r = launch_conv(op.cp, m->cfg.precision, st); break;
            case OP_UPSAMPLE_ADD: r = launch_upsample2x_add(op.in, op.skip, op.out, op.out_s, op.N, op.H, op.W, op.cin, st); break;
            case OP_ZERO_INSERT_ADD: r = launch_zero_insert2x_add(op.in, op.skip, op.out, op.out_s, op.N, op.H, op.W, op.cin, st); break;
            case OP_NOP: break;
            case OP_ADD_PAD: r = launch_add_pad_split(op.in, op.skip, op.out_s, op.N, op.H, op.W, op.cin, st, op.mixed); break;
            case OP_RING_LINES: r = launch_ring_lines(op.out_s, op.lines_h, op.lines_v, op.N, op.H, op.W, op.cin, st, op.mixed); break;
            case OP_HEAD_PACK: r = launch_head_pack(op.in, op.out_s, op.N, op.cin, op.H, op.W, op.k / 2, st, op.srcH, op.srcW, op.stride, op.src_planes); break;
            case OP_ADD_SPLIT: r = launch_add_split(op.in, op.skip, op.out, op.out_s, (int64_t)op.N * op.H * op.W * op.cin, st); break;
            case OP_SPADE_SHUFFLE: r = launch_spade_shuffle(op.in, op.skip, op.w, op.b, op.out, op.out_s, op.N, op.H, op.W, op.cout, st); break;
            case OP_SPADE_PRED: r = launch_spade_pred(op.in, op.skip, op.w, op.bias3, op.prev3, op.out, op.N, (int64_t)op.H * op.W, op.cin, st); break;
            case OP_LAYERNORM: r = launch_layernorm256(op.in, op.w, op.b, op.out, op.out_s, (int64_t)op.N * op.H * op.W, st); break;
            case OP_ATTENTION: r = launch_attention(op.in, op.q_stride, op.skip, op.w, op.kv_stride, op.out, op.out_s, op.N, op.H * op.W, op.Lk, st); break;
            case OP_ADD_POS: r = launch_add_pos(op.in, op.w, op.out, op.N, (int64_t)op.H * op.W * op.cin, st); break;
            case OP_AVG6: r = launch_avg6(op.six, op.out, (int64_t)op.N * op.H * op.W * op.cin, st); break;
            case OP_PRED: r = launch_pred(op.in, op.skip, op.w, op.bias0, op.out, (int64_t)op.N * op.H * op.W, op.cin, op.sigmoid, st); break;
            default: r = launch_hyper(op.kind - OP_HYPER_CONTEXT, op.hp, st); break;
        }
        if (r != EVK_OK) return r;
    }
    if (ev) {
        cudaEvent_t e; EVK_CHECK_CUDA(cudaEventCreate(&e)); EVK_CHECK_CUDA(cudaEventRecord(e, st)); ev->push_back(e);
    }
    return EVK_OK;
}

static std::string op_desc(const Op& op) {
    char b[160];
    switch (op.kind) {
        case OP_HEAD: snprintf(b, sizeof b, "head conv%dx%d %d->%d @%dx%d", op.k, op.k, op.cin, op.cout, op.H, op.W); break;
        case OP_CONV: {
            const ConvParams& p = op.cp;
            const char* e = p.epi == EPI_LSTM ? "lstm" : p.epi == EPI_GRU_UR ? "gru_ur" : p.epi == EPI_GRU_OUT ? "gru_out" : (p.res ? "linear+res" : "linear");
            if (p.stride_x)
                snprintf(b, sizeof b, "conv%dx5 s%d %d+0->%d %s @%dx%d as %dx3 over pixel pairs (%d-channel K rows) [tcgen05 bf16x3]", p.kh, p.stride, p.c1 / 2, p.cout, e,
                         p.Hout, p.Wout, p.kh, p.c1);
            else if (p.win_c > 0)
                snprintf(b, sizeof b, "conv%dx%d s1 %d+%d->%d %s @%dx%d, window K rows (4 px x 16 ch), %d pixels per GEMM row (N=%d) [tcgen05 bf16x3]", p.kh, p.kw_packed,
                         p.win_c, p.c2 ? p.win_c : 0, p.cout / p.kw_group, e, p.Hout, p.Wout * p.kw_group, p.kw_group, p.cout);
            else if (p.kw_packed)
                snprintf(b, sizeof b, "conv%dx%d s1 %d+0->%d head row-window%s @%dx%d, %d pixels per GEMM row (N=%d) [tcgen05 bf16x3]", p.kh, p.kw_packed,
                         op.cin, p.cout / (p.kw_group > 1 ? p.kw_group : 1), p.pred_out ? "+pred" : "", p.Hout, p.Wout * (p.kw_group > 1 ? p.kw_group : 1),
                         p.kw_group > 1 ? p.kw_group : 1, p.cout);
            else if (op.ring_line)
                snprintf(b, sizeof b, "conv1x5 %d->4x%d %s border correction of the next layer @%dx%d lines [tcgen05 bf16x3]", p.c1, p.cout / 4,
                         op.ring_line == 1 ? "horizontal" : "vertical", p.Hout, p.Wout);
            else if (p.phase4)
                snprintf(b, sizeof b, "conv5x5 s1 %d+0->%d linear%s @%dx%d as 4 stacked phases (N=%d%s) on %dx%d [tcgen05 %s]", p.c1, p.cout,
                         p.pred_out ? "+pred" : "", 2 * p.Hout, 2 * p.Wout, 4 * p.cout, p.phase4 == 2 ? ", 4 of 5 tap rows per tile" : p.phase4 == 3 ? ", 4x4 of 5x5 taps per tile" : "", p.Hout, p.Wout,
                         p.mixed ? "f16+2xf8" : "bf16x3");
            else
                snprintf(b, sizeof b, "conv%dx%d s%d %d+%d->%d %s%s @%dx%d [%s]", p.kh, p.kw, p.stride, p.c1, p.c2, p.cout, e,
                         p.pred_out ? "+pred" : "", p.Hout, p.Wout, p.tc ? (p.mixed ? "tcgen05 f16+2xf8" : "tcgen05 bf16x3") : "simt fp32");
            break;
        }
        case OP_UPSAMPLE_ADD: snprintf(b, sizeof b, "upsample2x_add C=%d @%dx%d", op.cin, 2 * op.H, 2 * op.W); break;
        case OP_ZERO_INSERT_ADD: snprintf(b, sizeof b, "zero_insert2x_add C=%d @%dx%d", op.cin, 2 * op.H, 2 * op.W); break;
        case OP_PRED: snprintf(b, sizeof b, "pred 1x1 %d->1 @%dx%d", op.cin, op.H, op.W); break;
        case OP_NOP: snprintf(b, sizeof b, "(pred 1x1 fused into the previous epilogue)"); break;
        case OP_ADD_PAD: snprintf(b, sizeof b, "add + replicate pad -> split bf16 C=%d @%dx%d", op.cin, op.H, op.W); break;
        case OP_RING_LINES: snprintf(b, sizeof b, "border lines of u_ext -> split bf16 C=%d @%dx%d", op.cin, 2 * op.H, 2 * op.W); break;
        case OP_HEAD_PACK: snprintf(b, sizeof b, "head pack NCHW -> row-window split bf16 @%dx%d", op.H, op.W); break;
        case OP_ADD_SPLIT: snprintf(b, sizeof b, "skip add -> split bf16 C=%d @%dx%d", op.cin, op.H, op.W); break;
        case OP_SPADE_SHUFFLE: snprintf(b, sizeof b, "pixel shuffle x2 + SPADE (BatchNorm, (1+gamma), beta) + ReLU C=%d @%dx%d", op.cout, 2 * op.H, 2 * op.W); break;
        case OP_SPADE_PRED: snprintf(b, sizeof b, "SPADE-E2VID prediction: relu(x + head) -> 1x1 %d->3 + BN + sigmoid, image = mean @%dx%d", op.cin, op.H, op.W); break;
        case OP_LAYERNORM: snprintf(b, sizeof b, "LayerNorm(%d) over %d tokens -> split bf16", op.cin, op.H * op.W); break;
        case OP_ATTENTION: snprintf(b, sizeof b, "attention 8 heads x 32, %d queries x %d keys (fp32, online softmax)", op.H * op.W, op.Lk); break;
        case OP_ADD_POS: snprintf(b, sizeof b, "tokens + sine position table, %d tokens", op.H * op.W); break;
        case OP_AVG6: snprintf(b, sizeof b, "mean of the six token sets (3 encoder + 3 decoder outputs), %d tokens", op.H * op.W); break;
        case OP_HYPER_CONTEXT: snprintf(b, sizeof b, "hyper context x0.25"); break;
        case OP_HYPER_ATOMS: snprintf(b, sizeof b, "hyper atoms A=%d K=%d L=%d @%dx%d", op.hp.A, op.hp.K, op.hp.L, op.hp.h, op.hp.w); break;
        case OP_HYPER_APPLY_U: snprintf(b, sizeof b, "hyper atoms + dynamic conv applied to U = conv1x1(x) (re-associated) C=%d A=%d Cout=%d @%dx%d", op.hp.C, op.hp.A, op.hp.CO, op.hp.h, op.hp.w); break;
        default: snprintf(b, sizeof b, "hyper dynamic conv C=%d A=%d @%dx%d", op.hp.C, op.hp.A, op.hp.h, op.hp.w); break;
    }
    return b;
}

}  // namespace evk

// ------------------------------------------------------------------ C ABI
extern "C" {

int evk_model_create(const evk_model_config* cfg, evk_model** out) {
    EVK_REQUIRE(cfg && out, EVK_ERR_ARG, "evk_model_create: null argument");
    EVK_REQUIRE(cfg->arch >= 0 && cfg->arch <= 4, EVK_ERR_ARG, "evk_model_create: unknown arch %d", cfg->arch);
    EVK_REQUIRE(cfg->batch >= 1 && cfg->height > 0 && cfg->width > 0 && cfg->num_bins > 0 && cfg->base_channels > 0 &&
                    cfg->base_channels % 4 == 0,
                EVK_ERR_ARG, "evk_model_create: bad config (batch=%d %dx%d bins=%d base=%d)", cfg->batch, cfg->height,
                cfg->width, cfg->num_bins, cfg->base_channels);
    int ndev = 0;
    EVK_CHECK_CUDA(cudaGetDeviceCount(&ndev));
    EVK_REQUIRE(ndev > 0, EVK_ERR_CUDA, "evk_model_create: no CUDA device (there is no CPU implementation)");
    evk_model* m = new evk_model();
    m->cfg = *cfg;
    const char* ng = getenv("EVK_NO_GRAPH");
    m->use_graph = !(ng && ng[0] == '1');
    const char* mx = getenv("EVK_MIXED");
    m->mixed_enabled = cfg->precision == 0 && !(mx && mx[0] == '0');
    *out = m;
    return EVK_OK;
}

int evk_model_load_tensor(evk_model* m, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    EVK_REQUIRE(m && name && host_data && ndim >= 0 && ndim <= 8, EVK_ERR_ARG, "evk_model_load_tensor: bad argument");
    EVK_REQUIRE(!m->finalized, EVK_ERR_STATE, "evk_model_load_tensor: model already finalized");
    HostTensor t;
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= shape[i]; }
    t.data.assign(host_data, host_data + n);
    m->sd[name] = std::move(t);
    return EVK_OK;
}

int evk_model_finalize(evk_model* m, void* stream) {
    EVK_REQUIRE(m, EVK_ERR_ARG, "evk_model_finalize: null model");
    EVK_REQUIRE(!m->finalized, EVK_ERR_STATE, "evk_model_finalize: already finalized");
    const evk_model_config& c = m->cfg;
    for (int k = 0; k < 2; ++k) {
        m->in_bufs[k] = m->dalloc((size_t)c.batch * c.num_bins * c.height * c.width);
        m->out_bufs[k] = m->dalloc((size_t)c.batch * c.height * c.width);
        EVK_REQUIRE(m->in_bufs[k] && m->out_bufs[k], EVK_ERR_CUDA, "evk_model_finalize: out of device memory");
    }
    m->in_buf = m->in_bufs[0]; m->out_buf = m->out_bufs[0]; m->prev_rec = m->out_bufs[1];
    static const bool build_timing = getenv("EVK_BUILD_TIMING") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    int r = c.arch == EVK_ARCH_UNET_RECURRENT ? build_unet(m) : c.arch == EVK_ARCH_SPADE_E2VID ? build_spade(m) : c.arch == EVK_ARCH_ETNET ? build_etnet(m)
                                                                                                  : build_firenet(m, c.arch == EVK_ARCH_FIRENET_LEGACY);
    if (r != EVK_OK) return r;
    // parity 1 reads / writes the other input / output buffer (and its "previous reconstruction" is parity 0's output)
    for (Op& op : m->ops[1]) {
        auto swap_io = [&](const float*& q) {
            if (q == m->in_bufs[0]) q = m->in_bufs[1];
            else if (q == m->out_bufs[0]) q = m->out_bufs[1];
            else if (q == m->out_bufs[1]) q = m->out_bufs[0];
        };
        auto swap_out = [&](float*& q) { if (q == m->out_bufs[0]) q = m->out_bufs[1]; };
        swap_io(op.in); swap_out(op.out);
        swap_io(op.hp.ev_nchw); swap_io(op.hp.prev);
    }
    const auto t_built = std::chrono::steady_clock::now();
    r = wire_tc(m);
    if (r != EVK_OK) return r;
    if (build_timing)
        fprintf(stderr, "evk_model_finalize: weight packing + program %.1f ms, tensor-core wiring + plans %.1f ms (%d tensor-core launches)\n",
                std::chrono::duration<double, std::milli>(t_built - t_start).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_built).count(), m->tc_convs);
    for (void* p : m->allocs) EVK_REQUIRE(p != nullptr, EVK_ERR_CUDA, "evk_model_finalize: out of device memory");
    m->flops = 0.0;
    for (const Op& op : m->ops[0]) m->flops += op.flops;
    EVK_CHECK_CUDA(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
    EVK_CHECK_CUDA(cudaStreamCreateWithFlags(&m->cap_stream2, cudaStreamNonBlocking));
    EVK_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    EVK_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    EVK_CHECK_CUDA(cudaDeviceSynchronize());
    m->sd.clear();
    m->finalized = true;
    (void)stream;
    return EVK_OK;
}

int evk_model_reset_states(evk_model* m, void* stream) {
    EVK_REQUIRE(m && m->finalized, EVK_ERR_STATE, "evk_model_reset_states: model not finalized");
    cudaStream_t st = (cudaStream_t)stream;
    for (const StateBuf& s : m->states) {
        const size_t bytes = sizeof(float) * (size_t)m->cfg.batch * s.H * s.W * s.C;
        EVK_CHECK_CUDA(cudaMemsetAsync(s.buf[0], 0, bytes, st));
        if (s.pingpong) EVK_CHECK_CUDA(cudaMemsetAsync(s.buf[1], 0, bytes, st));
        for (int k = 0; k < 2; ++k) {
            auto it = m->split_of.find(s.buf[k]);
            if (it != m->split_of.end()) EVK_CHECK_CUDA(cudaMemsetAsync(it->second, 0, m->split_bytes[s.buf[k]], st));
        }
    }
    for (int k = 0; k < 2; ++k)      // (HyperE2VID reads the other output buffer as the previous reconstruction: zeros after a reset)
        EVK_CHECK_CUDA(cudaMemsetAsync(m->out_bufs[k], 0, sizeof(float) * (size_t)m->cfg.batch * m->cfg.height * m->cfg.width, st));
    m->parity = 0;
    m->spade_first = true;
    return EVK_OK;
}

int evk_model_forward(evk_model* m, const float* voxel, float* image, void* stream) {
    EVK_REQUIRE(m && m->finalized, EVK_ERR_STATE, "evk_model_forward: model not finalized");
    EVK_REQUIRE(voxel && image, EVK_ERR_ARG, "evk_model_forward: null tensor");
    cudaStream_t st = (cudaStream_t)stream;
    const evk_model_config& c = m->cfg;
    const size_t in_bytes = sizeof(float) * (size_t)c.batch * c.num_bins * c.height * c.width;
    const size_t out_bytes = sizeof(float) * (size_t)c.batch * c.height * c.width;
    const int par = m->parity;
    if (voxel != m->in_bufs[par]) EVK_CHECK_CUDA(cudaMemcpyAsync(m->in_bufs[par], voxel, in_bytes, cudaMemcpyDeviceToDevice, st));
    if (c.arch == EVK_ARCH_SPADE_E2VID && m->spade_first) {
        // no previous reconstruction yet: x_org = the first three bins, shifted / scaled to [0, 1] in place (spade_e2v.py:140-145)
        int r = launch_spade_first_frame(m->in_bufs[par], m->prev3, c.batch, c.num_bins, (int64_t)c.height * c.width, st);
        if (r != EVK_OK) return r;
        m->spade_first = false;
    }
    if (m->use_graph) {
        if (!m->graph[par]) {
            cudaGraph_t g = nullptr;
            EVK_CHECK_CUDA(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
            static const bool fork_ok = getenv("EVK_NO_GRAPH_FORK") == nullptr;
            int r = run_ops(m, par, m->cap_stream, nullptr, fork_ok ? m->cap_stream2 : nullptr);
            cudaError_t e = cudaStreamEndCapture(m->cap_stream, &g);
            if (r != EVK_OK) { if (g) cudaGraphDestroy(g); return r; }
            EVK_CHECK_CUDA(e);
            EVK_CHECK_CUDA(cudaGraphInstantiate(&m->graph[par], g, 0));
            cudaGraphDestroy(g);
        }
        EVK_CHECK_CUDA(cudaGraphLaunch(m->graph[par], st));
    } else {
        int r = run_ops(m, par, st);
        if (r != EVK_OK) return r;
    }
    m->last_launches = 0;
    for (const Op& op : m->ops[par]) m->last_launches += op.kind != OP_NOP;
    if (image != m->out_bufs[par]) EVK_CHECK_CUDA(cudaMemcpyAsync(image, m->out_bufs[par], out_bytes, cudaMemcpyDeviceToDevice, st));
    m->parity ^= 1;
    return EVK_OK;
}

int evk_model_profile(evk_model* m, const float* voxel, float* image, void* stream, int max_ops, float* ms, double* flops,
                      int* n_ops) {
    EVK_REQUIRE(m && m->finalized, EVK_ERR_STATE, "evk_model_profile: model not finalized");
    EVK_REQUIRE(voxel && image && ms && flops && n_ops, EVK_ERR_ARG, "evk_model_profile: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const evk_model_config& c = m->cfg;
    const size_t in_bytes = sizeof(float) * (size_t)c.batch * c.num_bins * c.height * c.width;
    const size_t out_bytes = sizeof(float) * (size_t)c.batch * c.height * c.width;
    const int par = m->parity;
    if (voxel != m->in_bufs[par]) EVK_CHECK_CUDA(cudaMemcpyAsync(m->in_bufs[par], voxel, in_bytes, cudaMemcpyDeviceToDevice, st));
    if (c.arch == EVK_ARCH_SPADE_E2VID && m->spade_first) {
        int r0 = launch_spade_first_frame(m->in_bufs[par], m->prev3, c.batch, c.num_bins, (int64_t)c.height * c.width, st);
        if (r0 != EVK_OK) return r0;
        m->spade_first = false;
    }
    std::vector<cudaEvent_t> ev;
    int r = run_ops(m, par, st, &ev);
    if (image != m->out_bufs[par]) EVK_CHECK_CUDA(cudaMemcpyAsync(image, m->out_bufs[par], out_bytes, cudaMemcpyDeviceToDevice, st));
    cudaError_t se = cudaStreamSynchronize(st);
    const int n = (int)m->ops[par].size();
    *n_ops = n;
    if (r == EVK_OK && se == cudaSuccess && (int)ev.size() == n + 1)
        for (int i = 0; i < n && i < max_ops; ++i) {
            cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]);
            flops[i] = m->ops[par][i].flops;
        }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    if (r != EVK_OK) return r;
    EVK_CHECK_CUDA(se);
    m->last_launches = 0;
    for (const Op& op : m->ops[par]) m->last_launches += op.kind != OP_NOP;
    m->parity ^= 1;
    return EVK_OK;
}

int evk_model_op_desc(evk_model* m, int index, char* buf, int cap) {
    EVK_REQUIRE(m && m->finalized && buf && cap > 0 && index >= 0 && index < (int)m->ops[0].size(), EVK_ERR_ARG,
                "evk_model_op_desc: bad argument");
    const std::string d = op_desc(m->ops[0][index]);
    snprintf(buf, (size_t)cap, "%s", d.c_str());
    return EVK_OK;
}

int evk_model_io_buffers(evk_model* m, float** in, float** out) {
    EVK_REQUIRE(m && m->finalized, EVK_ERR_STATE, "evk_model_io_buffers: model not finalized");
    if (in) *in = m->in_bufs[m->parity];
    if (out) *out = m->out_bufs[m->parity];
    return EVK_OK;
}

int evk_model_num_states(evk_model* m) { return m ? (int)m->states.size() : EVK_ERR_ARG; }

int evk_model_state_shape(evk_model* m, int index, int64_t shape[4]) {
    EVK_REQUIRE(m && index >= 0 && index < (int)m->states.size(), EVK_ERR_ARG, "evk_model_state_shape: bad index %d", index);
    const StateBuf& s = m->states[index];
    shape[0] = m->cfg.batch; shape[1] = s.C; shape[2] = s.H; shape[3] = s.W;
    return EVK_OK;
}

int evk_model_last_launch_count(evk_model* m) { return m ? m->last_launches : EVK_ERR_ARG; }
double evk_model_flops(evk_model* m) { return m ? m->flops : 0.0; }

int evk_model_destroy(evk_model* m) {
    if (!m) return EVK_OK;
    for (int i = 0; i < 2; ++i)
        if (m->graph[i]) cudaGraphExecDestroy(m->graph[i]);
    if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
    if (m->cap_stream2) cudaStreamDestroy(m->cap_stream2);
    if (m->ev_fork) cudaEventDestroy(m->ev_fork);
    if (m->ev_join) cudaEventDestroy(m->ev_join);
    for (TcPlan* pl : m->plans) tc_plan_destroy(pl);
    m->arena.release();
    delete m;
    return EVK_OK;
}

}  // extern "C"

// state get/set need a layout transpose (NHWC <-> NCHW)
namespace evk {
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int C, int HW) {
    const int64_t total = (int64_t)N * C * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(i % HW);
        const int c = (int)((i / HW) % C);
        const int n = (int)(i / ((int64_t)HW * C));
        out[i] = in[((size_t)n * HW + p) * C + c];
    }
}
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int C, int HW) {
    const int64_t total = (int64_t)N * C * HW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int p = (int)((i / C) % HW);
        const int n = (int)(i / ((int64_t)HW * C));
        out[i] = in[((size_t)n * C + c) * HW + p];
    }
}
}  // namespace evk

extern "C" {

int evk_model_get_state(evk_model* m, int index, float* out_nchw, void* stream) {
    EVK_REQUIRE(m && m->finalized && index >= 0 && index < (int)m->states.size() && out_nchw, EVK_ERR_ARG, "evk_model_get_state: bad argument");
    const StateBuf& s = m->states[index];
    const float* cur = s.pingpong ? s.buf[m->parity] : s.buf[0];   // buffer the next forward will read
    const int64_t total = (int64_t)m->cfg.batch * s.C * s.H * s.W;
    nhwc_to_nchw_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1184), 256, 0, (cudaStream_t)stream>>>(cur, out_nchw, m->cfg.batch, s.C, s.H * s.W);
    EVK_CHECK_CUDA(cudaGetLastError());
    return EVK_OK;
}

int evk_model_set_state(evk_model* m, int index, const float* in_nchw, void* stream) {
    EVK_REQUIRE(m && m->finalized && index >= 0 && index < (int)m->states.size() && in_nchw, EVK_ERR_ARG, "evk_model_set_state: bad argument");
    const StateBuf& s = m->states[index];
    float* cur = s.pingpong ? s.buf[m->parity] : s.buf[0];
    const int64_t total = (int64_t)m->cfg.batch * s.C * s.H * s.W;
    nchw_to_nhwc_kernel<<<(unsigned)std::min<int64_t>(ceil_div64(total, 256), 1184), 256, 0, (cudaStream_t)stream>>>(in_nchw, cur, m->cfg.batch, s.C, s.H * s.W);
    EVK_CHECK_CUDA(cudaGetLastError());
    auto it = m->split_of.find(cur);
    if (it != m->split_of.end()) {
        auto mx = m->mixed_bufs.find(cur);
        if (mx != m->mixed_bufs.end()) return launch_split_mixed(cur, it->second, (int64_t)m->cfg.batch * s.H * s.W, mx->second, (cudaStream_t)stream);
        if (m->win_wp > 0) return launch_split_padded(cur, it->second, (int64_t)m->cfg.batch * s.H, s.W, s.C, m->win_wp, 1, (cudaStream_t)stream);
        return launch_split(cur, it->second, total, (cudaStream_t)stream);
    }
    return EVK_OK;
}

}  // extern "C"
