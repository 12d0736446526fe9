// HyperE2VID dynamic decoder pieces that are not plain convolutions.
// Reference semantics: model/hyper/hyper_dynamic.py:7-92, model/submodules.py:100-127.
#pragma once
#include <cuda_bf16.h>

#include "evk_common.cuh"

namespace evk {

struct HyperParams {
    int N = 0;
    // (0) context: cat(event tensor, previous reconstruction) -> bilinear x0.25 -> NHWC [N,H/4,W/4,8] (6 used)
    const float* ev_nchw = nullptr; const float* prev = nullptr; float* ctx = nullptr;
    int bins = 0, H = 0, W = 0;
    // (1) atoms: coef [N,h,w,A*K] (tanh'd basis coefficients, channel = a*K + k) x bases [K][L] -> atoms [N,h,w,A*L]
    const float* coef = nullptr; const float* bases = nullptr; float* atoms = nullptr;
    int h = 0, w = 0, A = 0, K = 0, L = 0, ks = 0;
    // (2) apply: inter[n,y,x,c*A+a] = sum_l atoms[n,y,x,a,l] * xu[n, y+dy(l), x+dx(l), c]   (zero padded)
    const float* xu = nullptr; float* inter = nullptr; int C = 0;
    __nv_bfloat16* inter_s = nullptr;   // optional split-bf16 copy of inter for the tensor-core 1x1 conv
    // (3) re-associated form (the shipped path): the static 1x1 compositional convolution runs FIRST on the tensor cores,
    //   U[n,y,x,a*CO+o] = sum_c W[o, c*A+a] * xu[n,y,x,c]                               (an ordinary conv_tc launch),
    // and the dynamic part is applied to its output:
    //   y[n,y,x,o] = relu(bias[o] + sum_a sum_l atoms[n,y,x,a,l] * U[n, y+dy(l), x+dx(l), a*CO+o]),
    //   atoms[.,a,l] = sum_k coef[., a*K+k] * bases[k][l]    (computed on the fly)
    // -- the same sums as (1)+(2)+1x1 in a different order: half the dynamic multiply-adds (A*L*CO instead of A*L*C per
    // pixel), and the [N,h,w,C*A] intermediate (1536 channels, split-bf16) is never built.
    const float* u = nullptr; const float* out_bias = nullptr; float* y = nullptr; int CO = 0;
};

// which: 0 context, 1 atoms, 2 apply, 3 apply to U (re-associated)
int launch_hyper(int which, const HyperParams& p, cudaStream_t st);

}  // namespace evk
