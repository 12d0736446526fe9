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
};

// which: 0 context, 1 atoms, 2 apply
int launch_hyper(int which, const HyperParams& p, cudaStream_t st);

}  // namespace evk
