#pragma once
#include "common.cuh"

namespace cidnet {

// IEL gate (net/LCA.py:61-65): t = project_in output, channels [x1 (hp) | x2 (hp)], hp = round_up(h, 16)
//   d  = dwconv(t)            (3x3 depthwise, zero pad)
//   x1 = tanh(dwconv1(d1)) + d1 ;  x2 = tanh(dwconv2(d2)) + d2 ;  g = x1 * x2     -> [P][hp]
struct IelGateArgs {
    const act_t* t[2]; act_t* g[2];
    const float* w0[2];      // dwconv  fp32 [9][2*hp]  (tap major; x1 channels then x2 channels)
    const float* w1[2];      // dwconv1 fp32 [9][hp]
    const float* w2[2];      // dwconv2 fp32 [9][hp]
    int B, H, W, hp, nprob;
    int rows_per_cta;        // set by launch_iel_gate (pick_strip_rows)
};
int launch_iel_gate(const IelGateArgs& a, cudaStream_t stream);

}  // namespace cidnet
