#pragma once
#include "common.cuh"

namespace cidnet {

// SpatialAttention of the MSSA variant (/root/reference/net/CIDNet_MSSA.py:10-25), applied IN PLACE to up to two
// NHWC tensors of one level (the I and the HV output of an up-block pair):
//   stats: per pixel (mean over C, max over C), fp32 [B*H*W][2]     (written by the up block's GEMM epilogue)
//   gate : x *= sigmoid(conv7x7([mean, max]), zero padding 3)     (2 -> 1 channels, no bias)
struct SaArgs {
    act_t* x[2];
    const float* w[2];      // fp32 [2][49]: taps of the mean plane, then of the max plane
    float2* stats[2];       // scratch, B*H*W entries each
    int B, H, W, C, pitch, nprob;
};
int launch_sa_gate(const SaArgs& a, cudaStream_t stream);

}  // namespace cidnet
