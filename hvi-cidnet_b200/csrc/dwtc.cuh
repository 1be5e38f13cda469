#pragma once
#include "common.cuh"
#include <vector>

namespace cidnet {

// diagonal weight tiles of a depthwise 3x3 conv over `channels` contiguous channels
struct DwtcWeights { act_t* w = nullptr; int channels = 0; int nblk = 0; };
int pack_dwtc_weights(DwtcWeights* out, const float* w_tapmajor, int pitch, int c_begin, int channels);

struct DwtcSeg {
    const act_t* in = nullptr; int in_pitch = 0;      // channel 0 of the segment, NHWC
    act_t* out = nullptr; int out_pitch = 0;
    const DwtcWeights* wt = nullptr;
    float* ssq = nullptr; int ssq_pitch = 0; int ssq_channels = 0;   // optional per-channel sum of squares [B][pitch]
};
struct DwtcLaunch { DwtcSeg seg[6]; int nseg = 0; int B = 0, H = 0, W = 0; };
int launch_dwtc(const DwtcLaunch& L, cudaStream_t stream);

}  // namespace cidnet
