#pragma once
#include "common.cuh"

namespace cidnet {

// depthwise 3x3 over the [q | k | v] pre-activations of up to two CAB problems
struct Dw3Args {
    const act_t* src[2][3];     // per problem: q_pre, k_pre, v_pre (channel 0 of each segment)
    int src_pitch;
    act_t* dst[2];              // per problem: [q | k | v] after the depthwise conv, pitch dst_pitch
    int dst_pitch;
    const float* w[2];          // fp32 [9][nv*8] tap major, segments in the same order
    float* sq[2]; float* sk[2]; // [B][Cp] sum of squares of q / k (pre-zeroed)
    int B, H, W, nv, seg_vecs, nprob;
    int stat_y0, stat_y1;       // rows whose squares enter sq / sk (row-strip sharding: the owned rows); 0,0 = all
};
int launch_dw3(const Dw3Args& a, cudaStream_t stream);

struct GramLaunch {
    const act_t* q[2]; const act_t* k[2];   // NHWC, channel 0 of q / k, common pitch
    int pitch;
    float* gram[2];                          // [B][heads][18][18] fp32, pre-zeroed
    int B, H, W, C, heads, nprob;
    long long img_stride_px;                 // pixels between images (0 = H*W; larger when H covers only the owned rows)
};
int launch_gram(const GramLaunch& L, cudaStream_t stream);

struct CabFoldArgs {
    const float* gram[2]; const float* sq[2]; const float* sk[2];
    const float* temp[2];                        // [heads]
    const float* wo[2];                          // project_out [C][C] fp32
    act_t* m_out[2];                             // [B][n_rows][kt] packed per-image weights
    int B, C, Cp, heads, nprob, n_rows, kt;
};
int launch_cab_fold(const CabFoldArgs& a, cudaStream_t stream);

}  // namespace cidnet
