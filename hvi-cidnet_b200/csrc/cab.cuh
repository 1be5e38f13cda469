#pragma once
#include "common.cuh"

namespace cidnet {

// One launch handles up to two independent CAB problems (the I_LCA / HV_LCA pair of a stage).
struct CabDwArgs {
    const act_t* q[2]; int q_pitch[2];          // q_pre (output of the folded q 1x1), channel 0 of q
    const act_t* k[2]; const act_t* v[2]; int kv_pitch[2];
    const float* wq[2]; const float* wk[2]; const float* wv[2];   // depthwise weights fp32 [9][Cp]
    act_t* v_out[2]; int v_pitch;
    float* gram[2];                              // [B][heads][18][18] fp32, pre-zeroed
    float* sq[2]; float* sk[2];                  // [B][Cp] fp32, pre-zeroed
    int B, H, W, C, Cp, heads, nprob;
    int tiles_x, tiles_y;
};
int launch_cab_dw_gram(CabDwArgs a, cudaStream_t stream);

struct CabFoldArgs {
    const float* gram[2]; const float* sq[2]; const float* sk[2];
    const float* temp[2];                        // [heads]
    const float* wo[2];                          // project_out [C][C] fp32
    act_t* m_out[2];                             // [B][n_rows][kt] packed per-image weights
    int B, C, Cp, heads, nprob, n_rows, kt;
};
int launch_cab_fold(const CabFoldArgs& a, cudaStream_t stream);

}  // namespace cidnet
