#pragma once
#include "common.cuh"

namespace cidnet {

// depthwise 3x3 over the [q | k | v] pre-activations of up to two CAB problems
struct Dw3Args {
    const act_t* src[2][3];     // per problem: q_pre, k_pre, v_pre (channel 0 of each segment)
    int src_pitch;
    act_t* dst_qk[2];           // per problem: [q | k] after the depthwise conv, pitch 2 * Cp
    act_t* dst_v[2];            // per problem: v after the depthwise conv, pitch Cp
    const float* w[2];          // fp32 [9][nv*8] tap major, segments in the same order
    int B, H, W, nv, seg_vecs, nprob;
    int rows_per_cta;           // set by launch_dw3 (pick_strip_rows)
};
int launch_dw3(const Dw3Args& a, cudaStream_t stream);

// Split-K Gram.  Every CTA writes ONE slab entry of E = heads * 324 + 2 * Cp floats:
//   [heads][18][18] partial Gram blocks | [Cp] partial sum q^2 | [Cp] partial sum k^2
// at slab[((prob * B + b) * nsplit + split) * E]; cab_fold_kernel sums the nsplit entries in a fixed order.
static constexpr int kGramMaxCtas = 480;     // upper bound of the one-wave split-K grid (<= 160 SMs x 3 CTAs)
int gram_max_slab_entries(int nprob_times_B);
struct GramLaunch {
    const act_t* q[2]; const act_t* k[2];   // NHWC, channel 0 of q / k, common pitch
    int pitch;
    float* slab;
    int B, H, W, C, heads, nprob;
    long long img_stride_px;                 // pixels between images (0 = H*W; larger when H covers only the owned rows)
};
int launch_gram(const GramLaunch& L, cudaStream_t stream, int* nsplit_out);

struct CabFoldArgs {
    const float* slab; int nsplit;               // see GramLaunch (nsplit = 1: an already reduced vector per problem / image)
    float* raw_out;                              // optional: the reduced [Gram | sq | sk] per (problem, image), E floats each
    const float* temp[2];                        // [heads]
    const float* wo[2];                          // project_out [C][C] fp32
    act_t* m_out[2];                             // [B][n_rows][kt] packed per-image weights
    int B, C, Cp, heads, nprob, n_rows, kt;
};
int launch_cab_fold(const CabFoldArgs& a, cudaStream_t stream);
// slab -> out[nvec][E] (nvec = nprob * B vectors): fixed-order sum over the nsplit entries (row-strip sharding)
int launch_cab_reduce(const float* slab, float* out, int nsplit, int E, int nvec, cudaStream_t stream);

}  // namespace cidnet
