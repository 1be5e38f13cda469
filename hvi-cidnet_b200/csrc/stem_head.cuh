#pragma once
#include "common.cuh"

namespace cidnet {

struct StemArgs {
    const void* rgb;     // fp32 [B,3,H,W] planar, or (in_u8) uint8 [B,h_src,w_src,3]
    float* hvi; act_t* i_enc0; act_t* hv_0;
    const float* w_hv;   // [27][36] fp32 (input-tap major)
    const float* w_i;    // [9][36]
    const float* k_dev;  // optional: density_k read on the device (no host sync)
    float k_host;
    int B, H, W, pitch;
    const uint2* bfrag;  // [3][5][32] per-lane MMA B fragments (pack_stem_bfrag)
    int in_u8 = 0, h_src = 0, w_src = 0; float gamma = 1.0f;    // 8-bit input: ToTensor + reflect pad + gamma fused into the load
};
int launch_stem(const StemArgs& a, cudaStream_t stream);

struct HeadArgs {
    const act_t* i_dec1; const act_t* hv_1; const float* hvi;
    void* rgb;           // fp32 [B,3,H,W] planar, or (out_u8) uint8 [B,h_dst,w_dst,3]
    float* out_hvi_dbg;  // optional fp32 NCHW tap of output_hvi
    const float* w_i;    // [9][36]
    const float* w_hv;   // [2][9][36]
    const float* k_dev; float k_host; float alpha_s; float alpha; int gated; int gated2;
    int B, H, W, pitch;
    const uint2* bfrag;  // [5][3][32] per-lane MMA B fragments (pack_head_bfrag)
    int out_u8 = 0, h_dst = 0, w_dst = 0;                       // 8-bit output: clamp + crop + quantise fused into the store
};
int launch_head(const HeadArgs& a, cudaStream_t stream);

// kernel entry addresses (CUDA-graph node identification: argument 0 of the stem is the input image,
// argument 3 of the head is the output image)
const void* stem_kernel_func();
const void* head_kernel_func();
static constexpr int kStemNumArgs = 16, kStemArgIn = 0;
static constexpr int kHeadNumArgs = 16, kHeadArgOut = 3;

// host-side packing of the MMA B fragments (weights as uploaded by api.cu: tap-input major fp32)
void pack_stem_bfrag(const float* w_hv, const float* w_i, uint2* out);
void pack_head_bfrag(const float* w_i, const float* w_hv, uint2* out);

}  // namespace cidnet
