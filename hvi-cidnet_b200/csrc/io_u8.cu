// 8-bit image I/O around the forward: the reference's callers convert, pad, gamma-correct,
// clamp, crop and quantise on the host / with a handful of ATen ops per image
//   pre : transforms.ToTensor (u8 HWC -> fp32 CHW / 255)  -> F.pad(..., (0,padw,0,padh), 'reflect') to a
//         multiple of 8 (data/eval_sets.py:22-27, demo.py:47-52) -> input ** gamma (eval.py:64, demo.py:57)
//   post: torch.clamp(output, 0, 1) (eval.py:69) -> output[:, :, :h, :w] (eval.py:71) ->
//         transforms.ToPILImage = mul(255).byte() (truncation) -> HWC u8
// Here each direction is ONE coalesced elementwise kernel; uint8 in/out cuts the host<->device traffic of
// an image from 24 to 6 bytes per pixel.
#include "common.cuh"

namespace cidnet {

__global__ void __launch_bounds__(256)
pre_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, int h, int w, int H, int W, float gamma,
              int apply_gamma) {
    const int b = blockIdx.z;
    const int y = blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= W) return;
    const int sy = y < h ? y : 2 * (h - 1) - y;         // 'reflect': padded row h+i mirrors row h-2-i
    const int sx = x < w ? x : 2 * (w - 1) - x;
    const uint8_t* p = src + (((long long)b * h + sy) * w + sx) * 3;
    const long long HW = (long long)H * W;
    float* o = dst + (long long)b * 3 * HW + (long long)y * W + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = __fdiv_rn((float)p[c], 255.0f);       // ToTensor: .to(float32).div(255)
        if (apply_gamma) v = powf(v, gamma);
        o[c * HW] = v;
    }
}

__global__ void __launch_bounds__(256)
post_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, int H, int W) {
    const int b = blockIdx.z;
    const int y = blockIdx.y;
    const int x = blockIdx.x * 256 + threadIdx.x;
    if (x >= w) return;
    const long long HW = (long long)H * W;
    const float* p = src + (long long)b * 3 * HW + (long long)y * W + x;
    uint8_t* o = dst + (((long long)b * h + y) * w + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = p[c * HW];
        v = fminf(fmaxf(v, 0.0f), 1.0f);                // clamp(0, 1); NaN -> 0 like the byte() cast of the reference path
        o[c] = (uint8_t)__float2int_rz(__fmul_rn(v, 255.0f));   // mul(255).byte(): truncation
    }
}

}  // namespace cidnet

using namespace cidnet;

extern "C" int cidnet_pre_u8(const uint8_t* src_hwc, float* dst_nchw, int B, int h, int w, int H, int W, float gamma,
                             void* stream) {
    CIDNET_CHECK(B >= 0 && h > 0 && w > 0 && H >= h && W >= w, CIDNET_ERR_INVALID, "pre_u8: bad shape");
    CIDNET_CHECK(H - h < h && W - w < w, CIDNET_ERR_INVALID, "pre_u8: reflect padding must be smaller than the image");
    if (B == 0) return CIDNET_OK;
    CIDNET_CHECK(src_hwc && dst_nchw, CIDNET_ERR_INVALID, "pre_u8: null pointer");
    dim3 grid(ceil_div(W, 256), H, B);
    pre_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_hwc, dst_nchw, h, w, H, W, gamma, gamma != 1.0f);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

extern "C" int cidnet_post_u8(const float* src_nchw, uint8_t* dst_hwc, int B, int h, int w, int H, int W, void* stream) {
    CIDNET_CHECK(B >= 0 && h > 0 && w > 0 && H >= h && W >= w, CIDNET_ERR_INVALID, "post_u8: bad shape");
    if (B == 0) return CIDNET_OK;
    CIDNET_CHECK(src_nchw && dst_hwc, CIDNET_ERR_INVALID, "post_u8: null pointer");
    dim3 grid(ceil_div(w, 256), h, B);
    post_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src_nchw, dst_hwc, h, w, H, W);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}
