// Standalone RGB<->HVI kernels (HBM-bound: 24 B/pixel each, fp32 planar NCHW).
// Replaces RGB_HVI.HVIT / PHVIT of /root/reference/net/HVI_transform.py:16-122,
// which issue ~290 / ~810 ATen calls with host syncs; here each is ONE launch.
#include "common.cuh"
#include "hvi_math.cuh"

namespace cidnet {

// streaming 128-bit accesses: every byte is touched exactly once
__device__ __forceinline__ float4 ldg_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool kInverse>
__global__ void __launch_bounds__(256)
hvi_vec4_kernel(const float* __restrict__ src, float* __restrict__ dst,
                int64_t quads_per_img, int64_t total_quads, int64_t hw, float k, const float* __restrict__ k_dev,
                PhvitParams pp) {
    if (k_dev) { k = __ldg(k_dev); pp.k = k; }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_quads; q += stride) {
        const int64_t img = q / quads_per_img;
        const int64_t off = img * 3 * hw + (q - img * quads_per_img) * 4;
        const float4 a = ldg_stream(src + off);
        const float4 b = ldg_stream(src + off + hw);
        const float4 c = ldg_stream(src + off + 2 * hw);
        float4 o0, o1, o2;
        if (!kInverse) {
            hvit_px(a.x, b.x, c.x, k, o0.x, o1.x, o2.x);
            hvit_px(a.y, b.y, c.y, k, o0.y, o1.y, o2.y);
            hvit_px(a.z, b.z, c.z, k, o0.z, o1.z, o2.z);
            hvit_px(a.w, b.w, c.w, k, o0.w, o1.w, o2.w);
        } else {
            phvit_px(a.x, b.x, c.x, pp, o0.x, o1.x, o2.x);
            phvit_px(a.y, b.y, c.y, pp, o0.y, o1.y, o2.y);
            phvit_px(a.z, b.z, c.z, pp, o0.z, o1.z, o2.z);
            phvit_px(a.w, b.w, c.w, pp, o0.w, o1.w, o2.w);
        }
        stg_stream(dst + off, o0);
        stg_stream(dst + off + hw, o1);
        stg_stream(dst + off + 2 * hw, o2);
    }
}

// generic path for H*W not a multiple of 4 or unaligned bases
template <bool kInverse>
__global__ void __launch_bounds__(256)
hvi_scalar_kernel(const float* __restrict__ src, float* __restrict__ dst,
                  int64_t total_px, int64_t hw, float k, const float* __restrict__ k_dev, PhvitParams pp) {
    if (k_dev) { k = __ldg(k_dev); pp.k = k; }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_px; p += stride) {
        const int64_t img = p / hw;
        const int64_t off = img * 3 * hw + (p - img * hw);
        float o0, o1, o2;
        if (!kInverse) hvit_px(src[off], src[off + hw], src[off + 2 * hw], k, o0, o1, o2);
        else           phvit_px(src[off], src[off + hw], src[off + 2 * hw], pp, o0, o1, o2);
        dst[off] = o0; dst[off + hw] = o1; dst[off + 2 * hw] = o2;
    }
}

template <bool kInverse>
static int launch_hvi(const float* src, float* dst, int B, int H, int W, float k, const float* k_dev,
                      const PhvitParams& pp, cudaStream_t stream) {
    CIDNET_CHECK(B >= 0 && H >= 0 && W >= 0, CIDNET_ERR_INVALID, "negative image dimension");
    const int64_t hw = (int64_t)H * W;
    const int64_t total = (int64_t)B * hw;
    if (total == 0) return CIDNET_OK;
    CIDNET_CHECK(src != nullptr && dst != nullptr, CIDNET_ERR_INVALID, "null image pointer");
    int dev = 0, sms = 148;
    CIDNET_CUDA_OK(cudaGetDevice(&dev));
    CIDNET_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool vec = (hw % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0);
    const int threads = 256;
    if (vec) {
        const int64_t quads = total / 4;
        // 8 resident CTAs/SM; grid-stride keeps it a multiple of the SM count
        int64_t blocks = (quads + threads - 1) / threads;
        const int64_t cap = (int64_t)sms * 8;
        if (blocks > cap) blocks = cap;
        hvi_vec4_kernel<kInverse><<<(unsigned)blocks, threads, 0, stream>>>(src, dst, hw / 4, quads, hw, k, k_dev, pp);
    } else {
        int64_t blocks = (total + threads - 1) / threads;
        const int64_t cap = (int64_t)sms * 8;
        if (blocks > cap) blocks = cap;
        hvi_scalar_kernel<kInverse><<<(unsigned)blocks, threads, 0, stream>>>(src, dst, total, hw, k, k_dev, pp);
    }
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet

extern "C" int cidnet_hvit(const float* rgb, float* hvi, int B, int H, int W, float k, const float* k_dev,
                           void* stream) {
    cidnet::PhvitParams pp{};
    return cidnet::launch_hvi<false>(rgb, hvi, B, H, W, k, k_dev, pp, (cudaStream_t)stream);
}

extern "C" int cidnet_phvit(const float* hvi, float* rgb, int B, int H, int W, float k, const float* k_dev,
                            int gated, float alpha_s, int gated2, float alpha, void* stream) {
    cidnet::PhvitParams pp{k, alpha_s, alpha, gated, gated2};
    return cidnet::launch_hvi<true>(hvi, rgb, B, H, W, k, k_dev, pp, (cudaStream_t)stream);
}
