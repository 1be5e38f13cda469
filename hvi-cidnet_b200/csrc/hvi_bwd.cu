// Backward (vector-Jacobian products) of the RGB<->HVI transform: what autograd derives from the tensor programs of
// /root/reference/net/HVI_transform.py:16-47 (HVIT) and :49-122 (PHVIT).  The reference's training loop differentiates
// both: `model.HVIT(output_rgb)` inside the loss (train.py:61-62) and `self.trans.PHVIT(output_hvi)` at the end of the
// forward (net/CIDNet.py:121).  There the backward is ~600 ATen calls over ~40 saved full-size temporaries; here each
// is ONE launch that recomputes the forward's per-pixel quantities from the input (HBM-bound: 36 B/pixel -- input 12,
// upstream gradient 12, result 12; nothing is saved by the forward).
//
// Autograd conventions reproduced (oracle/hvi_backward.py states them next to the reference lines):
//   max(1)[0] / min(1)[0] -> the first extremal channel gets the gradient; masked hue assignments -> the last one wins
//   (r==max > g==max > b==max), grey pixels and value==0 cut the hue / saturation paths; `%` passes the gradient;
//   clamp passes it on the closed interval; floor has none; pow(k) w.r.t. the parameter k is result*log(base), summed
//   over every pixel -- here: per-thread partial -> fixed-order block reduction -> one slot per CTA -> a one-CTA
//   finishing kernel that adds the slots in index order (no atomics: bit-reproducible).
// Transcendental parts.  First version: libdevice sincosf x2 / powf / logf / atan2f and ~10 IEEE divisions per pixel --
// ~350 instructions, 2.3-2.5 TB/s = 36 % of the measured HBM peak on B200 (compute-bound), parity at 0.5-1.4 % of the
// tolerance.  Default now (CIDNET_HVI_FAST, as the forward kernels): rcp + multiply for the divisions, x^k as
// ex2(k * lg2 x) with lg2 shared with the d/dk term, MUFU sin / cos of the range-reduced hue angle, the forward kernels'
// atan2 polynomial.  ONE accurate call stays: sinf(I * pi/2) -- at I = 1 it must be exactly 1.0f so that the base of the
// power, the power and `den` are exactly 1 and a clamped |H| = 1 stays ON the closed clamp interval of :63-64 (where
// autograd passes the gradient) instead of landing one ulp outside it.
#include "common.cuh"
#include "hvi_math.cuh"

namespace cidnet {

static constexpr int kBwdThreads = 256;
static constexpr int kBwdMaxBlocks = 2048;          // slots of the d/dk scratch buffer

__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg4_stream(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// x / y, 1 / y, x^k (and log x), sin / cos of 2*pi*t for t in [0, 1): accurate or fast building blocks
__device__ __forceinline__ float rcp_(float y) {
#if CIDNET_HVI_FAST
    return __frcp_rn(y);
#else
    return 1.f / y;
#endif
}
__device__ __forceinline__ float div_(float x, float y) {
#if CIDNET_HVI_FAST
    return x * __frcp_rn(y);
#else
    return x / y;
#endif
}
__device__ __forceinline__ void pow_log(float base, float k, float& p, float& ln) {
#if CIDNET_HVI_FAST
    const float lg = __log2f(base);
    p = exp2f(k * lg);
    ln = lg * 0.693147182464599609f;
#else
    p = powf(base, k);
    ln = logf(base);
#endif
}
__device__ __forceinline__ void sincos_turns(float t, float& sn, float& cn) {      // sin / cos(2*pi*t), t in [0, 1)
#if CIDNET_HVI_FAST
    if (t >= 0.5f) t -= 1.f;                                                          // exact; |angle| <= pi
    sn = __sinf(CIDNET_2PI_F * t);
    cn = __cosf(CIDNET_2PI_F * t);
#else
    sincosf(CIDNET_2PI_F * t, &sn, &cn);
#endif
}
__device__ __forceinline__ void sincos_quarter(float ang, float& sn, float& cn) {   // ang in [0, pi/2]
    sn = sinf(ang);                      // accurate on purpose (header comment): exactly 1.0f at pi/2
#if CIDNET_HVI_FAST
    cn = __cosf(ang);
#else
    cn = cosf(ang);
#endif
}

// d(H,V,I)/d(r,g,b)^T applied to (gH,gV,gI); returns this pixel's contribution to d/dk
__device__ __forceinline__ float hvit_bwd_px(float r, float g, float b, float k, float gH, float gV, float gI,
                                             float& gr, float& gg, float& gb) {
    const float value = fmaxf(fmaxf(r, g), b);
    const float vmin = fminf(fminf(r, g), b);
    const float d = (value - vmin) + CIDNET_EPS_F;
    const bool grey = vmin == value;
    const int branch = grey ? -1 : (r == value ? 0 : (g == value ? 1 : 2));    // which assignment wrote the hue last (:22-26)
    float num = 0.f, hue = 0.f;
    const float rd = rcp_(d);
    if (branch == 0)      { num = g - b; hue = num * rd; if (hue < 0.f) hue += 6.f; }      // python % 6, |num / d| < 1
    else if (branch == 1) { num = b - r; hue = 2.f + num * rd; }
    else if (branch == 2) { num = r - g; hue = 4.f + num * rd; }
    hue = hue * 0.16666667163372040f;
    const bool sat_live = value != 0.f;
    const float rve = rcp_(value + CIDNET_EPS_F);
    const float sat = sat_live ? (value - vmin) * rve : 0.f;
    const float ang = value * 0.5f * CIDNET_PI_F;
    float sn, cn;
    sincos_quarter(ang, sn, cn);
    const float sbase = sn + CIDNET_EPS_F;
    float cs, ln_sbase;
    pow_log(sbase, k, cs, ln_sbase);
    float cv, ch;
    sincos_turns(hue, cv, ch);

    const float g_css = gH * ch + gV * cv;                       // H = (cs*sat)*ch, V = (cs*sat)*cv
    const float g_cs = g_css * sat;
    const float g_sat = sat_live ? g_css * cs : 0.f;
    const float g_hue = grey ? 0.f : (gV * ch - gH * cv) * (cs * sat) * (CIDNET_2PI_F / 6.f);
    const float g_num = g_hue * rd;
    const float g_d = -g_num * num * rd;
    const float g_sv = g_sat * rve;
    float g_value = gI + g_d + g_sv - g_sv * sat
                  + g_cs * k * div_(cs, sbase) * cn * (0.5f * CIDNET_PI_F);
    float g_min = -g_d - g_sv;
    gr = branch == 1 ? -g_num : (branch == 2 ? g_num : 0.f);
    gg = branch == 0 ? g_num : (branch == 2 ? -g_num : 0.f);
    gb = branch == 0 ? -g_num : (branch == 1 ? g_num : 0.f);
    if (r == value) gr += g_value; else if (g == value) gg += g_value; else gb += g_value;    // first arg-max
    if (r == vmin)  gr += g_min;   else if (g == vmin)  gg += g_min;   else gb += g_min;      // first arg-min
    return g_cs * cs * ln_sbase;
}

// d(r,g,b)/d(H,V,I)^T applied to (gr,gg,gb)
__device__ __forceinline__ void phvit_bwd_px(float H0, float V0, float I0, const PhvitParams& p, float gr, float gg,
                                             float gb, float& gH, float& gV, float& gI) {
    if (p.gated2) { gr *= p.alpha; gg *= p.alpha; gb *= p.alpha; }
    const float Hc = fminf(fmaxf(H0, -1.f), 1.f), Vc = fminf(fmaxf(V0, -1.f), 1.f), v = fminf(fmaxf(I0, 0.f), 1.f);
    const float ang = v * 0.5f * CIDNET_PI_F;
    float sn, cn;
    sincos_quarter(ang, sn, cn);
    const float sbase = sn + CIDNET_EPS_F;
    float cs, ln_unused;
    pow_log(sbase, p.k, cs, ln_unused);
    const float rden = rcp_(cs + CIDNET_EPS_F);
    const float H2 = Hc * rden, V2 = Vc * rden;
    const float H3 = fminf(fmaxf(H2, -1.f), 1.f), V3 = fminf(fmaxf(V2, -1.f), 1.f);
    const float x = H3 + CIDNET_EPS_F, y = V3 + CIDNET_EPS_F;
#if CIDNET_HVI_FAST
    float h = atan2_poly(y, x) * 0.15915493667125702f;
    if (h < 0.f) h += 1.f;
#else
    float h = atan2f(y, x) / CIDNET_2PI_F;
    h = pymodf(h, 1.f);
#endif
    const float s_raw = sqrtf(H3 * H3 + V3 * V3 + CIDNET_EPS_F);
    const float s_pre = p.gated ? s_raw * p.alpha_s : s_raw;
    const float s = fminf(fmaxf(s_pre, 0.f), 1.f);
    const float h6 = h * 6.f;
    const float hi = floorf(h6);
    const float f = h6 - hi;
    float g_v = 0.f, g_p = 0.f, g_q = 0.f, g_t = 0.f;            // sextant table :92-114, transposed
    if      (hi == 0.f) { g_v = gr; g_t = gg; g_p = gb; }
    else if (hi == 1.f) { g_q = gr; g_v = gg; g_p = gb; }
    else if (hi == 2.f) { g_p = gr; g_v = gg; g_t = gb; }
    else if (hi == 3.f) { g_p = gr; g_q = gg; g_v = gb; }
    else if (hi == 4.f) { g_t = gr; g_p = gg; g_v = gb; }
    else if (hi == 5.f) { g_v = gr; g_p = gg; g_q = gb; }
    g_v += g_p * (1.f - s) + g_q * (1.f - f * s) + g_t * (1.f - (1.f - f) * s);
    float g_s = -g_p * v - g_q * v * f - g_t * v * (1.f - f);
    const float g_f = (g_t - g_q) * v * s;
    if (!(s_pre >= 0.f && s_pre <= 1.f)) g_s = 0.f;
    if (p.gated) g_s *= p.alpha_s;
    const float g_sr = g_s * rcp_(s_raw);
    float g_H3 = g_sr * H3, g_V3 = g_sr * V3;
    const float g_hr = g_f * (6.f / CIDNET_2PI_F) * rcp_(x * x + y * y);
    g_V3 += g_hr * x;
    g_H3 -= g_hr * y;
    const float g_H2 = (H2 >= -1.f && H2 <= 1.f) ? g_H3 : 0.f;
    const float g_V2 = (V2 >= -1.f && V2 <= 1.f) ? g_V3 : 0.f;
    const float g_den = -(g_H2 * H2 + g_V2 * V2) * rden;           // -(g_H2 Hc + g_V2 Vc) / den^2
    if (p.k != 0.f) g_v += g_den * p.k * div_(cs, sbase) * cn * (0.5f * CIDNET_PI_F);
    gH = (H0 >= -1.f && H0 <= 1.f) ? g_H2 * rden : 0.f;
    gV = (V0 >= -1.f && V0 <= 1.f) ? g_V2 * rden : 0.f;
    gI = (I0 >= 0.f && I0 <= 1.f) ? g_v : 0.f;
}

// fixed-order block sum (warp tree by shuffle, then warp 0 over the warps' partials)
__device__ __forceinline__ float block_sum(float v) {
    __shared__ float s_part[kBwdThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < kBwdThreads / 32 ? s_part[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    return t;          // valid in thread 0
}

template <bool kInverse, bool kVec>
__global__ void __launch_bounds__(kBwdThreads)
hvi_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gout, float* __restrict__ gin,
               float* __restrict__ gk_slots, int64_t units_per_img, int64_t total_units, int64_t hw, float k,
               const float* __restrict__ k_dev, PhvitParams pp) {
    if (k_dev) { k = __ldg(k_dev); pp.k = k; }
    float gk = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total_units; q += stride) {
        const int64_t img = q / units_per_img;
        const int64_t off = img * 3 * hw + (q - img * units_per_img) * (kVec ? 4 : 1);
        if (kVec) {
            const float4 a = ldg4_stream(x + off), b = ldg4_stream(x + off + hw), c = ldg4_stream(x + off + 2 * hw);
            const float4 ga = ldg4_stream(gout + off), gb = ldg4_stream(gout + off + hw), gc = ldg4_stream(gout + off + 2 * hw);
            float4 o0, o1, o2;
            if (!kInverse) {
                gk += hvit_bwd_px(a.x, b.x, c.x, k, ga.x, gb.x, gc.x, o0.x, o1.x, o2.x);
                gk += hvit_bwd_px(a.y, b.y, c.y, k, ga.y, gb.y, gc.y, o0.y, o1.y, o2.y);
                gk += hvit_bwd_px(a.z, b.z, c.z, k, ga.z, gb.z, gc.z, o0.z, o1.z, o2.z);
                gk += hvit_bwd_px(a.w, b.w, c.w, k, ga.w, gb.w, gc.w, o0.w, o1.w, o2.w);
            } else {
                phvit_bwd_px(a.x, b.x, c.x, pp, ga.x, gb.x, gc.x, o0.x, o1.x, o2.x);
                phvit_bwd_px(a.y, b.y, c.y, pp, ga.y, gb.y, gc.y, o0.y, o1.y, o2.y);
                phvit_bwd_px(a.z, b.z, c.z, pp, ga.z, gb.z, gc.z, o0.z, o1.z, o2.z);
                phvit_bwd_px(a.w, b.w, c.w, pp, ga.w, gb.w, gc.w, o0.w, o1.w, o2.w);
            }
            stg4_stream(gin + off, o0);
            stg4_stream(gin + off + hw, o1);
            stg4_stream(gin + off + 2 * hw, o2);
        } else {
            float o0, o1, o2;
            if (!kInverse) gk += hvit_bwd_px(x[off], x[off + hw], x[off + 2 * hw], k, gout[off], gout[off + hw], gout[off + 2 * hw], o0, o1, o2);
            else           phvit_bwd_px(x[off], x[off + hw], x[off + 2 * hw], pp, gout[off], gout[off + hw], gout[off + 2 * hw], o0, o1, o2);
            gin[off] = o0; gin[off + hw] = o1; gin[off + 2 * hw] = o2;
        }
    }
    if (!kInverse && gk_slots) {
        const float t = block_sum(gk);
        if (threadIdx.x == 0) gk_slots[blockIdx.x] = t;
    }
}

// d/dk: the CTAs' slots added in index order (one CTA: thread t takes slots t, t+256, ...; then the block tree)
__global__ void __launch_bounds__(kBwdThreads)
hvi_bwd_finish_kernel(const float* __restrict__ gk_slots, int n, float* __restrict__ grad_k) {
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += kBwdThreads) v += __ldcg(gk_slots + i);
    const float t = block_sum(v);
    if (threadIdx.x == 0) grad_k[0] = t;
}

template <bool kInverse>
static int launch_hvi_bwd(const float* x, const float* gout, float* gin, float* grad_k, float* scratch, int B, int H,
                          int W, float k, const float* k_dev, const PhvitParams& pp, cudaStream_t stream) {
    CIDNET_CHECK(B >= 0 && H >= 0 && W >= 0, CIDNET_ERR_INVALID, "negative image dimension");
    const int64_t hw = (int64_t)H * W, total = (int64_t)B * hw;
    CIDNET_CHECK(!(grad_k && !scratch), CIDNET_ERR_INVALID, "hvit_backward: grad_k needs the scratch buffer (cidnet_hvi_backward_scratch_bytes)");
    if (total == 0) {
        if (grad_k) CIDNET_CUDA_OK(cudaMemsetAsync(grad_k, 0, sizeof(float), stream));
        return CIDNET_OK;
    }
    CIDNET_CHECK(x && gout && gin, CIDNET_ERR_INVALID, "null image pointer");
    int dev = 0, sms = 148;
    CIDNET_CUDA_OK(cudaGetDevice(&dev));
    CIDNET_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const bool vec = (hw % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(gin)) % 16 == 0);
    const int64_t units = vec ? total / 4 : total;
    int64_t blocks = (units + kBwdThreads - 1) / kBwdThreads;
    int64_t cap = (int64_t)sms * 8;                     // grid-stride: a multiple of the SM count
    if (cap > kBwdMaxBlocks) cap = kBwdMaxBlocks;
    if (blocks > cap) blocks = cap;
    float* slots = (!kInverse && grad_k) ? scratch : nullptr;
    if (vec) hvi_bwd_kernel<kInverse, true><<<(unsigned)blocks, kBwdThreads, 0, stream>>>(x, gout, gin, slots, hw / 4, units, hw, k, k_dev, pp);
    else     hvi_bwd_kernel<kInverse, false><<<(unsigned)blocks, kBwdThreads, 0, stream>>>(x, gout, gin, slots, hw, units, hw, k, k_dev, pp);
    CIDNET_CUDA_OK(cudaGetLastError());
    if (slots) {
        hvi_bwd_finish_kernel<<<1, kBwdThreads, 0, stream>>>(slots, (int)blocks, grad_k);
        CIDNET_CUDA_OK(cudaGetLastError());
    }
    return CIDNET_OK;
}

}  // namespace cidnet

extern "C" int64_t cidnet_hvi_backward_scratch_bytes(void) { return (int64_t)cidnet::kBwdMaxBlocks * sizeof(float); }

extern "C" int cidnet_hvit_backward(const float* rgb, const float* grad_hvi, float* grad_rgb, float* grad_k, void* scratch,
                                    int B, int H, int W, float k, const float* k_dev, void* stream) {
    cidnet::PhvitParams pp{};
    return cidnet::launch_hvi_bwd<false>(rgb, grad_hvi, grad_rgb, grad_k, static_cast<float*>(scratch), B, H, W, k, k_dev,
                                         pp, (cudaStream_t)stream);
}

extern "C" int cidnet_phvit_backward(const float* hvi, const float* grad_rgb, float* grad_hvi, int B, int H, int W, float k,
                                     const float* k_dev, int gated, float alpha_s, int gated2, float alpha, void* stream) {
    cidnet::PhvitParams pp{k, alpha_s, alpha, gated, gated2};
    return cidnet::launch_hvi_bwd<true>(hvi, grad_rgb, grad_hvi, nullptr, nullptr, B, H, W, k, k_dev, pp, (cudaStream_t)stream);
}
