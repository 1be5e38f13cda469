// Per-pixel RGB<->HVI arithmetic shared by the standalone transform kernels and
// by the fused stem / head kernels of the forward path.
//
// Follows /root/reference/net/HVI_transform.py:16-47 (HVIT) and :49-122 (PHVIT)
// operation by operation in fp32.  Every product/sum that the reference performs
// as a separate tensor op is written with __fmul_rn/__fadd_rn/__fsub_rn so nvcc
// cannot contract it into an FMA (the reference rounds after every op).
#pragma once
#include <cuda_runtime.h>

namespace cidnet {

__device__ __forceinline__ float pymodf(float a, float b) {
    // torch.remainder for floats: fmod, then shift into the sign of b.
    float m = fmodf(a, b);
    if (m != 0.f && ((b < 0.f) != (m < 0.f))) m = __fadd_rn(m, b);
    return m;
}

#define CIDNET_PI_F     3.14159274101257324f   /* fp32(3.141592653589793)      */
#define CIDNET_2PI_F    6.28318548202514648f   /* fp32(2.0 * 3.141592653589793) */
#define CIDNET_EPS_F    1e-8f

// color_sensitive = (sin(v * 0.5 * pi) + eps) ** k        (:40 and :60)
__device__ __forceinline__ float color_sensitive(float v, float k) {
    float a = __fmul_rn(__fmul_rn(v, 0.5f), CIDNET_PI_F);
    return powf(__fadd_rn(sinf(a), CIDNET_EPS_F), k);
}

// RGB -> HVI, one pixel.  Mask priority (min==max) > (r==max) > (g==max) > (b==max)
// reproduces the assignment order of :23-27.
__device__ __forceinline__ void hvit_px(float r, float g, float b, float k,
                                        float& H, float& V, float& I) {
    const float value = fmaxf(fmaxf(r, g), b);
    const float vmin = fminf(fminf(r, g), b);
    const float d = __fadd_rn(__fsub_rn(value, vmin), CIDNET_EPS_F);
    float hue;
    if (vmin == value)      hue = 0.f;
    else if (r == value)    hue = pymodf(__fdiv_rn(__fsub_rn(g, b), d), 6.f);
    else if (g == value)    hue = __fadd_rn(2.f, __fdiv_rn(__fsub_rn(b, r), d));
    else                    hue = __fadd_rn(4.f, __fdiv_rn(__fsub_rn(r, g), d));
    hue = __fdiv_rn(hue, 6.f);
    float sat = __fdiv_rn(__fsub_rn(value, vmin), __fadd_rn(value, CIDNET_EPS_F));
    if (value == 0.f) sat = 0.f;
    const float cs = color_sensitive(value, k);
    const float ang = __fmul_rn(CIDNET_2PI_F, hue);
    float sn, cn;
    sincosf(ang, &sn, &cn);
    const float css = __fmul_rn(cs, sat);
    H = __fmul_rn(css, cn);
    V = __fmul_rn(css, sn);
    I = value;
}

struct PhvitParams {
    float k;        // this_k
    float alpha_s;  // applied when gated
    float alpha;    // applied when gated2
    int gated;
    int gated2;
};

// HVI -> RGB, one pixel (:49-122).
__device__ __forceinline__ void phvit_px(float H, float V, float I, const PhvitParams& p,
                                         float& r, float& g, float& b) {
    H = fminf(fmaxf(H, -1.f), 1.f);
    V = fminf(fmaxf(V, -1.f), 1.f);
    I = fminf(fmaxf(I, 0.f), 1.f);
    float v = I;
    const float cs = color_sensitive(v, p.k);
    const float den = __fadd_rn(cs, CIDNET_EPS_F);
    H = fminf(fmaxf(__fdiv_rn(H, den), -1.f), 1.f);
    V = fminf(fmaxf(__fdiv_rn(V, den), -1.f), 1.f);
    float h = __fdiv_rn(atan2f(__fadd_rn(V, CIDNET_EPS_F), __fadd_rn(H, CIDNET_EPS_F)), CIDNET_2PI_F);
    h = pymodf(h, 1.f);
    float s = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(H, H), __fmul_rn(V, V)), CIDNET_EPS_F));
    if (p.gated) s = __fmul_rn(s, p.alpha_s);
    s = fminf(fmaxf(s, 0.f), 1.f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    const float h6 = __fmul_rn(h, 6.f);
    const float hi = floorf(h6);
    const float f = __fsub_rn(h6, hi);
    const float pp = __fmul_rn(v, __fsub_rn(1.f, s));
    const float qq = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(f, s)));
    const float tt = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(__fsub_rn(1.f, f), s)));
    // sextant table :92-114; hi outside 0..5 (h%1 rounding to 1.0) leaves the pixel black.
    r = 0.f; g = 0.f; b = 0.f;
    if      (hi == 0.f) { r = v;  g = tt; b = pp; }
    else if (hi == 1.f) { r = qq; g = v;  b = pp; }
    else if (hi == 2.f) { r = pp; g = v;  b = tt; }
    else if (hi == 3.f) { r = pp; g = qq; b = v;  }
    else if (hi == 4.f) { r = tt; g = pp; b = v;  }
    else if (hi == 5.f) { r = v;  g = pp; b = qq; }
    if (p.gated2) { r = __fmul_rn(r, p.alpha); g = __fmul_rn(g, p.alpha); b = __fmul_rn(b, p.alpha); }
}

}  // namespace cidnet
