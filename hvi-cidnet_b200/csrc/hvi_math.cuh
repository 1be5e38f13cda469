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

// ---------------------------------------------------------------------------------------------
// Two implementations of the transcendental parts:
//   * CIDNET_HVI_ACCURATE: libdevice sinf/cosf/powf/atan2f and IEEE divisions, every reference op
//     rounded separately -- ~200 instructions per pixel, compute-bound at ~40 % of the HBM roof;
//   * default (fast): the same formulas with cheaper building blocks whose errors are budgeted against
//     the 1e-5 contract (measured max-abs vs the fp32 reference: see tests/test_hvi_gpu.py):
//       sin(v*pi/2), v in [0,1]  : odd minimax polynomial of degree 11, relative error 1.6e-7 (small
//                                  arguments matter: cs = (sin+eps)^k amplifies RELATIVE error only by k)
//       x^k                      : ex2.approx(k * lg2.approx(x)), relative error < 1e-6 for x in [1e-8,1]
//       cos/sin(2*pi*h)          : h reduced to [-0.5,0.5) (exact), then MUFU sin/cos, abs error 4e-7
//       atan2                    : min/max reduction + odd minimax polynomial of degree 15, abs error 1.7e-7
//       divisions                : rcp.approx + multiply (2 ulp)
// ---------------------------------------------------------------------------------------------
#ifdef CIDNET_HVI_ACCURATE
#define CIDNET_HVI_FAST 0
#else
#define CIDNET_HVI_FAST 1
#endif

__device__ __forceinline__ float sin_halfpi_poly(float x) {     // sin(x), x in [0, pi/2]
    const float u = x * x;
    float p = -2.3889498379503493e-08f;
    p = fmaf(p, u, 2.752528644123231e-06f);
    p = fmaf(p, u, -0.00019840861205011606f);
    p = fmaf(p, u, 0.008333330973982811f);
    p = fmaf(p, u, -0.1666666716337204f);
    p = fmaf(p, u, 1.0f);
    return p * x;
}
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float z = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    const float u = z * z;
    float p = -0.004668773151934147f;
    p = fmaf(p, u, 0.02416618913412094f);
    p = fmaf(p, u, -0.0593671016395092f);
    p = fmaf(p, u, 0.09906096756458282f);
    p = fmaf(p, u, -0.14016585052013397f);
    p = fmaf(p, u, 0.19969235360622406f);
    p = fmaf(p, u, -0.33331960439682007f);
    p = fmaf(p, u, 0.9999998807907104f);
    float a = p * z;
    if (ay > ax) a = 1.57079637050628662f - a;
    if (x < 0.f) a = 3.14159274101257324f - a;
    return y < 0.f ? -a : a;
}

// color_sensitive = (sin(v * 0.5 * pi) + eps) ** k        (:40 and :60)
__device__ __forceinline__ float color_sensitive(float v, float k) {
#if CIDNET_HVI_FAST
    const float s = sin_halfpi_poly(v * (0.5f * CIDNET_PI_F)) + CIDNET_EPS_F;
    return exp2f(k * __log2f(s));
#else
    float a = __fmul_rn(__fmul_rn(v, 0.5f), CIDNET_PI_F);
    return powf(__fadd_rn(sinf(a), CIDNET_EPS_F), k);
#endif
}

// RGB -> HVI, one pixel.  Mask priority (min==max) > (r==max) > (g==max) > (b==max)
// reproduces the assignment order of :23-27.
__device__ __forceinline__ void hvit_px(float r, float g, float b, float k,
                                        float& H, float& V, float& I) {
    const float value = fmaxf(fmaxf(r, g), b);
    const float vmin = fminf(fminf(r, g), b);
#if CIDNET_HVI_FAST
    {
        const float rd = __frcp_rn(value - vmin + CIDNET_EPS_F);
        float hue;
        if (vmin == value)   hue = 0.f;
        else if (r == value) { const float t = (g - b) * rd; hue = t < 0.f ? t + 6.f : t; }   // python %, |t| < 1
        else if (g == value) hue = 2.f + (b - r) * rd;
        else                 hue = 4.f + (r - g) * rd;
        float hh = hue * 0.16666667163372040f;                    // hue / 6
        float sat = __fdividef(value - vmin, value + CIDNET_EPS_F);
        if (value == 0.f) sat = 0.f;
        const float css = color_sensitive(value, k) * sat;
        I = value;
        if (hue > 5.999f) {
            // hue just below the 6 -> 0 wrap: the reference's V = cs*s*sin(fp32(2*pi)*fp32(hue/6)) is a tiny
            // NEGATIVE number made of fp32 rounding of the angle, and PHVIT's `h % 1 == 1.0 -> black pixel`
            // hole depends on it.  Reproduce the same roundings and an exact reduction by
            // 2*pi = hi + lo instead of trusting MUFU.SIN's absolute error there (rare branch).
            const float a = __fmul_rn(CIDNET_2PI_F, __fdiv_rn(hue, 6.f));
            const float rr = (a - CIDNET_2PI_F) - (-1.74845553e-07f);     // a - 2*pi, |rr| < 1.1e-3
            H = css * __cosf(rr);
            V = css * rr;                                                  // sin(rr) = rr (1 - 2e-7)
            return;
        }
        if (hh >= 0.5f) hh -= 1.f;                                // exact; same cos/sin, |angle| <= pi
        const float ang = CIDNET_2PI_F * hh;
        H = css * __cosf(ang);
        V = css * __sinf(ang);
        return;
    }
#endif
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
#if CIDNET_HVI_FAST
    {
        const float inv = __frcp_rn(color_sensitive(v, p.k) + CIDNET_EPS_F);
        H = fminf(fmaxf(H * inv, -1.f), 1.f);
        V = fminf(fmaxf(V * inv, -1.f), 1.f);
        float h = atan2_poly(V + CIDNET_EPS_F, H + CIDNET_EPS_F) * 0.15915493667125702f;   // / (2*pi)
        if (h < 0.f) h += 1.f;                                    // python % 1 (may round to exactly 1.0 -> black pixel)
        float s = sqrtf(fmaf(H, H, fmaf(V, V, CIDNET_EPS_F)));
        if (p.gated) s *= p.alpha_s;
        s = fminf(s, 1.f);
        const float h6 = h * 6.f;
        const float hi = floorf(h6);
        const float f = h6 - hi;
        const float pp = v * (1.f - s);
        const float qq = v * (1.f - f * s);
        const float tt = v * (1.f - (1.f - f) * s);
        r = 0.f; g = 0.f; b = 0.f;
        if      (hi == 0.f) { r = v;  g = tt; b = pp; }
        else if (hi == 1.f) { r = qq; g = v;  b = pp; }
        else if (hi == 2.f) { r = pp; g = v;  b = tt; }
        else if (hi == 3.f) { r = pp; g = qq; b = v;  }
        else if (hi == 4.f) { r = tt; g = pp; b = v;  }
        else if (hi == 5.f) { r = v;  g = pp; b = qq; }
        if (p.gated2) { r *= p.alpha; g *= p.alpha; b *= p.alpha; }
        return;
    }
#endif
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
