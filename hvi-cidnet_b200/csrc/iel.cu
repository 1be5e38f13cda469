// IEL gate: the whole depthwise / tanh / product chain of the IEL block in ONE kernel,
// staged in shared memory (replaces 3 depthwise convs, 2 tanh, 2 adds, 1 mul and the
// two chunk() views of net/LCA.py:61-65).
//
// CTA = 16x16 output pixels x 16 hidden channels (of x1 AND the matching 16 of x2).
//   t tile   (20x20, 2-pixel halo)  act_t  -> shared
//   d tile   (18x18, 1-pixel halo)  fp32   -> shared   (d is ZERO outside the image: that is the
//                                                        zero padding dwconv1/dwconv2 see)
//   g        (16x16)                       -> global NHWC
#include "iel.cuh"

namespace cidnet {

static constexpr int kIT = 16;
static constexpr int kT2 = kIT + 4;   // t tile edge
static constexpr int kT1 = kIT + 2;   // d tile edge
static constexpr int kCG = 16;        // channels per CTA (per half)

__global__ void __launch_bounds__(256)
iel_gate_kernel(const IelGateArgs a) {
    extern __shared__ __align__(16) uint8_t iel_smem[];
    act_t* s_t = reinterpret_cast<act_t*>(iel_smem);                         // [2][400][16]
    float* s_d = reinterpret_cast<float*>(s_t + 2 * kT2 * kT2 * kCG);        // [2][324][16]
    float* s_w = s_d + 2 * kT1 * kT1 * kCG;                                  // [2][3 (w0,w1/w2 unused slot)][9][16]
    const int tid = threadIdx.x;
    const int hp = a.hp;
    const int ngroups = hp / kCG;
    const int prob = blockIdx.z % a.nprob;
    const int b = blockIdx.z / a.nprob;
    const int cg = blockIdx.y % ngroups;
    const int tiles_x = (a.W + kIT - 1) / kIT;
    const int y0 = (blockIdx.x / tiles_x) * kIT, x0 = (blockIdx.x % tiles_x) * kIT;
    const int c0 = cg * kCG;
    const int pitch_t = 2 * hp;
    const long long hw = (long long)a.H * a.W;
    const act_t* t = a.t[prob] + (long long)b * hw * pitch_t;

    // weights: s_w[half][0][tap][16] = dwconv, s_w[half][1][tap][16] = dwconv1 / dwconv2
    for (int i = tid; i < 2 * 2 * 9 * kCG; i += 256) {
        const int c = i % kCG, tap = (i / kCG) % 9, which = (i / (kCG * 9)) % 2, half = i / (kCG * 9 * 2);
        float v;
        if (which == 0) v = a.w0[prob][tap * 2 * hp + half * hp + c0 + c];
        else v = (half == 0 ? a.w1[prob] : a.w2[prob])[tap * hp + c0 + c];
        s_w[i] = v;
    }
    // stage t: 400 pixels x 2 halves x 2 vectors of 8 channels
    for (int i = tid; i < kT2 * kT2 * 4; i += 256) {
        const int p = i >> 2, hv = i & 3, half = hv >> 1, v = hv & 1;
        const int y = y0 + p / kT2 - 2, x = x0 + p % kT2 - 2;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (y >= 0 && y < a.H && x >= 0 && x < a.W)
            val = *reinterpret_cast<const uint4*>(t + ((long long)y * a.W + x) * pitch_t + half * hp + c0 + v * 8);
        *reinterpret_cast<uint4*>(s_t + ((size_t)half * kT2 * kT2 + p) * kCG + v * 8) = val;
    }
    __syncthreads();
    // d = dwconv(t) on the 18x18 tile; zero outside the image
    for (int i = tid; i < kT1 * kT1 * 4; i += 256) {
        const int p = i >> 2, hv = i & 3, half = hv >> 1, v = hv & 1;
        const int py = p / kT1, px = p % kT1;
        const int y = y0 + py - 1, x = x0 + px - 1;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        if (y >= 0 && y < a.H && x >= 0 && x < a.W) {
            const float* w = s_w + (half * 2 + 0) * 9 * kCG + v * 8;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                float f[8];
                load8(s_t + ((size_t)half * kT2 * kT2 + (py + tap / 3) * kT2 + px + tap % 3) * kCG + v * 8, f);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(f[e], w[tap * kCG + e], acc[e]);
            }
        }
        float* d = s_d + ((size_t)half * kT1 * kT1 + p) * kCG + v * 8;
        *reinterpret_cast<float4*>(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(d + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    // gate
    act_t* g = a.g[prob] + (long long)b * hw * hp;
    for (int i = tid; i < kIT * kIT * 2; i += 256) {
        const int p = i >> 1, v = i & 1;
        const int py = p / kIT, px = p % kIT;
        const int y = y0 + py, x = x0 + px;
        if (y >= a.H || x >= a.W) continue;
        float out[8];
        float xs[2][8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const float* w = s_w + (half * 2 + 1) * 9 * kCG + v * 8;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float* d = s_d + ((size_t)half * kT1 * kT1 + (py + tap / 3) * kT1 + px + tap % 3) * kCG + v * 8;
                const float4 d0 = *reinterpret_cast<const float4*>(d), d1 = *reinterpret_cast<const float4*>(d + 4);
                const float dv[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(dv[e], w[tap * kCG + e], acc[e]);
            }
            const float* dc = s_d + ((size_t)half * kT1 * kT1 + (py + 1) * kT1 + px + 1) * kCG + v * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) xs[half][e] = tanhf(acc[e]) + dc[e];
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) out[e] = xs[0][e] * xs[1][e];
        store8(g + ((long long)y * a.W + x) * hp + c0 + v * 8, out);
    }
}

int launch_iel_gate(const IelGateArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.hp % kCG == 0, CIDNET_ERR_INVALID, "iel: hp % 16");
    const size_t smem = 2 * kT2 * kT2 * kCG * sizeof(act_t) + 2 * kT1 * kT1 * kCG * sizeof(float) +
                        2 * 2 * 9 * kCG * sizeof(float);
    static bool configured = false;
    if (!configured) {
        CIDNET_CUDA_OK(cudaFuncSetAttribute(iel_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int tiles = ceil_div(a.W, kIT) * ceil_div(a.H, kIT);
    dim3 grid(tiles, a.hp / kCG, a.B * a.nprob);
    iel_gate_kernel<<<grid, 256, smem, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
