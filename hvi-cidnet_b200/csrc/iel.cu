// IEL gate: the whole depthwise / tanh / product chain of the IEL block in ONE kernel
// (replaces 3 depthwise convs, 2 tanh, 2 adds, 1 mul and the chunk() views of net/LCA.py:61-65):
//
//   d  = dwconv(t)            (3x3 depthwise on [x1 | x2], zero pad)
//   x1 = tanh(dwconv1(d1)) + d1 ;  x2 = tanh(dwconv2(d2)) + d2 ;  g = x1 * x2
//
// CTA = 32 columns of d (30 output columns) x 32 hidden channels of each half, walking DOWN a
// strip of rows.  Stage 1: thread (half, 8-channel vector, column) slides a 3x3 window of t
// (read straight from global, neighbours are L1 hits) and writes one row of d (fp32) into a
// 4-row ring in shared memory -- d is ZERO outside the image, which is exactly the zero
// padding dwconv1/dwconv2 see.  Stage 2: thread (4-channel sub-vector, vector, column) reads the
// 3x3 neighbourhood of d for BOTH halves from the ring, applies dwconv1/2 + tanh + residual,
// multiplies and stores g.  One __syncthreads per row.
#include "iel.cuh"

namespace cidnet {

static constexpr int kCols = 32;       // d columns per CTA (outputs: kCols - 2)
static constexpr int kRows = 32;       // output rows per CTA
static constexpr int kCh = 32;         // channels per half per CTA
static constexpr int kThreadsIel = 256;

// tanh(x) = 1 - 2 / (exp(2x) + 1): two fast SFU ops, abs error ~1e-7 (vs ~5e-4 of tanh.approx)
__device__ __forceinline__ float fast_tanh(float x) {
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, e + 1.f);
}

__global__ void __launch_bounds__(kThreadsIel, 2)
iel_gate_kernel(const IelGateArgs a) {
    // ring: [4 rows][2 halves][4 vecs][2 planes][32 cols] float4
    __shared__ float4 s_d[4 * 2 * 4 * 2 * kCols];
    __shared__ __align__(16) float s_w0[9 * 2 * kCh];     // dwconv   [tap][half][32]
    __shared__ __align__(16) float s_w12[9 * 2 * kCh];    // dwconv1/2 [tap][half][32]
    const int tid = threadIdx.x;
    const int hp = a.hp, ngroups = hp / kCh;
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int cg = blockIdx.x % ngroups, strip = blockIdx.x / ngroups;
    const int c0 = cg * kCh;
    const int X0 = strip * (kCols - 2) - 1;                // image column of d column 0
    const int y0 = blockIdx.y * kRows;
    const int y1 = min(y0 + kRows, a.H);
    const int pitch_t = 2 * hp;
    const long long hw = (long long)a.H * a.W;

    for (int i = tid; i < 9 * 2 * kCh; i += kThreadsIel) {
        const int c = i % kCh, half = (i / kCh) & 1, tap = i / (2 * kCh);
        s_w0[i] = a.w0[prob][tap * 2 * hp + half * hp + c0 + c];
        s_w12[i] = (half == 0 ? a.w1[prob] : a.w2[prob])[tap * hp + c0 + c];
    }

    const int dx = tid & 31;                  // column (lane): a warp shares (half|sub, vec) -> smem broadcasts
    const int vec = (tid >> 5) & 3;           // 8-channel vector within the group
    const int hs = tid >> 7;                  // stage 1: half;  stage 2: 4-channel sub-vector (plane)
    const int xd = X0 + dx;                   // image column of this thread's d column
    const bool col_in = xd >= 0 && xd < a.W;
    const bool has_l = xd - 1 >= 0 && xd - 1 < a.W, has_r = xd + 1 >= 0 && xd + 1 < a.W;
    const act_t* tsrc = a.t[prob] + (long long)b * hw * pitch_t + hs * hp + c0 + vec * 8;
    act_t* gdst = a.g[prob] + (long long)b * hw * hp + c0 + vec * 8 + hs * 4;
    const bool writer = dx >= 1 && dx <= kCols - 2 && col_in;

    uint4 win0[3], win1[3], win2[3];          // raw 16-bit t values of three consecutive rows
    auto load_row = [&](int y, uint4* r) {
        r[0] = r[1] = r[2] = make_uint4(0, 0, 0, 0);
        if (y < 0 || y >= a.H) return;
        const act_t* p = tsrc + ((long long)y * a.W + xd) * pitch_t;
        if (col_in) r[1] = *reinterpret_cast<const uint4*>(p);
        if (has_l) r[0] = *reinterpret_cast<const uint4*>(p - pitch_t);
        if (has_r) r[2] = *reinterpret_cast<const uint4*>(p + pitch_t);
    };
    auto fma_row = [&](const uint4* r, int tap0, float* acc) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const act_t* h = reinterpret_cast<const act_t*>(&r[c]);
            const float4 wa = *reinterpret_cast<const float4*>(s_w0 + ((tap0 + c) * 2 + hs) * kCh + vec * 8);
            const float4 wb = *reinterpret_cast<const float4*>(s_w0 + ((tap0 + c) * 2 + hs) * kCh + vec * 8 + 4);
            acc[0] = fmaf(act2f(h[0]), wa.x, acc[0]); acc[1] = fmaf(act2f(h[1]), wa.y, acc[1]);
            acc[2] = fmaf(act2f(h[2]), wa.z, acc[2]); acc[3] = fmaf(act2f(h[3]), wa.w, acc[3]);
            acc[4] = fmaf(act2f(h[4]), wb.x, acc[4]); acc[5] = fmaf(act2f(h[5]), wb.y, acc[5]);
            acc[6] = fmaf(act2f(h[6]), wb.z, acc[6]); acc[7] = fmaf(act2f(h[7]), wb.w, acc[7]);
        }
    };
    // ring address of (row slot, half, vec, plane, col)
    auto ring = [&](int slot, int half, int v, int plane, int col) -> float4* {
        return s_d + ((((slot * 2 + half) * 4 + v) * 2 + plane) * kCols + col);
    };

    __syncthreads();
    // one iteration: d(r) -> ring, then output row r-1.  (w0, w1, w2) = t rows r-1, r, r+1.
    auto iter = [&](int r, uint4* w0, uint4* w1, uint4* w2) {
        // ---- stage 1: d(r) for (half = hs, vec, column dx)
        load_row(r + 1, w2);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        if (r >= 0 && r < a.H && col_in) {
            fma_row(w0, 0, acc);
            fma_row(w1, 3, acc);
            fma_row(w2, 6, acc);
        }
        *ring(r & 3, hs, vec, 0, dx) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *ring(r & 3, hs, vec, 1, dx) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        __syncthreads();
        // ---- stage 2: output row yo = r - 1, channels [4*hs, 4*hs+4) of vec, both halves
        const int yo = r - 1;
        if (yo >= y0 && writer) {
            float xs[2][4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const int slot = (yo - 1 + rr) & 3;
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        const float4 dv = *ring(slot, half, vec, hs, dx - 1 + cc);
                        const float4 wv = *reinterpret_cast<const float4*>(s_w12 + ((rr * 3 + cc) * 2 + half) * kCh + vec * 8 + hs * 4);
                        o[0] = fmaf(dv.x, wv.x, o[0]); o[1] = fmaf(dv.y, wv.y, o[1]);
                        o[2] = fmaf(dv.z, wv.z, o[2]); o[3] = fmaf(dv.w, wv.w, o[3]);
                    }
                }
                const float4 dc = *ring(yo & 3, half, vec, hs, dx);
                xs[half][0] = fast_tanh(o[0]) + dc.x; xs[half][1] = fast_tanh(o[1]) + dc.y;
                xs[half][2] = fast_tanh(o[2]) + dc.z; xs[half][3] = fast_tanh(o[3]) + dc.w;
            }
            uint2 raw;
            act_t* ov = reinterpret_cast<act_t*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) ov[e] = f2act(xs[0][e] * xs[1][e]);
            *reinterpret_cast<uint2*>(gdst + ((long long)yo * a.W + xd) * hp) = raw;
        }
        // the next iteration writes ring slot (r+1)&3, which the stage 2 above (rows r-2..r) never reads
    };
    // d rows r = y0-1 .. y1; the three window registers rotate roles (no register moves)
    load_row(y0 - 2, win0);
    load_row(y0 - 1, win1);
    for (int r = y0 - 1; r <= y1; r += 3) {
        iter(r, win0, win1, win2);
        if (r + 1 <= y1) iter(r + 1, win1, win2, win0);
        if (r + 2 <= y1) iter(r + 2, win2, win0, win1);
    }
}

int launch_iel_gate(const IelGateArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.hp % kCh == 0, CIDNET_ERR_INVALID, "iel: hp % 32");
    const int strips = ceil_div(a.W, kCols - 2);
    dim3 grid(strips * (a.hp / kCh), ceil_div(a.H, kRows), a.B * a.nprob);
    iel_gate_kernel<<<grid, kThreadsIel, 0, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
