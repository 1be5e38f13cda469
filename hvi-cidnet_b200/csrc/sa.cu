// SpatialAttention gates of the fork's MSSA variant (/root/reference/net/CIDNet_MSSA.py:10-25, used at :132-153):
//   avg = mean_C(x); mx = max_C(x); y = conv7x7(cat[avg, mx]) (zero padding 3, no bias); x * sigmoid(y)
// Both kernels are HBM-bound elementwise passes over an NHWC 16-bit tensor:
//   sa_stats_kernel  reads x once (2*pitch B/px), writes 8 B/px
//   sa_gate_kernel   reads the 7x7 neighbourhood of the statistics from a shared-memory tile, reads x, writes x
#include "sa.cuh"
#include "ptx_sm100.cuh"

namespace cidnet {

namespace {

constexpr int kStatPx = 128;        // pixels per CTA of the statistics kernel
constexpr int kMaxVec = 18;         // 144 channels / 8

struct SaStatsParams { const act_t* x[2]; float2* stats[2]; long long npx; int C, nv; };

__global__ void __launch_bounds__(256)
sa_stats_kernel(SaStatsParams p) {
    __shared__ float2 part[kStatPx * kMaxVec];
    const act_t* __restrict__ x = blockIdx.y ? p.x[1] : p.x[0];
    const long long px0 = (long long)blockIdx.x * kStatPx;
    const int npx = (int)min((long long)kStatPx, p.npx - px0);
    const int nvec = npx * p.nv;
    const uint4* src = reinterpret_cast<const uint4*>(x) + px0 * p.nv;     // pitch = 8 * nv: pixels are contiguous
    ptx::pdl_wait();
    ptx::pdl_trigger();
    for (int i = threadIdx.x; i < nvec; i += 256) {
        const uint4 raw = __ldcg(src + i);
        const act_t* a = reinterpret_cast<const act_t*>(&raw);
        const int c0 = (i % p.nv) * 8;
        float s = 0.f, m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (c0 + j < p.C) { const float v = act2f(a[j]); s += v; m = fmaxf(m, v); }
        }
        part[i] = make_float2(s, m);
    }
    __syncthreads();
    if (threadIdx.x < npx) {
        float s = 0.f, m = -INFINITY;
        for (int v = 0; v < p.nv; ++v) {
            const float2 q = part[threadIdx.x * p.nv + v];
            s += q.x; m = fmaxf(m, q.y);
        }
        (blockIdx.y ? p.stats[1] : p.stats[0])[px0 + threadIdx.x] = make_float2(s / (float)p.C, m);
    }
}

constexpr int kTW = 32, kTH = 8;    // pixel tile of the gate kernel

struct SaGateParams { act_t* x[2]; const float2* stats[2]; const float* w[2]; int B, H, W, nv; };

__global__ void __launch_bounds__(256)
sa_gate_kernel(SaGateParams p) {
    __shared__ float2 tile[(kTH + 6) * (kTW + 6)];
    __shared__ float wsm[98];
    __shared__ float gate[kTH * kTW];
    const int prob = blockIdx.z / p.B, b = blockIdx.z % p.B;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    const float2* __restrict__ st = (prob ? p.stats[1] : p.stats[0]) + (long long)b * p.H * p.W;
    if (threadIdx.x < 98) wsm[threadIdx.x] = (prob ? p.w[1] : p.w[0])[threadIdx.x];
    ptx::pdl_wait();
    ptx::pdl_trigger();
    for (int i = threadIdx.x; i < (kTH + 6) * (kTW + 6); i += 256) {
        const int ty = i / (kTW + 6), tx = i % (kTW + 6);
        const int y = y0 + ty - 3, x = x0 + tx - 3;
        tile[i] = (y >= 0 && y < p.H && x >= 0 && x < p.W) ? __ldcg(st + (long long)y * p.W + x) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    {
        const int ty = threadIdx.x / kTW, tx = threadIdx.x % kTW;
        float acc = 0.f;
#pragma unroll
        for (int dy = 0; dy < 7; ++dy)
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) {
                const float2 q = tile[(ty + dy) * (kTW + 6) + tx + dx];
                acc = fmaf(wsm[dy * 7 + dx], q.x, acc);
                acc = fmaf(wsm[49 + dy * 7 + dx], q.y, acc);
            }
        gate[threadIdx.x] = 1.0f / (1.0f + expf(-acc));
    }
    __syncthreads();
    act_t* __restrict__ xb = (prob ? p.x[1] : p.x[0]) + (long long)b * p.H * p.W * p.nv * 8;
    const int row_vecs = kTW * p.nv;
    for (int i = threadIdx.x; i < kTH * row_vecs; i += 256) {
        const int ty = i / row_vecs, r = i % row_vecs, tx = r / p.nv;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= p.H || x >= p.W) continue;
        uint4* ptr = reinterpret_cast<uint4*>(xb) + ((long long)y * p.W + x0) * p.nv + r;
        uint4 raw = __ldcg(ptr);
        act_t* a = reinterpret_cast<act_t*>(&raw);
        const float g = gate[ty * kTW + tx];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = f2act(act2f(a[j]) * g);
        *ptr = raw;
    }
}

int check(const SaArgs& a) {
    CIDNET_CHECK(a.nprob >= 1 && a.nprob <= 2 && a.B > 0 && a.H > 0 && a.W > 0, CIDNET_ERR_INVALID, "spatial attention: bad shape");
    CIDNET_CHECK(a.pitch % 8 == 0 && a.pitch / 8 <= kMaxVec && a.C > 0 && a.C <= a.pitch, CIDNET_ERR_INVALID,
                 "spatial attention: channel pitch must be a multiple of 8 and <= 144");
    return CIDNET_OK;
}

}  // namespace

int launch_sa_stats(const SaArgs& a, cudaStream_t stream) {
    int rc = check(a);
    if (rc) return rc;
    SaStatsParams p;
    for (int i = 0; i < 2; ++i) { p.x[i] = a.x[i]; p.stats[i] = a.stats[i]; }
    p.npx = (long long)a.B * a.H * a.W; p.C = a.C; p.nv = a.pitch / 8;
    dim3 grid((unsigned)((p.npx + kStatPx - 1) / kStatPx), a.nprob);
    return launch_k(sa_stats_kernel, grid, dim3(256), 0, stream, p);
}

int launch_sa_gate(const SaArgs& a, cudaStream_t stream) {
    int rc = check(a);
    if (rc) return rc;
    SaGateParams p;
    for (int i = 0; i < 2; ++i) { p.x[i] = a.x[i]; p.stats[i] = a.stats[i]; p.w[i] = a.w[i]; }
    p.B = a.B; p.H = a.H; p.W = a.W; p.nv = a.pitch / 8;
    CIDNET_CHECK((long long)a.B * a.nprob <= 65535, CIDNET_ERR_INVALID, "spatial attention: batch too large for one launch");
    dim3 grid(ceil_div(a.W, kTW), ceil_div(a.H, kTH), a.B * a.nprob);
    return launch_k(sa_gate_kernel, grid, dim3(256), 0, stream, p);
}

}  // namespace cidnet
