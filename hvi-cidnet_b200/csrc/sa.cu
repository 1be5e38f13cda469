// SpatialAttention gates of the fork's MSSA variant (/root/reference/net/CIDNet_MSSA.py:10-25, used at :132-153):
//   avg = mean_C(x); mx = max_C(x); y = conv7x7(cat[avg, mx]) (zero padding 3, no bias); x * sigmoid(y)
// The per-pixel channel statistics (mean, max) are produced by the up block's GEMM epilogue (conv_gemm.cu, EPI_UP with
// ConvGemmLaunch::sa_stats: each epilogue thread holds its pixel's whole channel vector); what is left is ONE
// HBM-bound elementwise kernel per up-block pair:
//   sa_gate_kernel   reads the 7x7 neighbourhood of the statistics from a shared-memory tile, reads x, writes x
#include "sa.cuh"
#include "ptx_sm100.cuh"

namespace cidnet {

namespace {

constexpr int kMaxVec = 18;         // 144 channels / 8

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
