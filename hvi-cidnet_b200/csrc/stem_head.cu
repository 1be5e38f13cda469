// Fused first / last stages of CIDNet.forward (full resolution, CUDA cores, fp32 math).
//
// stem:  rgb (fp32 NCHW) -> HVIT -> { hvi (fp32 NCHW, kept for the global residual),
//                                      i_enc0 = IE_block0(I)   [B,H,W,36] NHWC act_t,
//                                      hv_0   = HVE_block0(hvi) [B,H,W,36] NHWC act_t }
//        = net/CIDNet.py:73-78 (HVIT, ReplicationPad2d(1)+conv3x3 1->36 and 3->36).
// head:  i_dec1, hv_1 (NHWC) -> ID_block0 (36->1) , HVD_block0 (36->2) -> cat + hvi -> PHVIT -> rgb
//        = net/CIDNet.py:115,117,119,120.
// Replicate padding == clamping the source coordinate, done while staging the tile.
#include "stem_head.cuh"
#include "hvi_math.cuh"

namespace cidnet {

#ifdef CIDNET_ACT_BF16
#define CIDNET_FHFMA_SH "fma.rn.f32.bf16"
#else
#define CIDNET_FHFMA_SH "fma.rn.f32.f16"
#endif
// acc0/1 += lo/hi(a) * lo/hi(b): mixed-precision FMA (SASS FHFMA), 16-bit operands stay packed
__device__ __forceinline__ void fhfma2(float& acc0, float& acc1, uint32_t a, uint32_t b) {
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
        "mov.b32 {al, ah}, %2;\n\t"
        "mov.b32 {bl, bh}, %3;\n\t"
        CIDNET_FHFMA_SH " %0, al, bl, %0;\n\t"
        CIDNET_FHFMA_SH " %1, ah, bh, %1;\n\t}"
        : "+f"(acc0), "+f"(acc1) : "r"(a), "r"(b));
}

static constexpr int kTile = 16;          // 16x16 output pixels per CTA, 256 threads
static constexpr int kHalo = kTile + 2;

// ------------------------------------------------------------------ stem ----
// weights in shared memory as [27 or 9 taps-inputs][36] fp32 (tap-input major so one
// thread reads 36 consecutive floats = 9 x LDS.128 broadcast per input value)
__global__ void __launch_bounds__(256)
stem_kernel(const float* __restrict__ rgb, float* __restrict__ hvi, act_t* __restrict__ i_enc0,
            act_t* __restrict__ hv_0, const float* __restrict__ w_hv /*[27][36]*/,
            const float* __restrict__ w_i /*[9][36]*/, const float* __restrict__ k_dev, float k_host,
            int H, int W, int pitch) {
    __shared__ float s_hvi[3][kHalo * kHalo];
    __shared__ __align__(16) float s_whv[27 * 36];
    __shared__ __align__(16) float s_wi[9 * 36];
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
    const int tid = threadIdx.x;
    const float k = k_dev ? __ldg(k_dev) : k_host;
    const long long hw = (long long)H * W;
    const float* img = rgb + (long long)b * 3 * hw;

    for (int i = tid; i < 27 * 36; i += 256) s_whv[i] = w_hv[i];
    for (int i = tid; i < 9 * 36; i += 256) s_wi[i] = w_i[i];
    for (int i = tid; i < kHalo * kHalo; i += 256) {
        const int hy = i / kHalo, hx = i - hy * kHalo;
        const int y = min(max(y0 + hy - 1, 0), H - 1);
        const int x = min(max(x0 + hx - 1, 0), W - 1);
        const long long o = (long long)y * W + x;
        float hh, vv, ii;
        hvit_px(img[o], img[o + hw], img[o + 2 * hw], k, hh, vv, ii);
        s_hvi[0][i] = hh; s_hvi[1][i] = vv; s_hvi[2][i] = ii;
    }
    __syncthreads();

    const int ty = tid / kTile, tx = tid - ty * kTile;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) return;
    const long long pix = (long long)y * W + x;
    {   // the HVI image itself (centre of the halo tile)
        const int c = (ty + 1) * kHalo + tx + 1;
        float* o = hvi + (long long)b * 3 * hw + pix;
        o[0] = s_hvi[0][c]; o[hw] = s_hvi[1][c]; o[2 * hw] = s_hvi[2][c];
    }
    float acc[36];
    // HVE_block0: 3 -> 36
#pragma unroll
    for (int j = 0; j < 36; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float v = s_hvi[c][(ty + t / 3) * kHalo + tx + t % 3];
            const float4* wr = reinterpret_cast<const float4*>(s_whv + (c * 9 + t) * 36);
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const float4 w4 = wr[j];
                acc[4 * j + 0] = fmaf(v, w4.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(v, w4.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(v, w4.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(v, w4.w, acc[4 * j + 3]);
            }
        }
    }
    {
        act_t* o = hv_0 + ((long long)b * hw + pix) * pitch;
        float pad[8] = {acc[32], acc[33], acc[34], acc[35], 0.f, 0.f, 0.f, 0.f};
        store8(o, acc); store8(o + 8, acc + 8); store8(o + 16, acc + 16); store8(o + 24, acc + 24);
        store8(o + 32, pad);
    }
    // IE_block0: 1 -> 36 (input = I channel)
#pragma unroll
    for (int j = 0; j < 36; ++j) acc[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float v = s_hvi[2][(ty + t / 3) * kHalo + tx + t % 3];
        const float4* wr = reinterpret_cast<const float4*>(s_wi + t * 36);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float4 w4 = wr[j];
            acc[4 * j + 0] = fmaf(v, w4.x, acc[4 * j + 0]);
            acc[4 * j + 1] = fmaf(v, w4.y, acc[4 * j + 1]);
            acc[4 * j + 2] = fmaf(v, w4.z, acc[4 * j + 2]);
            acc[4 * j + 3] = fmaf(v, w4.w, acc[4 * j + 3]);
        }
    }
    {
        act_t* o = i_enc0 + ((long long)b * hw + pix) * pitch;
        float pad[8] = {acc[32], acc[33], acc[34], acc[35], 0.f, 0.f, 0.f, 0.f};
        store8(o, acc); store8(o + 8, acc + 8); store8(o + 16, acc + 16); store8(o + 24, acc + 24);
        store8(o + 32, pad);
    }
}

const void* stem_kernel_func() { return reinterpret_cast<const void*>(&stem_kernel); }

int launch_stem(const StemArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.pitch == 40, CIDNET_ERR_INVALID, "stem: pitch must be 40");
    dim3 grid(ceil_div(a.W, kTile), ceil_div(a.H, kTile), a.B);
    stem_kernel<<<grid, 256, 0, stream>>>(a.rgb, a.hvi, a.i_enc0, a.hv_0, a.w_hv, a.w_i, a.k_dev, a.k_host,
                                          a.H, a.W, a.pitch);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

// ------------------------------------------------------------------ head ----
// smem tiles [18*18][40] act_t per branch; weights [9][36] per output channel, fp32.
__global__ void __launch_bounds__(256)
head_kernel(const act_t* __restrict__ i_dec1, const act_t* __restrict__ hv_1, const float* __restrict__ hvi,
            float* __restrict__ rgb, float* __restrict__ out_hvi_dbg, const float* __restrict__ w_i /*[9][36]*/,
            const float* __restrict__ w_hv /*[2][9][36]*/, const float* __restrict__ k_dev, PhvitParams pp,
            int H, int W, int pitch) {
    extern __shared__ __align__(16) uint8_t head_smem[];
    act_t* s_i = reinterpret_cast<act_t*>(head_smem);
    act_t* s_hv = s_i + kHalo * kHalo * 40;
    act_t* s_w = s_hv + kHalo * kHalo * 40;                               // [I | H | V][9 taps][40 ch] 16-bit, zero padded
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
    const int tid = threadIdx.x;
    const long long hw = (long long)H * W;
    if (k_dev) pp.k = __ldg(k_dev);

    for (int i = tid; i < 3 * 9 * 40; i += 256) {
        const int o = i / 360, r = i - o * 360, t = r / 40, c = r - t * 40;
        const float wv = c < 36 ? (o == 0 ? w_i[t * 36 + c] : w_hv[((o - 1) * 9 + t) * 36 + c]) : 0.f;
        s_w[i] = f2act(wv);
    }
    // stage both tiles with 16-byte vectors: 5 vectors per pixel per branch
    for (int i = tid; i < kHalo * kHalo * 5; i += 256) {
        const int p = i / 5, v = i - p * 5;
        const int hy = p / kHalo, hx = p - hy * kHalo;
        const int y = min(max(y0 + hy - 1, 0), H - 1);
        const int x = min(max(x0 + hx - 1, 0), W - 1);
        const long long g = (((long long)b * hw) + (long long)y * W + x) * pitch + v * 8;
        uint4 vi = *reinterpret_cast<const uint4*>(i_dec1 + g);
        uint4 vh = *reinterpret_cast<const uint4*>(hv_1 + g);
        if (v == 4) { vi.z = vi.w = 0u; vh.z = vh.w = 0u; }     // channels 36..39 are pitch padding nobody writes
        reinterpret_cast<uint4*>(s_i)[i] = vi;
        reinterpret_cast<uint4*>(s_hv)[i] = vh;
    }
    __syncthreads();

    const int ty = tid / kTile, tx = tid - ty * kTile;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) return;
    // 16-bit x 16-bit products accumulated in fp32 (FHFMA): operands stay packed, no conversions; the weights are
    // rounded to the activation type like every tensor-core layer's.  Two partial sums per output (even / odd lanes
    // of the packed pairs); channels 36..39 of data and weights are zero.
    float oi0 = 0.f, oi1 = 0.f, oh0 = 0.f, oh1 = 0.f, ov0 = 0.f, ov1 = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int p = (ty + t / 3) * kHalo + tx + t % 3;
        const uint4* di = reinterpret_cast<const uint4*>(s_i + p * 40);
        const uint4* dh = reinterpret_cast<const uint4*>(s_hv + p * 40);
        const uint4* wi = reinterpret_cast<const uint4*>(s_w + t * 40);
        const uint4* wh = reinterpret_cast<const uint4*>(s_w + 360 + t * 40);
        const uint4* wv = reinterpret_cast<const uint4*>(s_w + 720 + t * 40);
#pragma unroll
        for (int v = 0; v < 5; ++v) {
            const uint4 a = di[v], h = dh[v], x = wi[v], y = wh[v], z = wv[v];
            fhfma2(oi0, oi1, a.x, x.x); fhfma2(oi0, oi1, a.y, x.y); fhfma2(oi0, oi1, a.z, x.z); fhfma2(oi0, oi1, a.w, x.w);
            fhfma2(oh0, oh1, h.x, y.x); fhfma2(oh0, oh1, h.y, y.y); fhfma2(oh0, oh1, h.z, y.z); fhfma2(oh0, oh1, h.w, y.w);
            fhfma2(ov0, ov1, h.x, z.x); fhfma2(ov0, ov1, h.y, z.y); fhfma2(ov0, ov1, h.z, z.z); fhfma2(ov0, ov1, h.w, z.w);
        }
    }
    const float oi = oi0 + oi1, oh = oh0 + oh1, ov = ov0 + ov1;
    const long long pix = (long long)y * W + x;
    const float* hp = hvi + (long long)b * 3 * hw + pix;
    const float Hh = oh + hp[0], Vv = ov + hp[hw], Ii = oi + hp[2 * hw];   // cat([hv_0, i_dec0]) + hvi
    if (out_hvi_dbg) {
        float* d = out_hvi_dbg + (long long)b * 3 * hw + pix;
        d[0] = Hh; d[hw] = Vv; d[2 * hw] = Ii;
    }
    float r, g, bl;
    phvit_px(Hh, Vv, Ii, pp, r, g, bl);
    float* o = rgb + (long long)b * 3 * hw + pix;
    o[0] = r; o[hw] = g; o[2 * hw] = bl;
}

const void* head_kernel_func() { return reinterpret_cast<const void*>(&head_kernel); }

int launch_head(const HeadArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.pitch == 40, CIDNET_ERR_INVALID, "head: pitch must be 40");
    dim3 grid(ceil_div(a.W, kTile), ceil_div(a.H, kTile), a.B);
    PhvitParams pp{a.k_host, a.alpha_s, a.alpha, a.gated, a.gated2};
    const size_t smem = 2 * kHalo * kHalo * 40 * sizeof(act_t) + 3 * 9 * 40 * sizeof(act_t);
    static bool configured = false;
    if (!configured) {
        CIDNET_CUDA_OK(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    head_kernel<<<grid, 256, smem, stream>>>(a.i_dec1, a.hv_1, a.hvi, a.rgb, a.out_hvi_dbg, a.w_i, a.w_hv, a.k_dev, pp,
                                          a.H, a.W, a.pitch);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
