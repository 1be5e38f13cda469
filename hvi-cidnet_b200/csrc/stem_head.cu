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
#include "ptx_sm100.cuh"
#include "hvi_math.cuh"

#include <cstdlib>
#include <cstring>

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
static constexpr int kHeadRows = 21 * 16 + 1;   // head: pixel rows of a staged halo tile (18 x 18 = 324, padded to 21 m-tiles of 16 + 1)

// Both block0 stages are tiny GEMMs per pixel (stem: K = 27 / 9, N = 36; head: K = 9 x 36, N = 1 / 2).  The fp32-FMA
// kernels of round 1 spent 1300-1700 instructions per pixel on them (ncu: issue-bound at 1.2-1.5 TB/s, 4-5x their HBM
// floor; profiles/r01_ncu_full_head_fma_vs_mma.csv) and are gone.  m16n8k16 `mma.sync` (16-bit operands, fp32
// accumulate, like every other conv of the path) needs ~30 (stem) / ~110 (head) MMAs per 32 pixels instead; operands
// come straight from the halo tiles in shared memory (stem: gathered fp32 -> packed 16-bit A fragments; head: ldmatrix
// on the NHWC tile) and the weights are pre-arranged as per-lane B fragments.  tcgen05 would need an im2col copy of the
// tile in the UMMA layout plus TMEM round trips for N <= 36 -- the tensor pipe is idle either way, the instruction
// count is what matters here.
#ifdef CIDNET_ACT_BF16
#define CIDNET_MMA_16816 "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32"
#else
#define CIDNET_MMA_16816 "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32"
#endif
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(CIDNET_MMA_16816 " {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_act2(float lo, float hi) {
#ifdef CIDNET_ACT_BF16
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
#else
    __half2 h = __floats2half2_rn(lo, hi);
#endif
    return *reinterpret_cast<uint32_t*>(&h);
}

// stem K order: k 0..8 = taps of the I channel (c = 2), 9..17 = H (c = 0), 18..26 = V (c = 1), 27..31 = zero.
// With the I taps first, IE_block0 (1 -> 36 on I) re-uses the A fragment of HVE_block0's first k-step.
__device__ __forceinline__ int stem_k_to_ct(int k, int* c, int* t) {
    if (k >= 27) return 0;
    *c = k < 9 ? 2 : (k < 18 ? 0 : 1);
    *t = k % 9;
    return 1;
}

__global__ void __launch_bounds__(256)
stem_mma_kernel(const void* __restrict__ rgb_any, float* __restrict__ hvi, act_t* __restrict__ i_enc0,
                act_t* __restrict__ hv_0, const float* __restrict__ w_hv /*[27][36]*/,
                const float* __restrict__ w_i /*[9][36]*/, const float* __restrict__ k_dev, float k_host,
                int H, int W, int pitch, const uint2* __restrict__ bfrag /*[3][5][32], cidnet_pack_stem_bfrag*/,
                int in_u8, int h_src, int w_src, float gamma) {
    // in_u8 == 0: rgb_any = fp32 [B,3,H,W] planar.  in_u8 == 1: rgb_any = uint8 [B,h_src,w_src,3] (what a decoder yields):
    // ToTensor (/255), reflect padding of the bottom / right edge up to (H, W) and `** gamma` happen in the halo fill below
    // (data/eval_sets.py:22-27, eval.py:64) -- the padded fp32 image never exists in memory.
    const float* __restrict__ rgb = reinterpret_cast<const float*>(rgb_any);
    const uint8_t* __restrict__ rgb8 = reinterpret_cast<const uint8_t*>(rgb_any);
    __shared__ float s_hvi[3 * kHalo * kHalo];
    __shared__ __align__(16) uint2 s_bfrag[3][5][32];      // [HVE k-step 0, HVE k-step 1, IE][n8 tile][lane] = {b0, b1}
    __shared__ __align__(16) act_t s_out[8][32 * 40];      // per warp: [pixel][40 channels]
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const long long hw = (long long)H * W;
    const float* img = rgb + (long long)b * 3 * hw;

    if (tid < 3 * 5 * 32 / 2)      // per-lane B fragments, packed once at weight-finalisation time
        reinterpret_cast<uint4*>(&s_bfrag[0][0][0])[tid] = __ldg(reinterpret_cast<const uint4*>(bfrag) + tid);
    ptx::pdl_wait();               // constants above; the image, k and the outputs belong to the stream's earlier work
    ptx::pdl_trigger();
    const float k = k_dev ? __ldcg(k_dev) : k_host;
    // halo fill: 324 pixels on 256 threads -> a thread owns pixel `tid` and (tid < 68) pixel `tid + 256`.  All loads of both
    // pixels are issued before either is transformed (one exposed memory latency per CTA instead of two).
    float cr[2], cg[2], cb[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + j * 256;
        cr[j] = cg[j] = cb[j] = 0.f;
        if (i < kHalo * kHalo) {
            const int hy = i / kHalo, hx = i - hy * kHalo;
            const int y = min(max(y0 + hy - 1, 0), H - 1);
            const int x = min(max(x0 + hx - 1, 0), W - 1);
            if (in_u8) {
                const int sy = y < h_src ? y : 2 * (h_src - 1) - y;      // 'reflect': padded row h+i mirrors row h-2-i
                const int sx = x < w_src ? x : 2 * (w_src - 1) - x;
                const uint8_t* p = rgb8 + (((long long)b * h_src + sy) * w_src + sx) * 3;
                cr[j] = (float)__ldcg(p); cg[j] = (float)__ldcg(p + 1); cb[j] = (float)__ldcg(p + 2);
            } else {
                const long long o = (long long)y * W + x;
                cr[j] = __ldcg(img + o); cg[j] = __ldcg(img + o + hw); cb[j] = __ldcg(img + o + 2 * hw);   // L2 loads: see ptx::pdl_wait
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int i = tid + j * 256;
        if (i < kHalo * kHalo) {
            float r = cr[j], g_ = cg[j], b_ = cb[j], hh, vv, ii;
            if (in_u8) {
                r = __fdiv_rn(r, 255.0f); g_ = __fdiv_rn(g_, 255.0f); b_ = __fdiv_rn(b_, 255.0f);
                if (gamma != 1.0f) { r = powf(r, gamma); g_ = powf(g_, gamma); b_ = powf(b_, gamma); }
            }
            hvit_px(r, g_, b_, k, hh, vv, ii);
            s_hvi[i] = hh; s_hvi[kHalo * kHalo + i] = vv; s_hvi[2 * kHalo * kHalo + i] = ii;
        }
    }
    __syncthreads();

    const int ty = tid / kTile, tx = tid - ty * kTile;      // this thread's own pixel: warp w owns tile rows 2w, 2w + 1
    const int y = y0 + ty, x = x0 + tx;
    const bool inside = y < H && x < W;
    const long long pix = (long long)y * W + x;
    if (inside) {   // the HVI image itself (centre of the halo tile)
        const int c = (ty + 1) * kHalo + tx + 1;
        float* o = hvi + (long long)b * 3 * hw + pix;
        o[0] = s_hvi[c]; o[hw] = s_hvi[kHalo * kHalo + c]; o[2 * hw] = s_hvi[2 * kHalo * kHalo + c];
    }
    // A fragments: a0 = (row g, k0, k0+1), a1 = (row g+8, ...), a2 = (row g, k0+8, k0+9), a3 = (row g+8, ...), k0 = 16 ks + 2 tig
    int koff[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kk = (j >> 2) * 16 + 2 * tig + (j & 1) + ((j >> 1) & 1) * 8;
        int c = 0, t = 0;
        koff[j] = stem_k_to_ct(kk, &c, &t) ? c * kHalo * kHalo + (t / 3) * kHalo + (t % 3) : -1;
    }
    uint32_t afrag[2][2][4];                                // [m tile][k-step][a0..a3]
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const int base = (2 * warp + m) * kHalo + g;        // halo-tile offset of (tile row, column g), tap (0, 0)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            float v[8];                                     // (k0, k0+1, k0+8, k0+9) x (row g, row g+8)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int off = koff[ks * 4 + e];
                v[e] = off >= 0 ? s_hvi[off + base] : 0.f;
                v[4 + e] = off >= 0 ? s_hvi[off + base + 8] : 0.f;
            }
            afrag[m][ks][0] = pack_act2(v[0], v[1]);
            afrag[m][ks][1] = pack_act2(v[4], v[5]);
            afrag[m][ks][2] = pack_act2(v[2], v[3]);
            afrag[m][ks][3] = pack_act2(v[6], v[7]);
        }
    }
    act_t* so = s_out[warp];
#pragma unroll
    for (int conv = 0; conv < 2; ++conv) {                  // 0: HVE_block0 (3 -> 36), 1: IE_block0 (1 -> 36)
        float acc[2][5][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 5; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
#pragma unroll
        for (int n = 0; n < 5; ++n) {
            if (conv == 0) {
                const uint2 b0 = s_bfrag[0][n][lane], b1 = s_bfrag[1][n][lane];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    mma16816(acc[m][n], afrag[m][0], b0.x, b0.y);
                    mma16816(acc[m][n], afrag[m][1], b1.x, b1.y);
                }
            } else {
                const uint2 b2 = s_bfrag[2][n][lane];
#pragma unroll
                for (int m = 0; m < 2; ++m) mma16816(acc[m][n], afrag[m][0], b2.x, b2.y);
            }
        }
        // C fragment: c0, c1 = (row g, channels 8n + 2 tig, +1), c2, c3 = (row g + 8, ...) -> per-warp staging, then
        // every lane stores its own pixel's 40 channels with 16-byte vectors (channels 36..39 are exact zeros)
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int n = 0; n < 5; ++n) {
                *reinterpret_cast<uint32_t*>(so + (m * 16 + g) * 40 + n * 8 + 2 * tig) = pack_act2(acc[m][n][0], acc[m][n][1]);
                *reinterpret_cast<uint32_t*>(so + (m * 16 + g + 8) * 40 + n * 8 + 2 * tig) = pack_act2(acc[m][n][2], acc[m][n][3]);
            }
        __syncwarp();
        if (inside) {
            act_t* o = (conv == 0 ? hv_0 : i_enc0) + ((long long)b * hw + pix) * pitch;
            const uint4* src = reinterpret_cast<const uint4*>(so + lane * 40);
#pragma unroll
            for (int v4 = 0; v4 < 5; ++v4) reinterpret_cast<uint4*>(o)[v4] = src[v4];
        }
    }
}

// Per-lane B fragments of the m16n8k16 MMAs, built on the host when the weights are finalised.
//   b0 = (k = 2 tig, 2 tig + 1; n = g), b1 = (k + 8; n = g) with g = lane / 4, tig = lane % 4.
static uint32_t host_pack2(float lo, float hi) {
    const act_t a = f2act(lo), b = f2act(hi);
    uint16_t ua, ub;
    memcpy(&ua, &a, 2); memcpy(&ub, &b, 2);
    return (uint32_t)ua | ((uint32_t)ub << 16);
}
void pack_stem_bfrag(const float* w_hv /*[27][36] (c*9+t major)*/, const float* w_i /*[9][36]*/, uint2* out /*[3][5][32]*/) {
    for (int i = 0; i < 3 * 5 * 32; ++i) {
        const int slot = i / 160, n = (i / 32) % 5, l = i & 31;
        const int o = n * 8 + (l >> 2), tg = l & 3;
        float wv[4];
        for (int e = 0; e < 4; ++e) {
            const int kk = (slot == 1 ? 16 : 0) + 2 * tg + (e & 1) + (e >> 1) * 8;
            float v = 0.f;
            if (o < 36 && kk < 27) {
                const int c = kk < 9 ? 2 : (kk < 18 ? 0 : 1), t = kk % 9;      // stem K order: I taps, H taps, V taps
                if (slot < 2) v = w_hv[(c * 9 + t) * 36 + o];
                else if (c == 2) v = w_i[t * 36 + o];                          // IE_block0 sees only the I taps (k < 9)
            }
            wv[e] = v;
        }
        out[i] = make_uint2(host_pack2(wv[0], wv[1]), host_pack2(wv[2], wv[3]));
    }
}
// head: the taps live in the N dimension (see head_mma_kernel).  n-tiles: 0, 1 = ID_block0 (n = tap 0..8), 2, 3, 4 =
// HVD_block0 (n = out * 9 + tap, out 0 = H, 1 = V); layout [n-tile 5][k-step 3][lane 32].
void pack_head_bfrag(const float* w_i /*[9][36]*/, const float* w_hv /*[2][9][36]*/, uint2* out /*[5][3][32]*/) {
    for (int i = 0; i < 5 * 3 * 32; ++i) {
        const int l = i & 31, ks = (i >> 5) % 3, nt = i / 96;
        const int g = l >> 2, tg = l & 3;
        const bool is_i = nt < 2;
        const int n = (is_i ? nt : nt - 2) * 8 + g;             // column inside the branch
        float wv[4];
        for (int e = 0; e < 4; ++e) {
            const int c = ks * 16 + 2 * tg + (e & 1) + (e >> 1) * 8;
            float v = 0.f;
            if (c < 36) {
                if (is_i && n < 9) v = w_i[n * 36 + c];
                if (!is_i && n < 18) v = w_hv[n * 36 + c];      // [(out * 9 + tap)][36]
            }
            wv[e] = v;
        }
        out[i] = make_uint2(host_pack2(wv[0], wv[1]), host_pack2(wv[2], wv[3]));
    }
}

const void* stem_kernel_func() { return reinterpret_cast<const void*>(&stem_mma_kernel); }

int launch_stem(const StemArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.pitch == 40, CIDNET_ERR_INVALID, "stem: pitch must be 40");
    dim3 grid(ceil_div(a.W, kTile), ceil_div(a.H, kTile), a.B);
    return launch_k(stem_mma_kernel, grid, dim3(256), 0, stream, a.rgb, a.hvi, a.i_enc0, a.hv_0, a.w_hv, a.w_i, a.k_dev,
                    a.k_host, a.H, a.W, a.pitch, a.bfrag, a.in_u8, a.h_src, a.w_src, a.gamma);
}

// ------------------------------------------------------------------ head ----
// head on warp-level tensor cores, TAPS IN THE N DIMENSION.  Both convs have 1 / 2 output channels, so instead of one
// 16-pixel x 8 MMA per tap (9 x 3 k-steps x 2 branches = 54 ldmatrix.x4 + 54 MMAs per 16 pixels, of which 5 / 8 of every
// MMA's columns are padding: ncu on that version showed the LSU / shared-memory pipe at 77 %), the per-pixel partial products
//     P[halo pixel][tap]        = sum_c  i_dec1[pixel][c] * W_I[tap][c]              (9 columns)
//     P[halo pixel][9 + 9o + t] = sum_c  hv_1[pixel][c]   * W_HV[o][t][c]            (18 columns, o = H, V)
// are computed ONCE per halo pixel -- A = the NHWC halo tile itself (row = pixel, 80-byte pitch -> conflict-free ldmatrix),
// 3 ldmatrix.x4 per 16 halo pixels and branch feed 2 + 3 n-tiles -- and every output pixel then sums its nine shifted
// entries per output from a 27-float-pitch table (odd pitch: conflict-free LDS.32).  Per 16 x 16 tile: 126 ldmatrix.x4 and
// 315 MMAs instead of 864 and 864.  Channels 36..39 of the tiles are zeroed when staged, channels 40..47 of the third k-step
// belong to the next pixel and meet zero weights.
__global__ void __launch_bounds__(256, 2)
head_mma_kernel(const act_t* __restrict__ i_dec1, const act_t* __restrict__ hv_1, const float* __restrict__ hvi,
                void* __restrict__ rgb_any, float* __restrict__ out_hvi_dbg, const float* __restrict__ w_i /*[9][36]*/,
                const float* __restrict__ w_hv /*[2][9][36]*/, const float* __restrict__ k_dev, PhvitParams pp,
                int H, int W, int pitch, const uint2* __restrict__ bfrag /*[5][3][32], cidnet_pack_head_bfrag*/,
                int out_u8, int h_dst, int w_dst) {
    // out_u8 == 0: rgb_any = fp32 [B,3,H,W] planar.  out_u8 == 1: rgb_any = uint8 [B,h_dst,w_dst,3]: clamp(0,1), crop to
    // [:h_dst,:w_dst] and ToPILImage's mul(255).byte() (eval.py:69-73) happen in the store below.
    float* __restrict__ rgb = reinterpret_cast<float*>(rgb_any);
    extern __shared__ __align__(16) uint8_t head_smem[];
    act_t* s_i = reinterpret_cast<act_t*>(head_smem);                        // [21 m-tiles x 16 = 336 pixel rows + 1 pad pixel][40]
    act_t* s_hv = s_i + kHeadRows * 40;
    uint2* s_bfrag = reinterpret_cast<uint2*>(s_hv + kHeadRows * 40);        // [5 n-tiles][3 k-steps][32 lanes]
    float* s_P = reinterpret_cast<float*>(s_bfrag + 5 * 3 * 32);            // [336 halo pixels][27]: I taps | H taps | V taps
    const int b = blockIdx.z;
    const int y0 = blockIdx.y * kTile, x0 = blockIdx.x * kTile;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tig = lane & 3;
    const long long hw = (long long)H * W;

    for (int i = tid; i < 5 * 3 * 32 / 2; i += 256)          // per-lane B fragments, packed once at weight-finalisation time
        reinterpret_cast<uint4*>(s_bfrag)[i] = __ldg(reinterpret_cast<const uint4*>(bfrag) + i);
    ptx::pdl_wait();               // constants above; the activations and k belong to the stream's earlier work
    ptx::pdl_trigger();
    if (k_dev) pp.k = __ldcg(k_dev);
    // stage both halo tiles: a tile row is 18 pixels x 5 16-byte vectors; threads 0..179 copy two rows per step with a
    // fixed (pixel, vector) each -- no index arithmetic inside the loop
    if (tid < 180) {
        const int r = tid / 90, tv = tid - r * 90, px = tv / 5, v = tv - px * 5;
        const int x = min(max(x0 + px - 1, 0), W - 1);
        const long long col = (long long)x * pitch + v * 8;
        // cp.async: all 9 rows x 2 tensors of a thread are in flight at once, no registers in between; the last vector of
        // a pixel copies 8 bytes and zero-fills the rest (channels 36..39 are pitch padding nobody writes)
        const uint32_t nsrc = v == 4 ? 8u : 16u;
        const uint32_t di = static_cast<uint32_t>(__cvta_generic_to_shared(s_i)) + (uint32_t)tv * 16u;
        const uint32_t dh = static_cast<uint32_t>(__cvta_generic_to_shared(s_hv)) + (uint32_t)tv * 16u;
#pragma unroll
        for (int hy = r; hy < kHalo; hy += 2) {
            const int y = min(max(y0 + hy - 1, 0), H - 1);
            const long long gi = ((long long)b * hw + (long long)y * W) * pitch + col;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(di + (uint32_t)hy * 1440u), "l"(i_dec1 + gi), "r"(nsrc) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dh + (uint32_t)hy * 1440u), "l"(hv_1 + gi), "r"(nsrc) : "memory");
        }
    } else if (tid < 180 + 13 * 5) {                                 // the 13 pad pixel rows behind each tile (m-tile 20 reads them)
        reinterpret_cast<uint4*>(s_i)[kHalo * 90 + tid - 180] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(s_hv)[kHalo * 90 + tid - 180] = make_uint4(0, 0, 0, 0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // this thread's own pixel (phase 2): its three fp32 residual values are requested NOW, so their L2 / DRAM latency runs
    // under the tile staging and the MMA phase instead of in front of the PHVIT arithmetic
    const int ty = tid / kTile, tx = tid - ty * kTile;
    const int y = y0 + ty, x = x0 + tx;
    const bool inside = y < H && x < W;
    const long long pix = (long long)y * W + x;
    float res_h = 0.f, res_v = 0.f, res_i = 0.f;
    if (inside) {
        const float* hp = hvi + (long long)b * 3 * hw + pix;
        res_h = __ldcg(hp); res_v = __ldcg(hp + hw); res_i = __ldcg(hp + 2 * hw);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ldmatrix lane address: matrices 0 / 1 = rows 0-7 / 8-15 at k 0-7, matrices 2 / 3 = the same rows at k 8-15
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lk = (lane >> 4) * 8;
    const uint32_t si_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_i));
    const uint32_t shv_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_hv));
    // ---- phase 1: partial products of every halo pixel (21 m-tiles of 16 pixel rows, round-robin over the 8 warps)
    uint2 bfr[5][3];                                          // this lane's B fragments: loaded once, used by all its m-tiles
#pragma unroll
    for (int nt = 0; nt < 5; ++nt)
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) bfr[nt][ks] = s_bfrag[(nt * 3 + ks) * 32 + lane];
    for (int mt = warp; mt < 21; mt += 8) {
        uint32_t ai[3][4], ah[3][4];
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            const uint32_t off = (uint32_t)(((mt * 16 + lrow) * 40 + ks * 16 + lk) * 2);
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(ai[ks][0]), "=r"(ai[ks][1]), "=r"(ai[ks][2]), "=r"(ai[ks][3]) : "r"(si_base + off));
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                         : "=r"(ah[ks][0]), "=r"(ah[ks][1]), "=r"(ah[ks][2]), "=r"(ah[ks][3]) : "r"(shv_base + off));
        }
        float* p0 = s_P + (mt * 16 + g) * 27;                 // C fragment: c0, c1 = (row g, columns 2 tig, +1), c2, c3 = (row g + 8, ..)
        float* p1 = p0 + 8 * 27;
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                mma16816(acc, nt < 2 ? ai[ks] : ah[ks], bfr[nt][ks].x, bfr[nt][ks].y);
            }
            const int n = (nt < 2 ? nt : nt - 2) * 8 + 2 * tig;       // column inside the branch
            const int lim = nt < 2 ? 9 : 18, col = (nt < 2 ? 0 : 9) + n;
            if (n < lim) { p0[col] = acc[0]; p1[col] = acc[2]; }
            if (n + 1 < lim) { p0[col + 1] = acc[1]; p1[col + 1] = acc[3]; }
        }
    }
    __syncthreads();
    // ---- phase 2: every thread sums the nine shifted entries of its own pixel, per output
    if (!inside) return;
    float oi = 0.f, oh = 0.f, ov = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float* p = s_P + ((ty + t / 3) * kHalo + tx + t % 3) * 27 + t;
        oi += p[0]; oh += p[9]; ov += p[18];
    }
    const float Hh = oh + res_h, Vv = ov + res_v, Ii = oi + res_i;   // cat([hv_0, i_dec0]) + hvi
    if (out_hvi_dbg) {
        float* d = out_hvi_dbg + (long long)b * 3 * hw + pix;
        d[0] = Hh; d[hw] = Vv; d[2 * hw] = Ii;
    }
    float r, gg, bl;
    phvit_px(Hh, Vv, Ii, pp, r, gg, bl);
    if (out_u8) {
        if (y < h_dst && x < w_dst) {
            uint8_t* o8 = reinterpret_cast<uint8_t*>(rgb_any) + (((long long)b * h_dst + y) * w_dst + x) * 3;
            o8[0] = (uint8_t)__float2int_rz(__fmul_rn(fminf(fmaxf(r, 0.0f), 1.0f), 255.0f));      // NaN -> 0 like .byte()
            o8[1] = (uint8_t)__float2int_rz(__fmul_rn(fminf(fmaxf(gg, 0.0f), 1.0f), 255.0f));
            o8[2] = (uint8_t)__float2int_rz(__fmul_rn(fminf(fmaxf(bl, 0.0f), 1.0f), 255.0f));
        }
        return;
    }
    float* o = rgb + (long long)b * 3 * hw + pix;
    o[0] = r; o[hw] = gg; o[2 * hw] = bl;
}

const void* head_kernel_func() { return reinterpret_cast<const void*>(&head_mma_kernel); }

int launch_head(const HeadArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.pitch == 40, CIDNET_ERR_INVALID, "head: pitch must be 40");
    dim3 grid(ceil_div(a.W, kTile), ceil_div(a.H, kTile), a.B);
    PhvitParams pp{a.k_host, a.alpha_s, a.alpha, a.gated, a.gated2};
    const size_t smem = 2 * kHeadRows * 40 * sizeof(act_t) + 5 * 3 * 32 * sizeof(uint2) + 336 * 27 * sizeof(float);
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(head_mma_kernel), (int)smem);
    if (rc) return rc;
    return launch_k(head_mma_kernel, grid, dim3(256), smem, stream, a.i_dec1, a.hv_1, a.hvi, a.rgb, a.out_hvi_dbg, a.w_i,
                    a.w_hv, a.k_dev, pp, a.H, a.W, a.pitch, a.bfrag, a.out_u8, a.h_dst, a.w_dst);
}

}  // namespace cidnet
