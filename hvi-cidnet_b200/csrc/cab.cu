// CAB (channel cross-attention, net/LCA.py:19-41) in three kernels:
//
//   dw3x3_kernel    : depthwise 3x3 (zero pad) of q_pre | k_pre | v_pre, NHWC in, NHWC out.
//                     Sliding window down the rows: one thread owns (column x, 8 channels), keeps
//                     the 3x3 neighbourhood and its 72 weights in registers and reads each new
//                     row straight from global memory (the +-1 column neighbours are L1 hits).
//                     Also accumulates sum(q^2), sum(k^2) per channel (F.normalize denominators).
//   gram_kernel     : G[Cq, Ck] = sum_pixels q k^T on the tensor cores.  The NHWC tiles
//                     [64 pixels][64 channels] land by TMA (SWIZZLE_128B) exactly in the UMMA
//                     *MN-major* canonical layout, so the contraction over pixels needs no
//                     transposition: tcgen05.mma kind::f16 with a_major = b_major = MN, fp32
//                     accumulation in TMEM over the CTA's whole pixel range (split-K across CTAs),
//                     then the per-head 18x18 diagonal blocks are atomically added to global.
//   cab_fold_kernel : L2-normalise (eps 1e-12), temperature, softmax, and fold attn into
//                     project_out:  M_b = W_o * blockdiag(attn_b)   (one CxC matrix per image, written
//                     in the packed-weight layout of the conv GEMM), so project_out(attn @ v)
//                     becomes a per-image 1x1 conv on v (identity verified to 6e-7).
#include "cab.cuh"
#include "ptx_sm100.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace cidnet {

// ------------------------------------------------------------- depthwise ----
static constexpr int kDwThreads = 128;
static constexpr int kDwRows = 32;     // rows per CTA strip

#ifdef CIDNET_ACT_BF16
#define CIDNET_FHFMA "fma.rn.f32.bf16"
#else
#define CIDNET_FHFMA "fma.rn.f32.f16"
#endif
// acc0/1 += lo/hi(a) * lo/hi(b): mixed-precision FMA (SASS FHFMA), operands stay packed 16-bit
__device__ __forceinline__ void fhfma2(float& acc0, float& acc1, uint32_t a, uint32_t b) {
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
        "mov.b32 {al, ah}, %2;\n\t"
        "mov.b32 {bl, bh}, %3;\n\t"
        CIDNET_FHFMA " %0, al, bl, %0;\n\t"
        CIDNET_FHFMA " %1, ah, bh, %1;\n\t}"
        : "+f"(acc0), "+f"(acc1) : "r"(a), "r"(b));
}
__device__ __forceinline__ void fhfma8(float* acc, const uint4& t, const uint4& w) {
    fhfma2(acc[0], acc[1], t.x, w.x);
    fhfma2(acc[2], acc[3], t.y, w.y);
    fhfma2(acc[4], acc[5], t.z, w.z);
    fhfma2(acc[6], acc[7], t.w, w.w);
}

// one thread = (column x, 8 channels): 9 packed weight vectors + a raw 3x3 window in registers,
// FHFMA (16-bit x 16-bit + fp32) so no conversion instructions are needed
__global__ void __launch_bounds__(kDwThreads, 3)
dw3x3_f32acc_kernel(const Dw3Args a) {
    __shared__ float s_ssq[2 * 144];
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int nv = a.nv;                                   // 16-byte vectors per pixel (all segments)
    const int idx = blockIdx.x * kDwThreads + threadIdx.x; // vector index along the row
    const int x = idx / nv, v = idx - x * nv;
    const bool active = x < a.W;
    const int seg = active ? v / a.seg_vecs : 0;           // 0 = q, 1 = k, 2 = v
    const int c0 = (v - seg * a.seg_vecs) * 8;             // channel within the segment
    const long long hw = (long long)a.H * a.W;
    const act_t* src = a.src[prob][seg] + (long long)b * hw * a.src_pitch + c0;
    act_t* dst = a.dst[prob] + (long long)b * hw * a.dst_pitch + seg * a.seg_vecs * 8 + c0;
    const int y0 = blockIdx.y * kDwRows;
    const int y1 = min(y0 + kDwRows, a.H);

    for (int i = threadIdx.x; i < 2 * 144; i += kDwThreads) s_ssq[i] = 0.f;
    __syncthreads();

    uint4 w[9];
    {
        const float* wp = a.w[prob] + seg * a.seg_vecs * 8 + c0;   // [9][nv*8] tap major
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            act_t* h = reinterpret_cast<act_t*>(&w[t]);
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = f2act(active ? __ldg(wp + t * nv * 8 + e) : 0.f);
        }
    }
    float ssq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ssq[e] = 0.f;

    if (active) {
        const bool has_l = x > 0, has_r = x + 1 < a.W;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        uint4 win0[3], win1[3], win2[3];
        auto load_row = [&](int y, uint4* r) {
            r[0] = r[1] = r[2] = zero4;
            if (y < 0 || y >= a.H) return;
            const act_t* p = src + ((long long)y * a.W + x) * a.src_pitch;
            r[1] = *reinterpret_cast<const uint4*>(p);
            if (has_l) r[0] = *reinterpret_cast<const uint4*>(p - a.src_pitch);
            if (has_r) r[2] = *reinterpret_cast<const uint4*>(p + a.src_pitch);
        };
        auto step = [&](const uint4* r0, const uint4* r1, uint4* r2, int y) {
            load_row(y + 1, r2);
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                fhfma8(acc, r0[c], w[c]);
                fhfma8(acc, r1[c], w[3 + c]);
                fhfma8(acc, r2[c], w[6 + c]);
            }
            uint4 raw;
            act_t* ov = reinterpret_cast<act_t*>(&raw);
#pragma unroll
            for (int e = 0; e < 8; ++e) ov[e] = f2act(acc[e]);
            if (y >= a.stat_y0 && y < a.stat_y1) fhfma8(ssq, raw, raw);   // sum of squares of the ROUNDED values
            *reinterpret_cast<uint4*>(dst + ((long long)y * a.W + x) * a.dst_pitch) = raw;
        };
        load_row(y0 - 1, win0);
        load_row(y0, win1);
        for (int y = y0; y < y1; y += 3) {          // window roles rotate, no register moves
            step(win0, win1, win2, y);
            if (y + 1 < y1) step(win1, win2, win0, y + 1);
            if (y + 2 < y1) step(win2, win0, win1, y + 2);
        }
        if (seg < 2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_ssq[seg * 144 + c0 + e], ssq[e]);
        }
    }
    __syncthreads();
    // one global atomic per channel per CTA
    const int Cp = a.seg_vecs * 8;
    for (int i = threadIdx.x; i < 2 * Cp; i += kDwThreads) {
        const int sg = i / Cp, c = i - sg * Cp;
        const float val = s_ssq[sg * 144 + c];
        if (val != 0.f) atomicAdd((sg == 0 ? a.sq[prob] : a.sk[prob]) + (long long)b * Cp + c, val);
    }
}


#ifndef CIDNET_ACT_BF16
// fp16 build: the same sliding window with PACKED fp16 math (HFMA2: two multiply-adds per issue slot,
// accumulators are the packed output vector -> no conversions, ~100 registers instead of 162 so five
// CTAs fit an SM and hide the global-load latency the fp32-accumulate kernel was bound by).  The 9-tap
// sum is rounded to fp16 after every tap (two chains); end-to-end effect measured against the fp32
// oracle: max-abs 1.0e-4 -> 1.2e-4 (see iel.cu v5).  sum q^2 / sum k^2 stay fp32 (FHFMA).
__device__ __forceinline__ uint32_t hfma2u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hmul2u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hadd2u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint4 hmul8(const uint4& t, const uint4& w) {
    return make_uint4(hmul2u(t.x, w.x), hmul2u(t.y, w.y), hmul2u(t.z, w.z), hmul2u(t.w, w.w));
}
__device__ __forceinline__ void hfma8(uint4& acc, const uint4& t, const uint4& w) {
    acc.x = hfma2u(t.x, w.x, acc.x); acc.y = hfma2u(t.y, w.y, acc.y);
    acc.z = hfma2u(t.z, w.z, acc.z); acc.w = hfma2u(t.w, w.w, acc.w);
}

// kWsmem: the 9 weight vectors are re-read from shared memory at every row (9 conflict-free LDS.128) instead of
// living in 36 registers -> ~100 registers, 5 CTAs / SM WITH the one-row-ahead prefetch.  ncu on the register-weight
// kernel (L1, cfg 2): issue slots 25 % busy, 21 % of the warp slots occupied, long_scoreboard 9.6 stall cycles per
// issue -> bound by the latency of its own global loads, i.e. by how many rows are in flight per SM.
template <bool kPrefetch, int kMinBlocks, bool kWsmem = false>
__global__ void __launch_bounds__(kDwThreads, kMinBlocks)
dw3x3_kernel(const Dw3Args a) {
    __shared__ float s_ssq[2 * 144];
    __shared__ __align__(16) act_t s_w[kWsmem ? 9 * 432 : 8];
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int nv = a.nv;                                   // 16-byte vectors per pixel (all segments)
    const int idx = blockIdx.x * kDwThreads + threadIdx.x; // vector index along the row
    const int x = idx / nv, v = idx - x * nv;
    const bool active = x < a.W;
    const int seg = active ? v / a.seg_vecs : 0;           // 0 = q, 1 = k, 2 = v
    const int c0 = (v - seg * a.seg_vecs) * 8;             // channel within the segment
    const long long hw = (long long)a.H * a.W;
    const act_t* src = a.src[prob][seg] + (long long)b * hw * a.src_pitch + c0;
    act_t* dst = a.dst[prob] + (long long)b * hw * a.dst_pitch + seg * a.seg_vecs * 8 + c0;
    const int y0 = blockIdx.y * kDwRows;
    const int y1 = min(y0 + kDwRows, a.H);

    for (int i = threadIdx.x; i < 2 * 144; i += kDwThreads) s_ssq[i] = 0.f;
    if (kWsmem) {
        const float* wp = a.w[prob];                                // [9][nv*8] tap major, fp32
        for (int i = threadIdx.x; i < 9 * nv * 8; i += kDwThreads) s_w[i] = f2act(__ldg(wp + i));
    }
    __syncthreads();

    uint4 wreg[kWsmem ? 1 : 9];
    if (!kWsmem) {
        const float* wp = a.w[prob] + seg * a.seg_vecs * 8 + c0;   // [9][nv*8] tap major
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            act_t* h = reinterpret_cast<act_t*>(&wreg[t]);
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = f2act(active ? __ldg(wp + t * nv * 8 + e) : 0.f);
        }
    }
    const uint4* wsm = reinterpret_cast<const uint4*>(s_w) + (active ? v : 0);     // tap t: wsm[t * nv]
    // asm volatile: the loads must stay where they are used (hoisted out of the row loop they would occupy the
    // 36 registers this variant exists to free)
    const uint32_t wsm_addr = ptx::smem_u32(wsm);
    auto W9 = [&](int t) -> uint4 {
        if (!kWsmem) return wreg[kWsmem ? 0 : t];
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(wsm_addr + (uint32_t)(t * nv * 16)));
        return r;
    };
    float ssq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ssq[e] = 0.f;

    if (active) {
        const bool has_l = x > 0, has_r = x + 1 < a.W;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        uint4 win0[3], win1[3], win2[3], pend[3];
        // loads run one row ahead of their use
        auto issue = [&](int y) {
            pend[0] = pend[1] = pend[2] = zero4;
            if (y < 0 || y >= a.H) return;
            const act_t* p = src + ((long long)y * a.W + x) * a.src_pitch;
            pend[1] = *reinterpret_cast<const uint4*>(p);
            if (has_l) pend[0] = *reinterpret_cast<const uint4*>(p - a.src_pitch);
            if (has_r) pend[2] = *reinterpret_cast<const uint4*>(p + a.src_pitch);
        };
        auto step = [&](const uint4* r0, const uint4* r1, uint4* r2, int y) {
            if (kPrefetch) { r2[0] = pend[0]; r2[1] = pend[1]; r2[2] = pend[2]; issue(y + 2); }   // row y+1 arrives, y+2 leaves
            else           { issue(y + 1); r2[0] = pend[0]; r2[1] = pend[1]; r2[2] = pend[2]; }
            uint4 pa = hmul8(r0[0], W9(0));
            uint4 pb = hmul8(r2[0], W9(6));
            hfma8(pa, r0[1], W9(1)); hfma8(pb, r2[1], W9(7));
            hfma8(pa, r0[2], W9(2)); hfma8(pb, r2[2], W9(8));
            hfma8(pa, r1[0], W9(3));
            hfma8(pa, r1[1], W9(4));
            hfma8(pa, r1[2], W9(5));
            const uint4 raw = make_uint4(hadd2u(pa.x, pb.x), hadd2u(pa.y, pb.y), hadd2u(pa.z, pb.z), hadd2u(pa.w, pb.w));
            if (y >= a.stat_y0 && y < a.stat_y1) fhfma8(ssq, raw, raw);   // sum of squares of the STORED values
            *reinterpret_cast<uint4*>(dst + ((long long)y * a.W + x) * a.dst_pitch) = raw;
        };
        issue(y0 - 1); win0[0] = pend[0]; win0[1] = pend[1]; win0[2] = pend[2];
        issue(y0);     win1[0] = pend[0]; win1[1] = pend[1]; win1[2] = pend[2];
        if (kPrefetch) issue(y0 + 1);
        for (int y = y0; y < y1; y += 3) {          // window roles rotate, no register moves
            step(win0, win1, win2, y);
            if (y + 1 < y1) step(win1, win2, win0, y + 1);
            if (y + 2 < y1) step(win2, win0, win1, y + 2);
        }
        if (seg < 2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_ssq[seg * 144 + c0 + e], ssq[e]);
        }
    }
    __syncthreads();
    // one global atomic per channel per CTA
    const int Cp = a.seg_vecs * 8;
    for (int i = threadIdx.x; i < 2 * Cp; i += kDwThreads) {
        const int sg = i / Cp, c = i - sg * Cp;
        const float val = s_ssq[sg * 144 + c];
        if (val != 0.f) atomicAdd((sg == 0 ? a.sq[prob] : a.sk[prob]) + (long long)b * Cp + c, val);
    }
}
#endif

#ifndef CIDNET_ACT_BF16
// ------------------------------------------------------------------------------------------------
// dw3x3 with a per-thread cp.async ring (CIDNET_DW_VARIANT=8): the same thread <-> (column, 8 channels) mapping and
// sliding window as dw3x3_kernel, but the three 16-byte loads of every input row (left, centre, right) are issued
// kDepth rows ahead with cp.async (zero fill outside the image = the conv's padding) into the thread's own shared-
// memory slots -- no registers held by loads in flight, no barriers (a thread only reads what it requested itself).
// ------------------------------------------------------------------------------------------------
template <int kDepth, int kMinBlocks>
__global__ void __launch_bounds__(kDwThreads, kMinBlocks)
dw3x3_cpasync_kernel(const Dw3Args a) {
    extern __shared__ __align__(16) uint8_t dw_smem[];
    uint4* ring = reinterpret_cast<uint4*>(dw_smem);                       // [kDepth][3][kDwThreads]
    act_t* s_w = reinterpret_cast<act_t*>(ring + kDepth * 3 * kDwThreads);  // [9][nv * 8]
    __shared__ float s_ssq[2 * 144];
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int nv = a.nv;
    const int idx = blockIdx.x * kDwThreads + threadIdx.x;
    const int x = idx / nv, v = idx - x * nv;
    const bool active = x < a.W;
    const int seg = active ? v / a.seg_vecs : 0;
    const int c0 = (v - seg * a.seg_vecs) * 8;
    const long long hw = (long long)a.H * a.W;
    const act_t* src = a.src[prob][seg] + (long long)b * hw * a.src_pitch + c0;
    act_t* dst = a.dst[prob] + (long long)b * hw * a.dst_pitch + seg * a.seg_vecs * 8 + c0;
    const int y0 = blockIdx.y * kDwRows;
    const int y1 = min(y0 + kDwRows, a.H);

    for (int i = threadIdx.x; i < 2 * 144; i += kDwThreads) s_ssq[i] = 0.f;
    {
        const float* wp = a.w[prob];
        for (int i = threadIdx.x; i < 9 * nv * 8; i += kDwThreads) s_w[i] = f2act(__ldg(wp + i));
    }
    __syncthreads();
    const uint32_t wsm_addr = ptx::smem_u32(s_w) + (uint32_t)((active ? v : 0) * 16);
    auto W9 = [&](int t) -> uint4 {
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(wsm_addr + (uint32_t)(t * nv * 16)));
        return r;
    };
    float ssq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ssq[e] = 0.f;

    if (active) {
        const bool has_l = x > 0, has_r = x + 1 < a.W;
        const uint32_t slot0 = ptx::smem_u32(ring) + threadIdx.x * 16;
        const int n_in = (y1 - y0) + 2;                      // input rows y0-1 .. y1
        // request input row k (image row y0 - 1 + k) into ring slot k % kDepth; always commits one group
        auto issue = [&](int k) {
            if (k < n_in) {
                const int y = y0 - 1 + k;
                const bool row_ok = y >= 0 && y < a.H;
                const act_t* p = src + ((long long)(row_ok ? y : 0) * a.W + x) * a.src_pitch;
                const uint32_t d = slot0 + (uint32_t)((k % kDepth) * 3 * kDwThreads * 16);
                const uint32_t nl = (row_ok && has_l) ? 16u : 0u, nc = row_ok ? 16u : 0u, nr = (row_ok && has_r) ? 16u : 0u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(has_l ? p - a.src_pitch : p), "r"(nl) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d + kDwThreads * 16), "l"(p), "r"(nc) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d + 2 * kDwThreads * 16), "l"(has_r ? p + a.src_pitch : p), "r"(nr) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // take input row k out of the ring (it has landed once at most kDepth - 1 younger groups are pending) and
        // re-use its slot for row k + kDepth
        auto fetch = [&](int k, uint4* r) {
            asm volatile("cp.async.wait_group %0;" :: "n"(kDepth - 1) : "memory");
            const uint32_t d = slot0 + (uint32_t)((k % kDepth) * 3 * kDwThreads * 16);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[c].x), "=r"(r[c].y), "=r"(r[c].z), "=r"(r[c].w) : "r"(d + c * kDwThreads * 16));
            issue(k + kDepth);
        };
        uint4 win0[3], win1[3], win2[3];
        auto step = [&](const uint4* r0, const uint4* r1, uint4* r2, int y) {
            fetch(y - y0 + 2, r2);                           // row y + 1
            uint4 pa = hmul8(r0[0], W9(0));
            uint4 pb = hmul8(r2[0], W9(6));
            hfma8(pa, r0[1], W9(1)); hfma8(pb, r2[1], W9(7));
            hfma8(pa, r0[2], W9(2)); hfma8(pb, r2[2], W9(8));
            hfma8(pa, r1[0], W9(3));
            hfma8(pa, r1[1], W9(4));
            hfma8(pa, r1[2], W9(5));
            const uint4 raw = make_uint4(hadd2u(pa.x, pb.x), hadd2u(pa.y, pb.y), hadd2u(pa.z, pb.z), hadd2u(pa.w, pb.w));
            if (y >= a.stat_y0 && y < a.stat_y1) fhfma8(ssq, raw, raw);
            *reinterpret_cast<uint4*>(dst + ((long long)y * a.W + x) * a.dst_pitch) = raw;
        };
#pragma unroll
        for (int k = 0; k < kDepth; ++k) issue(k);
        fetch(0, win0);
        fetch(1, win1);
        for (int y = y0; y < y1; y += 3) {
            step(win0, win1, win2, y);
            if (y + 1 < y1) step(win1, win2, win0, y + 1);
            if (y + 2 < y1) step(win2, win0, win1, y + 2);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (seg < 2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_ssq[seg * 144 + c0 + e], ssq[e]);
        }
    }
    __syncthreads();
    const int Cp = a.seg_vecs * 8;
    for (int i = threadIdx.x; i < 2 * Cp; i += kDwThreads) {
        const int sg = i / Cp, c = i - sg * Cp;
        const float val = s_ssq[sg * 144 + c];
        if (val != 0.f) atomicAdd((sg == 0 ? a.sq[prob] : a.sk[prob]) + (long long)b * Cp + c, val);
    }
}
#endif

#ifndef CIDNET_ACT_BF16
// ------------------------------------------------------------------------------------------------
// dw3x3 v2 (fp16 build, CIDNET_DW_VARIANT=3): the IEL gate's data path applied to the q|k|v depthwise conv.
// A producer warp streams {16 channels, 34 columns, 4 rows} TMA boxes (SWIZZLE_32B, zero fill = the conv's
// padding) of two 16-channel groups into a shared-memory ring; four compute warps (group x 8-channel vector,
// lane = image column) read their column and both neighbours with conflict-free LDS.128, keep the 9 weight
// vectors in registers and evaluate the 3x3 in scatter form (three running sums).  Compared with the
// register-window kernel above nothing waits on a global load: the ring keeps kDw2Stages x 4 rows in flight.
// ------------------------------------------------------------------------------------------------
static constexpr int kDw2Cols = 32;          // output columns per warp (all 32 lanes produce one)
static constexpr int kDw2BoxCols = 40;       // 34 needed; 40 keeps the row pitch (1280 B) a multiple of the swizzle period
static constexpr int kDw2RB = 4;             // rows per box
static constexpr int kDw2Stages = 4;
static constexpr int kDw2Rows = 32;          // output rows per CTA
static constexpr uint32_t kDw2GroupBytes = kDw2RB * kDw2BoxCols * 32;   // one 16-channel group box
static constexpr uint32_t kDw2StageBytes = 2 * kDw2GroupBytes;
static constexpr int kDw2Threads = 160;

struct Dw2Args {
    CUtensorMap tm[2][3];       // per problem, per segment: {C channels, W, H, B}, box {16, 40, 4, 1}, SWIZZLE_32B
    Dw3Args g;
    int groups_per_seg;         // 16-channel groups per segment = ceil(Cp / 16)
};

template <int kMinBlocks>
__global__ void __launch_bounds__(kDw2Threads, kMinBlocks)
dw3x3_v2_kernel(const __grid_constant__ Dw2Args A) {
    const Dw3Args& a = A.g;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + kDw2Stages * kDw2StageBytes);
    uint64_t* empty = full + kDw2Stages;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Cp = a.seg_vecs * 8;
    const int gps = A.groups_per_seg, ngroups = 3 * gps, npairs = (ngroups + 1) / 2;
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int pair = blockIdx.x % npairs, strip = blockIdx.x / npairs;
    const int X0 = strip * kDw2Cols;                         // image column of lane 0
    const int y0 = blockIdx.y * kDw2Rows;
    const int y1 = min(y0 + kDw2Rows, a.H);
    const int nrows = (y1 - y0) + 2;                         // input rows y0-1 .. y1
    const int nblocks = (nrows + kDw2RB - 1) / kDw2RB;

    if (tid == 0) {
        for (int s = 0; s < kDw2Stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 4); }
        ptx::fence_barrier_init();
    }
    __syncthreads();

    if (warp == 4) {
        if (lane == 0) {
            int gseg[2], gc0[2];
            for (int g = 0; g < 2; ++g) {
                const int grp = 2 * pair + g;
                gseg[g] = grp < ngroups ? grp / gps : -1;
                gc0[g] = grp < ngroups ? (grp - gseg[g] * gps) * 16 : 0;
            }
            const uint32_t bytes = (gseg[0] >= 0 ? kDw2GroupBytes : 0u) + (gseg[1] >= 0 ? kDw2GroupBytes : 0u);
            for (int k = 0; k < nblocks; ++k) {
                const int s = k % kDw2Stages;
                ptx::mbar_wait(&empty[s], ((k / kDw2Stages) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&full[s], bytes);
                uint8_t* dst = ring + (size_t)s * kDw2StageBytes;
                const int yb = y0 - 1 + k * kDw2RB;
                for (int g = 0; g < 2; ++g)
                    if (gseg[g] >= 0)
                        ptx::tma_load_4d(dst + g * kDw2GroupBytes, &A.tm[prob][gseg[g]], &full[s], gc0[g], X0 - 1, yb, b);
            }
        }
        return;
    }
    const int g = warp >> 1, vec = warp & 1;
    const int grp = 2 * pair + g;
    const int seg = grp < ngroups ? grp / gps : 0;
    const int c0 = (grp - seg * gps) * 16 + vec * 8;           // channel within the segment
    const bool ch_live = grp < ngroups && c0 < Cp;             // whole-vector granularity (Cp is a multiple of 8)
    const int x = X0 + lane;
    const bool col_in = x < a.W;
    const long long hw = (long long)a.H * a.W;
    act_t* dst = a.dst[prob] + (long long)b * hw * a.dst_pitch + seg * Cp + c0;

    uint4 w[9];
    {
        const float* wp = a.w[prob] + seg * Cp + c0;             // [9][3*Cp] tap major
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0;
            if (ch_live) {
                f0 = __ldg(reinterpret_cast<const float4*>(wp + t * 3 * Cp));
                f1 = __ldg(reinterpret_cast<const float4*>(wp + t * 3 * Cp) + 1);
            }
            __half2 h;
            h = __floats2half2_rn(f0.x, f0.y); w[t].x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f0.z, f0.w); w[t].y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f1.x, f1.y); w[t].z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(f1.z, f1.w); w[t].w = *reinterpret_cast<uint32_t*>(&h);
        }
    }
    auto swz = [](uint32_t o) { return o ^ (((o >> 7) & 1u) << 4); };
    const uint32_t off_l = swz((uint32_t)lane * 32u + (uint32_t)vec * 16u);
    const uint32_t off_c = swz((uint32_t)(lane + 1) * 32u + (uint32_t)vec * 16u);
    const uint32_t off_r = swz((uint32_t)(lane + 2) * 32u + (uint32_t)vec * 16u);
    float ssq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ssq[e] = 0.f;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);

    // iteration j (= arriving input row y0-1+j): finishes output row y0-2+j.  pn / pm / pf: running sums
    auto iter = [&](int j, uint4& pn, uint4& pm, uint4& pf) {
        const int k = j / kDw2RB, rr = j - k * kDw2RB;
        const int s = k % kDw2Stages;
        if (rr == 0) ptx::mbar_wait(&full[s], (k / kDw2Stages) & 1u);
        const uint8_t* base = ring + (size_t)s * kDw2StageBytes + g * kDw2GroupBytes + rr * (kDw2BoxCols * 32);
        const uint4 tl = *reinterpret_cast<const uint4*>(base + off_l);
        const uint4 tc = *reinterpret_cast<const uint4*>(base + off_c);
        const uint4 tr = *reinterpret_cast<const uint4*>(base + off_r);
        if (rr == kDw2RB - 1 || j == nrows - 1) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[s]);
        }
        pn = hmul8(tl, w[0]); hfma8(pm, tl, w[3]); hfma8(pf, tl, w[6]);
        hfma8(pn, tc, w[1]);  hfma8(pm, tc, w[4]); hfma8(pf, tc, w[7]);
        hfma8(pn, tr, w[2]);  hfma8(pm, tr, w[5]); hfma8(pf, tr, w[8]);
        const int yo = y0 - 2 + j;
        if (ch_live && col_in && yo >= y0 && yo < y1) {
            if (seg < 2 && yo >= a.stat_y0 && yo < a.stat_y1) fhfma8(ssq, pf, pf);     // squares of the STORED values
            *reinterpret_cast<uint4*>(dst + ((long long)yo * a.W + x) * a.dst_pitch) = pf;
        }
    };
    uint4 pA = zero4, pB = zero4, pC = zero4;
    for (int j = 0; j < nrows; j += 3) {
        iter(j, pA, pB, pC);
        if (j + 1 < nrows) iter(j + 1, pC, pA, pB);
        if (j + 2 < nrows) iter(j + 2, pB, pC, pA);
    }
    if (seg < 2 && ch_live) {
        // sum over the warp's 32 columns, one global atomic per channel per warp
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = ssq[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v != 0.f) atomicAdd((seg == 0 ? a.sq[prob] : a.sk[prob]) + (long long)b * Cp + c0 + e, v);
        }
    }
}

int encode_map_generic_swz(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, int swizzle_bytes);   // conv_gemm.cu

static int launch_dw3_v2(const Dw3Args& a, cudaStream_t stream, int min_blocks) {
    Dw2Args A;
    memset(&A, 0, sizeof A);
    A.g = a;
    const int Cp = a.seg_vecs * 8;
    A.groups_per_seg = ceil_div(Cp, 16);
    const long long hw = (long long)a.H * a.W;
    for (int p = 0; p < a.nprob; ++p)
        for (int sg = 0; sg < 3; ++sg) {
            const uint64_t pb = (uint64_t)a.src_pitch * sizeof(act_t);
            const uint64_t dims[4] = {(uint64_t)Cp, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
            const uint64_t str[3] = {pb, pb * a.W, pb * hw};
            const uint32_t box[4] = {16, (uint32_t)kDw2BoxCols, (uint32_t)kDw2RB, 1};
            int rc = encode_map_generic_swz(&A.tm[p][sg], a.src[p][sg], 4, dims, str, box, 32);
            if (rc) return rc;
        }
    const size_t smem = 1024 + (size_t)kDw2Stages * kDw2StageBytes + 2 * kDw2Stages * sizeof(uint64_t) + 64;
    static bool configured = false;
    if (!configured) {
        CIDNET_CUDA_OK(cudaFuncSetAttribute(dw3x3_v2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CIDNET_CUDA_OK(cudaFuncSetAttribute(dw3x3_v2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int npairs = (3 * A.groups_per_seg + 1) / 2;
    dim3 grid(ceil_div(a.W, kDw2Cols) * npairs, ceil_div(a.H, kDw2Rows), a.B * a.nprob);
    if (min_blocks == 4) dw3x3_v2_kernel<4><<<grid, kDw2Threads, smem, stream>>>(A);
    else                 dw3x3_v2_kernel<3><<<grid, kDw2Threads, smem, stream>>>(A);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}
#endif

int launch_dw3(const Dw3Args& a_in, cudaStream_t stream) {
    Dw3Args a = a_in;
    if (a.stat_y1 <= 0) { a.stat_y0 = 0; a.stat_y1 = a.H; }
    CIDNET_CHECK(a.seg_vecs * 8 <= 144 && a.nv == 3 * a.seg_vecs, CIDNET_ERR_INVALID, "dw3: bad channel layout");
    dim3 grid(ceil_div(a.W * a.nv, kDwThreads), ceil_div(a.H, kDwRows), a.B * a.nprob);
#ifndef CIDNET_ACT_BF16
    // 8 (default): per-thread cp.async ring, 4 input rows in flight, weights re-read from smem, 5 CTAs / SM (L1 launch at cfg 2:
    //    181 -> 147 us; 16x400x600: 712 -> 640 us); 9: the same with 6 rows in flight, 4 CTAs / SM;
    // 7: register loads one row ahead, weights re-read from smem, 4 CTAs / SM;
    // 0: weights in registers, no prefetch, 4 CTAs / SM; 1: weights in registers, one-row prefetch, 3 CTAs / SM;
    // 2: fp32-accumulate FHFMA kernel
    static const int variant = getenv("CIDNET_DW_VARIANT") ? atoi(getenv("CIDNET_DW_VARIANT")) : 8;
    if (variant == 3 || variant == 4) return launch_dw3_v2(a, stream, variant == 3 ? 4 : 3);   // v2 (TMA ring), 4 / 3 CTAs per SM
    if (variant == 8 || variant == 9) {               // per-thread cp.async ring, 4 rows (8) / 6 rows (9) in flight
        const int depth = variant == 8 ? 4 : 6;
        const size_t smem = (size_t)depth * 3 * kDwThreads * 16 + (size_t)9 * a.nv * 8 * sizeof(act_t);
        static bool configured = false;
        if (!configured) {
            CIDNET_CUDA_OK(cudaFuncSetAttribute(dw3x3_cpasync_kernel<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            CIDNET_CUDA_OK(cudaFuncSetAttribute(dw3x3_cpasync_kernel<6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            configured = true;
        }
        if (variant == 8) dw3x3_cpasync_kernel<4, 5><<<grid, kDwThreads, smem, stream>>>(a);
        else              dw3x3_cpasync_kernel<6, 4><<<grid, kDwThreads, smem, stream>>>(a);
        CIDNET_CUDA_OK(cudaGetLastError());
        return CIDNET_OK;
    }
    // (5 CTAs / SM with register loads needs spills and measured slower: profiles/r01_dw3x3_variants.txt, variants 5 / 6)
    if (variant == 7)      dw3x3_kernel<true, 4, true><<<grid, kDwThreads, 0, stream>>>(a);    // weights in smem, prefetch, 4 CTAs / SM
    else if (variant == 0) dw3x3_kernel<false, 4><<<grid, kDwThreads, 0, stream>>>(a);
    else if (variant == 1) dw3x3_kernel<true, 3><<<grid, kDwThreads, 0, stream>>>(a);
    else                   dw3x3_f32acc_kernel<<<grid, kDwThreads, 0, stream>>>(a);
#else
    dw3x3_f32acc_kernel<<<grid, kDwThreads, 0, stream>>>(a);
#endif
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

// ------------------------------------------------------------------ Gram ----
struct GramArgs {
    CUtensorMap tmQ[2];      // per problem: q  [B][HW][pitch] viewed {C, HW, B}
    CUtensorMap tmK[2];
    float* gram[2];          // [B][heads][18][18]
    int C, heads, hw, chunks_per_cta, nchunks;
};

static constexpr int kGramThreads = 192;
static constexpr uint32_t kBlk = 64 * 128;     // one [64 pixels][64 channels] swizzled block = 8 KB
static constexpr int kGramStages = 3;

// MN-major SWIZZLE_128B operand: 64 MN-elements (128 B) contiguous per K row, 8 K rows per
// 1024-byte group (SBO), next 64 MN-elements LBO bytes further.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kGramThreads, 1)
gram_kernel(const __grid_constant__ GramArgs a) {
    // 1024-byte alignment is required by the SWIZZLE_128B TMA / UMMA tiles; using the array directly
    // (no integer round trip) keeps the accesses in the shared state space (LDS / STS, not generic LD / ST)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    const int C = a.C;
    const int mtiles_ = (C + 127) / 128;
    const int nA = 2 * mtiles_;                   // 64-channel blocks of q (blocks beyond C are TMA zero fill)
    const int N = (C + 15) / 16 * 16;             // MMA N (k channels)
    const int nB = (N + 63) / 64;
    const int mtiles = (C + 127) / 128;
    const uint32_t stage_bytes = (uint32_t)(nA + nB) * kBlk;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kGramStages * stage_bytes);
    uint64_t* empty = full + kGramStages;
    uint64_t* done = empty + kGramStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.y, b = blockIdx.z;
    const int c_begin = blockIdx.x * a.chunks_per_cta;
    const int c_end = min(c_begin + a.chunks_per_cta, a.nchunks);
    const int niter = c_end - c_begin;
    if (niter <= 0) return;

    uint32_t ncols = 32;
    while (ncols < (uint32_t)(mtiles * N)) ncols <<= 1;

    // q blocks that lie entirely beyond C (C = 36 / 72: the upper 64 of the M = 128 rows) are never loaded: every
    // TMA row request costs the same whether it moves data or zero fill, so they are zeroed once here instead
    const int nA_load = min(nA, (C + 63) / 64);
    if (nA_load < nA) {
        for (int s = 0; s < kGramStages; ++s) {
            uint4* z = reinterpret_cast<uint4*>(smem + (size_t)s * stage_bytes + (size_t)nA_load * kBlk);
            for (int i = threadIdx.x; i < (int)((nA - nA_load) * kBlk / 16); i += kGramThreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy zeros -> visible to the tensor core
    }
    if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&a.tmQ[prob]); ptx::prefetch_tensormap(&a.tmK[prob]); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kGramStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
            ptx::mbar_init(done, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, ncols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < niter; ++i) {
                const int s = i % kGramStages;
                const uint32_t ph = (uint32_t)(i / kGramStages) & 1u;
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&full[s], (uint32_t)(nA_load + nB) * kBlk);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                const int p0 = (c_begin + i) * 64;
                for (int j = 0; j < nA_load; ++j) ptx::tma_load_3d(st + j * kBlk, &a.tmQ[prob], &full[s], j * 64, p0, b);
                for (int j = 0; j < nB; ++j) ptx::tma_load_3d(st + (nA + j) * kBlk, &a.tmK[prob], &full[s], j * 64, p0, b);
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: fp32 accumulate, A and B MN-major (bits 15, 16), M = 128
        const uint32_t idesc = ptx::umma_idesc_f16(CIDNET_UMMA_FMT, (uint32_t)N) | (1u << 15) | (1u << 16);
        // the whole warp runs the loop convergently, one elected lane's tcgen05.mma / commit take effect
        // (see conv_gemm.cu: a divergent `if (lane == 0)` costs an ELECT retry loop per instruction)
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;
        for (int i = 0; i < niter; ++i) {
            const int s = i % kGramStages;
            const uint32_t ph = (uint32_t)(i / kGramStages) & 1u;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            const uint32_t sa = ptx::smem_u32(smem + (size_t)s * stage_bytes);
            const uint64_t dB0 = umma_desc_mn_sw128(sa + nA * kBlk, kBlk);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                if (mt < mtiles) {
                    const uint64_t dA0 = umma_desc_mn_sw128(sa + 2 * mt * kBlk, kBlk);
                    const uint32_t dt = tmem_base + mt * N;
                    if (i == 0) ptx::umma_f16_lead<false>(leader, dt, dA0, dB0, idesc);
                    else        ptx::umma_f16_lead<true>(leader, dt, dA0, dB0, idesc);
                    ptx::umma_f16_lead<true>(leader, dt, dA0 + (2048 >> 4), dB0 + (2048 >> 4), idesc);
                    ptx::umma_f16_lead<true>(leader, dt, dA0 + (4096 >> 4), dB0 + (4096 >> 4), idesc);
                    ptx::umma_f16_lead<true>(leader, dt, dA0 + (6144 >> 4), dB0 + (6144 >> 4), idesc);
                }
            }
            ptx::umma_commit_lead(leader, &empty[s]);
            if (i == niter - 1) ptx::umma_commit_lead(leader, done);
        }
    } else {
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        ptx::mbar_wait(done, 0);
        ptx::tc_fence_after();
        float* gdst = a.gram[prob] + (long long)b * a.heads * 324;
        // Only the per-head 18x18 diagonal blocks of the C x C product are wanted.  The 32 q channels of a warp
        // belong to at most 3 heads, so the warp reads just the k-column window of those heads (<= 5 x 16 columns
        // instead of all N) and every thread adds its own 18 columns with 8-byte vector reductions (row bases are
        // multiples of 18 floats = 72 B, pairs start on even columns).  ncu on the previous all-columns / scalar-RED
        // epilogue: ~70 % of the kernel's stall samples (index arithmetic per element, RED drain at EXIT).
        for (int mt = 0; mt < mtiles; ++mt) {
            const int ch_lo = mt * 128 + q4 * 32;                 // warp-uniform
            if (ch_lo >= C) break;
            const int ch_hi = min(ch_lo + 31, C - 1);
            const int col_lo = ((ch_lo / 18) * 18) & ~15;
            const int col_hi = min((ch_hi / 18) * 18 + 18, N);
            const int ch = mt * 128 + r;
            const int head = ch / 18;
            const int kc0 = head * 18;                            // this thread's first k column
            float* rowp = gdst + (long long)ch * 18;              // (head * 18 + qi) * 18 == ch * 18
            const bool row_ok = ch < C;
            for (int cc = col_lo; cc < col_hi; cc += 16) {
                float v[16];
                ptx::tmem_ld16(taddr + mt * N + cc, v);
                const int rel = cc - kc0;                         // column of v[0] relative to the thread's block
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    if (row_ok && (unsigned)(rel + j) < 18u)
                        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(rowp + rel + j), "f"(v[j]), "f"(v[j + 1]) : "memory");
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, ncols);
}

int encode_map_generic(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box);   // conv_gemm.cu

int launch_gram(const GramLaunch& L, cudaStream_t stream) {
    CIDNET_CHECK(L.C == 36 || L.C == 72 || L.C == 144, CIDNET_ERR_INVALID, "gram: C must be 36/72/144");
    GramArgs a;
    memset(&a, 0, sizeof a);
    a.C = L.C; a.heads = L.heads; a.hw = L.H * L.W;
    a.nchunks = ceil_div(a.hw, 64);
    for (int p = 0; p < L.nprob; ++p) {
        const uint64_t dims[3] = {(uint64_t)L.C, (uint64_t)a.hw, (uint64_t)L.B};
        const uint64_t img_px = L.img_stride_px > 0 ? (uint64_t)L.img_stride_px : (uint64_t)a.hw;
        const uint64_t str[2] = {(uint64_t)L.pitch * sizeof(act_t), (uint64_t)L.pitch * sizeof(act_t) * img_px};
        const uint32_t box[3] = {64, 64, 1};
        int rc = encode_map_generic(&a.tmQ[p], L.q[p], 3, dims, str, box);
        if (rc) return rc;
        if ((rc = encode_map_generic(&a.tmK[p], L.k[p], 3, dims, str, box))) return rc;
        a.gram[p] = L.gram[p];
    }
    const int nA = 2 * ceil_div(L.C, 128), nB = ceil_div(round_up(L.C, 16), 64);
    const size_t smem = 1024 + (size_t)kGramStages * (nA + nB) * kBlk + 64;
    // split-K over CTAs, sized for ONE wave: shared memory allows 3 / 2 / 1 CTAs per SM for C = 36 / 72 / 144 (and
    // C = 144 needs all 512 TMEM columns).  A partial second wave doubled the kernel time at the coarse levels, and
    // every extra CTA costs C x 18 more atomic adds.
    int sms = 148;
    { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const int per_sm = L.C >= 144 ? 1 : (int)std::min<size_t>(3, (227 * 1024) / smem);
    int nsplit = (sms * per_sm) / (L.nprob * L.B);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > a.nchunks) nsplit = a.nchunks;
    a.chunks_per_cta = ceil_div(a.nchunks, nsplit);
    nsplit = ceil_div(a.nchunks, a.chunks_per_cta);
    static bool configured = false;
    if (!configured) {
        CIDNET_CUDA_OK(cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    dim3 grid(nsplit, L.nprob, L.B);
    gram_kernel<<<grid, kGramThreads, smem, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

// ---------------------------------------------------------------- fold ------
__global__ void __launch_bounds__(256)
cab_fold_kernel(const CabFoldArgs a) {
    __shared__ float s_attn[144 * 18];
    const int prob = blockIdx.x, b = blockIdx.y;
    const int C = a.C, tid = threadIdx.x;
    const float* G = (prob ? a.gram[1] : a.gram[0]) + (long long)b * a.heads * 324;
    const float* sq = (prob ? a.sq[1] : a.sq[0]) + (long long)b * a.Cp;
    const float* sk = (prob ? a.sk[1] : a.sk[0]) + (long long)b * a.Cp;
    const float* wo = prob ? a.wo[1] : a.wo[0];
    // each CTA of the z dimension writes a slice of the rows (the softmax is recomputed by every slice: cheap)
    const int rows_per = (a.n_rows + gridDim.z - 1) / gridDim.z;
    const int i_begin = blockIdx.z * rows_per * a.kt;
    const int i_end = min((int)(blockIdx.z + 1) * rows_per, a.n_rows) * a.kt;
    // The kernel is a chain of dependent latencies (Gram loads -> softmax -> barrier -> W_o loads -> store), not work
    // (ncu: ~9 us per CTA at every level).  So: the W_o segments of this thread's (<= 2) output elements are requested
    // FIRST and arrive while the softmax runs, and the softmax uses the fast reciprocal / rsqrt / exp2 paths (the
    // attention weights keep ~1e-6 relative accuracy; contract 2e-3 on the image).
    float wv[2][18];
    int wj[2], wh[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = i_begin + tid + u * 256;
        const int o = i / a.kt, kin = i - o * a.kt;
        const bool nz = i < i_end && o < C && kin < C;
        wh[u] = nz ? kin / 18 : -1;
        wj[u] = nz ? kin - wh[u] * 18 : 0;
#pragma unroll
        for (int c = 0; c < 18; ++c) wv[u][c] = nz ? __ldg(wo + o * C + wh[u] * 18 + c) : 0.f;
    }
    if (tid < C) {
        const int head = tid / 18;
        const float inv_nq = rsqrtf(fmaxf(__ldg(sq + tid), 1e-24f));      // 1 / max(sqrt(sum q^2), 1e-12)
        const float temp = (prob ? a.temp[1] : a.temp[0])[head];
        float logit[18], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 18; ++j) {
            const float inv_nk = rsqrtf(fmaxf(__ldg(sk + head * 18 + j), 1e-24f));
            logit[j] = __ldcg(G + tid * 18 + j) * (inv_nq * inv_nk) * temp;
            mx = fmaxf(mx, logit[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 18; ++j) { logit[j] = __expf(logit[j] - mx); sum += logit[j]; }
        const float inv = __fdividef(1.0f, sum);
#pragma unroll
        for (int j = 0; j < 18; ++j) s_attn[tid * 18 + j] = logit[j] * inv;
    }
    __syncthreads();
    // M[o][kin] = sum_{c' in head(kin)} Wo[o][head*18 + c'] * attn[head*18 + c'][kin - head*18]
    act_t* m = (prob ? a.m_out[1] : a.m_out[0]) + (long long)b * a.n_rows * a.kt;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int i = i_begin + tid + u * 256;
        if (i >= i_end) break;
        float acc = 0.f;
        if (wh[u] >= 0) {
#pragma unroll
            for (int c = 0; c < 18; ++c) acc = fmaf(wv[u][c], s_attn[(wh[u] * 18 + c) * 18 + wj[u]], acc);
        }
        m[i] = f2act(acc);
    }
    for (int i = i_begin + tid + 512; i < i_end; i += 256) {      // slices larger than 512 elements (not used by launch_cab_fold)
        const int o = i / a.kt, kin = i - o * a.kt;
        float acc = 0.f;
        if (o < C && kin < C) {
            const int head = kin / 18, j = kin - head * 18;
#pragma unroll
            for (int c = 0; c < 18; ++c) acc = fmaf(wo[o * C + head * 18 + c], s_attn[(head * 18 + c) * 18 + j], acc);
        }
        m[i] = f2act(acc);
    }
}

int launch_cab_fold(const CabFoldArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.C <= 144, CIDNET_ERR_INVALID, "fold: C too large");
    // row slices per (problem, image): <= 2 output elements per thread (the kernel is a chain of dependent L2
    // latencies, not work: ncu 13 us at C = 144 with 16 slices of ~7 elements per thread)
    // rows_per * kt <= 512 elements per slice where possible
    const int slices = std::min(a.n_rows, std::max(16, ceil_div(a.n_rows, std::max(1, 512 / a.kt))));
    dim3 grid(a.nprob, a.B, slices);
    cab_fold_kernel<<<grid, 256, 0, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
