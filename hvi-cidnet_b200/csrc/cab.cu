// CAB (channel cross-attention, net/LCA.py:19-41) in three kernels:
//
//   dw3x3 kernel    : depthwise 3x3 (zero pad) of q_pre | k_pre | v_pre, NHWC in; [q | k] and v go to two
//                     separate NHWC buffers (the Gram's TMA boxes then never drag v's bytes along).
//                     Sliding window down the rows: one thread owns (column x, 8 channels).
//   gram_kernel     : G[Cq, Ck] = sum_pixels q k^T on the tensor cores.  The NHWC tiles
//                     [64 pixels][64 channels] land by TMA (SWIZZLE_128B) exactly in the UMMA
//                     *MN-major* canonical layout, so the contraction over pixels needs no
//                     transposition: tcgen05.mma kind::f16 with a_major = b_major = MN, fp32
//                     accumulation in TMEM over the CTA's whole pixel range (split-K across CTAs).
//                     While the tiles are resident the four epilogue warps also accumulate sum q^2 and
//                     sum k^2 per channel (the F.normalize denominators).  Every CTA writes its partial
//                     [per-head 18x18 blocks | sum q^2 | sum k^2] to its own slab entry with plain stores:
//                     no atomics, no pre-zeroed buffers, and the later summation order is fixed ->
//                     the attention is bit-reproducible run to run.
//   cab_fold_kernel : one CTA per (problem, head, image): sums the slab entries in a FIXED order,
//                     L2-normalise (eps 1e-12), temperature, softmax, and fold attn into
//                     project_out:  M_b = W_o * blockdiag(attn_b)   (one CxC matrix per image, written
//                     in the packed-weight layout of the conv GEMM), so project_out(attn @ v)
//                     becomes a per-image 1x1 conv on v (identity verified to 6e-7).
//   cab_reduce_kernel (row-strip sharding only): slab -> one raw [Gram | sum q^2 | sum k^2] vector per
//                     rank, which the host all-reduces before the fold.
#include "cab.cuh"
#include "ptx_sm100.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace cidnet {

// ------------------------------------------------------------- depthwise ----
static constexpr int kDwThreads = 128;

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

// destination of a thread's 16-byte vector: segments 0 / 1 (q, k) -> the [q | k] buffer, segment 2 (v) -> the v buffer
__device__ __forceinline__ act_t* dw_dst(const Dw3Args& a, int prob, int b, int seg, int c0, long long hw, int* pitch) {
    const int Cp = a.seg_vecs * 8;
    if (seg < 2) { *pitch = 2 * Cp; return a.dst_qk[prob] + (long long)b * hw * (2 * Cp) + seg * Cp + c0; }
    *pitch = Cp;
    return a.dst_v[prob] + (long long)b * hw * Cp + c0;
}

#ifdef CIDNET_ACT_BF16
// bf16 build: one thread = (column x, 8 channels): 9 packed weight vectors + a raw 3x3 window in registers,
// FHFMA (16-bit x 16-bit + fp32) so no conversion instructions are needed; fp32 accumulation
__global__ void __launch_bounds__(kDwThreads, 3)
dw3x3_f32acc_kernel(const Dw3Args a) {
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int nv = a.nv;                                   // 16-byte vectors per pixel (all segments)
    const int idx = blockIdx.x * kDwThreads + threadIdx.x; // vector index along the row
    const int x = idx / nv, v = idx - x * nv;
    ptx::pdl_wait();
    ptx::pdl_trigger();
    if (x >= a.W) return;
    const int seg = v / a.seg_vecs;                        // 0 = q, 1 = k, 2 = v
    const int c0 = (v - seg * a.seg_vecs) * 8;             // channel within the segment
    const long long hw = (long long)a.H * a.W;
    const act_t* src = a.src[prob][seg] + (long long)b * hw * a.src_pitch + c0;
    int dpitch;
    act_t* dst = dw_dst(a, prob, b, seg, c0, hw, &dpitch);
    const int y0 = blockIdx.y * a.rows_per_cta;
    const int y1 = min(y0 + a.rows_per_cta, a.H);

    uint4 w[9];
    {
        const float* wp = a.w[prob] + seg * a.seg_vecs * 8 + c0;   // [9][nv*8] tap major
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            act_t* h = reinterpret_cast<act_t*>(&w[t]);
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = f2act(__ldg(wp + t * nv * 8 + e));
        }
    }
    const bool has_l = x > 0, has_r = x + 1 < a.W;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    uint4 win0[3], win1[3], win2[3];
    auto load_row = [&](int y, uint4* r) {
        r[0] = r[1] = r[2] = zero4;
        if (y < 0 || y >= a.H) return;
        const act_t* p = src + ((long long)y * a.W + x) * a.src_pitch;
        r[1] = __ldcg(reinterpret_cast<const uint4*>(p));
        if (has_l) r[0] = __ldcg(reinterpret_cast<const uint4*>(p - a.src_pitch));
        if (has_r) r[2] = __ldcg(reinterpret_cast<const uint4*>(p + a.src_pitch));
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
        *reinterpret_cast<uint4*>(dst + ((long long)y * a.W + x) * dpitch) = raw;
    };
    load_row(y0 - 1, win0);
    load_row(y0, win1);
    for (int y = y0; y < y1; y += 3) {          // window roles rotate, no register moves
        step(win0, win1, win2, y);
        if (y + 1 < y1) step(win1, win2, win0, y + 1);
        if (y + 2 < y1) step(win2, win0, win1, y + 2);
    }
}
#else
// fp16 build: PACKED fp16 math (HFMA2: two multiply-adds per issue slot, the accumulators are the packed output
// vector -> no conversions).  The 9-tap sum is rounded to fp16 after every tap (two chains); end-to-end effect
// measured against the fp32 oracle: max-abs 1.0e-4 -> 1.2e-4.
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

// Per-thread cp.async ring: the 16-byte loads of every input row are issued kDepth rows ahead with cp.async (zero fill
// outside the image = the conv's padding) into the thread's own shared-memory slots -- no registers held by loads in
// flight, no barriers (a thread only reads what it requested itself); the 9 weight vectors are re-read from shared memory
// every row.
// Two pixels per thread (x0 = 2 * pair, x0 + 1): ncu on the one-pixel-per-thread kernel of round 1 / early round 2 showed the
// LSU / shared-memory data pipe as the limiter (l1tex data-pipe wavefronts 74 %: per row and thread 3 cp.async + 3 ring LDS.128 + 9 weight LDS.128 for ONE output
// vector).  With two adjacent pixels the four columns x0-1 .. x0+2 serve both outputs and every weight vector is loaded
// once for two uses: 4 + 4 + 9 = 17 LSU operations per TWO outputs instead of 30.  Measured (cfg 2, event-timed): L1 70.8 -> 55.5 us,
// L2 32.0 -> 25.8 us, L3 23.7 -> 21.2 us per launch; the step 1.620 -> 1.573 ms.  Same tap order and rounding points: bit-identical output.
template <int kDepth, int kMinBlocks>
__global__ void __launch_bounds__(kDwThreads, kMinBlocks)
dw3x3_cpasync2_kernel(const Dw3Args a) {
    extern __shared__ __align__(16) uint8_t dw_smem[];
    uint4* ring = reinterpret_cast<uint4*>(dw_smem);                       // [kDepth][4][kDwThreads]
    act_t* s_w = reinterpret_cast<act_t*>(ring + kDepth * 4 * kDwThreads);  // [9][nv * 8]
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int nv = a.nv;
    const int idx = blockIdx.x * kDwThreads + threadIdx.x;
    const int xp = idx / nv, v = idx - xp * nv;
    const int x = 2 * xp;
    const bool active = x < a.W;
    const int seg = active ? v / a.seg_vecs : 0;
    const int c0 = (v - seg * a.seg_vecs) * 8;
    const long long hw = (long long)a.H * a.W;
    const act_t* src = a.src[prob][seg] + (long long)b * hw * a.src_pitch + c0;
    int dpitch;
    act_t* dst = dw_dst(a, prob, b, seg, c0, hw, &dpitch);
    const int y0 = blockIdx.y * a.rows_per_cta;
    const int y1 = min(y0 + a.rows_per_cta, a.H);
    {
        const float* wp = a.w[prob];
        for (int i = threadIdx.x; i < 9 * nv * 8; i += kDwThreads) s_w[i] = f2act(__ldg(wp + i));
    }
    __syncthreads();
    ptx::pdl_wait();          // the weights above are constants; everything below reads the previous kernel's output
    ptx::pdl_trigger();
    if (!active) return;
    const uint32_t wsm_addr = ptx::smem_u32(s_w) + (uint32_t)(v * 16);
    auto W9 = [&](int t) -> uint4 {
        uint4 r;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(wsm_addr + (uint32_t)(t * nv * 16)));
        return r;
    };
    const bool has0 = x > 0, has2 = x + 1 < a.W, has3 = x + 2 < a.W;      // columns x-1, (x always), x+1, x+2 inside the image
    const uint32_t slot0 = ptx::smem_u32(ring) + threadIdx.x * 16;
    const int n_in = (y1 - y0) + 2;
    auto issue = [&](int k) {
        if (k < n_in) {
            const int y = y0 - 1 + k;
            const bool row_ok = y >= 0 && y < a.H;
            const act_t* p = src + ((long long)(row_ok ? y : 0) * a.W + x) * a.src_pitch;
            const uint32_t d = slot0 + (uint32_t)((k % kDepth) * 4 * kDwThreads * 16);
            const uint32_t n0 = (row_ok && has0) ? 16u : 0u, n1 = row_ok ? 16u : 0u, n2 = (row_ok && has2) ? 16u : 0u, n3 = (row_ok && has3) ? 16u : 0u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(has0 ? p - a.src_pitch : p), "r"(n0) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d + kDwThreads * 16), "l"(p), "r"(n1) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d + 2 * kDwThreads * 16), "l"(has2 ? p + a.src_pitch : p), "r"(n2) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d + 3 * kDwThreads * 16), "l"(has3 ? p + 2 * a.src_pitch : p), "r"(n3) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto fetch = [&](int k, uint4* r) {
        asm volatile("cp.async.wait_group %0;" :: "n"(kDepth - 1) : "memory");
        const uint32_t d = slot0 + (uint32_t)((k % kDepth) * 4 * kDwThreads * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r[c].x), "=r"(r[c].y), "=r"(r[c].z), "=r"(r[c].w) : "r"(d + c * kDwThreads * 16));
        issue(k + kDepth);
    };
    uint4 win0[4], win1[4], win2[4];
    auto step = [&](const uint4* r0, const uint4* r1, uint4* r2, int y) {
        fetch(y - y0 + 2, r2);                           // row y + 1
        // same tap order and rounding points as the one-pixel kernel (chain a: rows 0-1, chain b: row 2)
        uint4 w = W9(0);
        uint4 pa0 = hmul8(r0[0], w), pa1 = hmul8(r0[1], w);
        w = W9(6);
        uint4 pb0 = hmul8(r2[0], w), pb1 = hmul8(r2[1], w);
        w = W9(1); hfma8(pa0, r0[1], w); hfma8(pa1, r0[2], w);
        w = W9(7); hfma8(pb0, r2[1], w); hfma8(pb1, r2[2], w);
        w = W9(2); hfma8(pa0, r0[2], w); hfma8(pa1, r0[3], w);
        w = W9(8); hfma8(pb0, r2[2], w); hfma8(pb1, r2[3], w);
        w = W9(3); hfma8(pa0, r1[0], w); hfma8(pa1, r1[1], w);
        w = W9(4); hfma8(pa0, r1[1], w); hfma8(pa1, r1[2], w);
        w = W9(5); hfma8(pa0, r1[2], w); hfma8(pa1, r1[3], w);
        act_t* o = dst + ((long long)y * a.W + x) * dpitch;
        *reinterpret_cast<uint4*>(o) = make_uint4(hadd2u(pa0.x, pb0.x), hadd2u(pa0.y, pb0.y), hadd2u(pa0.z, pb0.z), hadd2u(pa0.w, pb0.w));
        if (has2)
            *reinterpret_cast<uint4*>(o + dpitch) = make_uint4(hadd2u(pa1.x, pb1.x), hadd2u(pa1.y, pb1.y), hadd2u(pa1.z, pb1.z), hadd2u(pa1.w, pb1.w));
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
}
#endif

int launch_dw3(const Dw3Args& a_in, cudaStream_t stream) {
    Dw3Args a = a_in;
    CIDNET_CHECK(a.seg_vecs * 8 <= 144 && a.nv == 3 * a.seg_vecs, CIDNET_ERR_INVALID, "dw3: bad channel layout");
#ifndef CIDNET_ACT_BF16
    const size_t smem2 = (size_t)4 * 4 * kDwThreads * 16 + (size_t)9 * a.nv * 8 * sizeof(act_t);
    const int gx = ceil_div(ceil_div(a.W, 2) * a.nv, kDwThreads);
    a.rows_per_cta = pick_strip_rows(a.H, (long long)gx * a.B * a.nprob, 4 * device_sm_count(), 2, 2, 16, 96);
    dim3 grid2(gx, ceil_div(a.H, a.rows_per_cta), a.B * a.nprob);
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(dw3x3_cpasync2_kernel<4, 4>), 64 * 1024);
    if (rc) return rc;
    if ((rc = launch_k(dw3x3_cpasync2_kernel<4, 4>, grid2, dim3(kDwThreads), smem2, stream, a))) return rc;
#else
    const int gx = ceil_div(a.W * a.nv, kDwThreads);
    a.rows_per_cta = pick_strip_rows(a.H, (long long)gx * a.B * a.nprob, 3 * device_sm_count(), 2, 2, 16, 96);
    dim3 grid(gx, ceil_div(a.H, a.rows_per_cta), a.B * a.nprob);
    int rc = launch_k(dw3x3_f32acc_kernel, grid, dim3(kDwThreads), 0, stream, a);
    if (rc) return rc;
#endif
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

// ------------------------------------------------------------------ Gram ----
struct GramArgs {
    CUtensorMap tmQ[2];      // per problem: q  [B][HW][pitch] viewed {C, HW, B}
    CUtensorMap tmK[2];
    float* slab;             // [nprob][B][nsplit][E]: per CTA [heads x 18 x 18 | Cp sum q^2 | Cp sum k^2]
    int C, Cp, heads, hw, chunks_per_cta, nchunks, nsplit, B;
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
    const int mtiles = (C + 127) / 128;
    const int nA = 2 * mtiles;                    // 64-channel blocks of q (blocks beyond C are TMA zero fill)
    const int N = (C + 15) / 16 * 16;             // MMA N (k channels)
    const int nB = (N + 63) / 64;
    const uint32_t stage_bytes = (uint32_t)(nA + nB) * kBlk;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kGramStages * stage_bytes);
    uint64_t* empty = full + kGramStages;
    uint64_t* done = empty + kGramStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    float* s_ss = reinterpret_cast<float*>(smem);                 // [4 warps][2 (q, k)][192 channels] partial sums of squares:
                                                                  // re-uses stage 0 once every MMA has retired

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.y, b = blockIdx.z;
    const int c_begin = blockIdx.x * a.chunks_per_cta;
    const int c_end = min(c_begin + a.chunks_per_cta, a.nchunks);
    const int niter = c_end - c_begin;            // >= 1: the host sizes the grid so that no CTA is empty

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
            // a stage is released by the MMA commit AND by the four statistics warps that read it
            for (int s = 0; s < kGramStages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 5); }
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
    ptx::pdl_wait();
    ptx::pdl_trigger();
    const uint32_t tmem_base = *tmem_slot;
    float* slab = a.slab + (((long long)prob * a.B + b) * a.nsplit + blockIdx.x) * (long long)(a.heads * 324 + 2 * a.Cp);

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
            __syncwarp();
        }
    } else {
        const int q4 = warp & 3;
        // ---- sum q^2 / sum k^2 while the tiles are resident: warp q4 takes pixels [16 q4, 16 q4 + 16) of every
        // 64-pixel block, lane = channel pair (one conflict-free LDS.32 per pixel: the 32 lanes read the 32 words
        // of one 128-byte row; the swizzle only permutes its 16-byte chunks).  FHFMA: exact 16-bit products, fp32 sums.
        float sq[3][2], sk[3][2];
#pragma unroll
        for (int j = 0; j < 3; ++j) { sq[j][0] = sq[j][1] = sk[j][0] = sk[j][1] = 0.f; }
        const uint32_t lane_off = (uint32_t)(lane & 3) * 4u;
        const uint32_t lane_chunk = (uint32_t)(lane >> 2);
        for (int i = 0; i < niter; ++i) {
            const int s = i % kGramStages;
            ptx::mbar_wait(&full[s], (uint32_t)(i / kGramStages) & 1u);
            const uint8_t* st = smem + (size_t)s * stage_bytes;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                if (j < nA_load) {
                    const uint8_t* blk = st + j * kBlk;
#pragma unroll
                    for (int p = 0; p < 16; ++p) {
                        const uint32_t px = (uint32_t)(q4 * 16 + p);
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(blk + px * 128u + ((lane_chunk ^ (px & 7u)) << 4) + lane_off);
                        fhfma2(sq[j][0], sq[j][1], v, v);
                    }
                }
                if (j < nB) {
                    const uint8_t* blk = st + (nA + j) * kBlk;
#pragma unroll
                    for (int p = 0; p < 16; ++p) {
                        const uint32_t px = (uint32_t)(q4 * 16 + p);
                        const uint32_t v = *reinterpret_cast<const uint32_t*>(blk + px * 128u + ((lane_chunk ^ (px & 7u)) << 4) + lane_off);
                        fhfma2(sk[j][0], sk[j][1], v, v);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[s]);
        }
        // all MMAs have retired (and every TMA load was consumed before that), and -- the barrier -- all four statistics
        // warps have finished READING the last chunk: only then may stage 0 be re-used as s_ss
        ptx::mbar_wait(done, 0);
        ptx::tc_fence_after();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // per-warp partials -> shared memory; the four are summed in a fixed order below
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float* dq = s_ss + (q4 * 2 + 0) * 192 + j * 64 + 2 * lane;
            float* dk = s_ss + (q4 * 2 + 1) * 192 + j * 64 + 2 * lane;
            dq[0] = sq[j][0]; dq[1] = sq[j][1];
            dk[0] = sk[j][0]; dk[1] = sk[j][1];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        {
            float* sqk = slab + a.heads * 324;
            for (int i = threadIdx.x - 64; i < 2 * a.Cp; i += 128) {
                const int which = i >= a.Cp ? 1 : 0, c = i - which * a.Cp;
                float v = 0.f;
                if (c < C) v = ((s_ss[(0 * 2 + which) * 192 + c] + s_ss[(1 * 2 + which) * 192 + c]) +
                                s_ss[(2 * 2 + which) * 192 + c]) + s_ss[(3 * 2 + which) * 192 + c];
                sqk[i] = v;
            }
        }
        // ---- the Gram: only the per-head 18x18 diagonal blocks of the C x C product are wanted.  The 32 q channels
        // of a warp belong to at most 3 heads, so the warp reads just the k-column window of those heads (<= 5 x 16
        // columns instead of all N); every thread stores the 18 columns of its own q channel.
        const int r = q4 * 32 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        for (int mt = 0; mt < mtiles; ++mt) {
            const int ch_lo = mt * 128 + q4 * 32;                 // warp-uniform
            if (ch_lo >= C) break;
            const int ch_hi = min(ch_lo + 31, C - 1);
            const int col_lo = ((ch_lo / 18) * 18) & ~15;
            const int col_hi = min((ch_hi / 18) * 18 + 18, N);
            const int ch = mt * 128 + r;
            const int head = ch / 18;
            const int kc0 = head * 18;                            // this thread's first k column
            float* rowp = slab + (long long)ch * 18;              // (head * 18 + qi) * 18 == ch * 18
            const bool row_ok = ch < C;
            for (int cc = col_lo; cc < col_hi; cc += 16) {
                float v[16];
                ptx::tmem_ld16(taddr + mt * N + cc, v);
                const int rel = cc - kc0;                         // column of v[0] relative to the thread's block
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    if (row_ok && (unsigned)(rel + j) < 18u)      // rel is even (cc and kc0 are): pairs never straddle the block
                        *reinterpret_cast<float2*>(rowp + rel + j) = make_float2(v[j], v[j + 1]);
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

int gram_max_slab_entries(int nprob_times_B) { return std::max(kGramMaxCtas, nprob_times_B); }

int launch_gram(const GramLaunch& L, cudaStream_t stream, int* nsplit_out) {
    CIDNET_CHECK(L.C == 36 || L.C == 72 || L.C == 144, CIDNET_ERR_INVALID, "gram: C must be 36/72/144");
    CIDNET_CHECK(L.slab != nullptr && nsplit_out != nullptr, CIDNET_ERR_INVALID, "gram: null slab");
    GramArgs a;
    memset(&a, 0, sizeof a);
    a.C = L.C; a.Cp = act_pitch(L.C); a.heads = L.heads; a.hw = L.H * L.W; a.B = L.B;
    a.nchunks = ceil_div(a.hw, 64);
    a.slab = L.slab;
    for (int p = 0; p < L.nprob; ++p) {
        const uint64_t dims[3] = {(uint64_t)L.C, (uint64_t)a.hw, (uint64_t)L.B};
        const uint64_t img_px = L.img_stride_px > 0 ? (uint64_t)L.img_stride_px : (uint64_t)a.hw;
        const uint64_t str[2] = {(uint64_t)L.pitch * sizeof(act_t), (uint64_t)L.pitch * sizeof(act_t) * img_px};
        const uint32_t box[3] = {64, 64, 1};
        int rc = encode_map_generic(&a.tmQ[p], L.q[p], 3, dims, str, box);
        if (rc) return rc;
        if ((rc = encode_map_generic(&a.tmK[p], L.k[p], 3, dims, str, box))) return rc;
    }
    const int nA = 2 * ceil_div(L.C, 128), nB = ceil_div(round_up(L.C, 16), 64);
    const size_t smem = 1024 + (size_t)kGramStages * (nA + nB) * kBlk + 64;
    // split-K over CTAs, sized for ONE wave: shared memory allows 3 / 2 / 1 CTAs per SM for C = 36 / 72 / 144 (and
    // C = 144 needs all 512 TMEM columns).  A partial second wave doubled the kernel time at the coarse levels.
    const int sms = device_sm_count();
    const int per_sm = L.C >= 144 ? 1 : (int)std::min<size_t>(3, (227 * 1024) / smem);
    int nsplit = std::min(sms * per_sm, kGramMaxCtas) / (L.nprob * L.B);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > a.nchunks) nsplit = a.nchunks;
    a.chunks_per_cta = ceil_div(a.nchunks, nsplit);
    nsplit = ceil_div(a.nchunks, a.chunks_per_cta);          // no empty CTA: every slab entry gets written
    a.nsplit = nsplit;
    *nsplit_out = nsplit;
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(gram_kernel), 200 * 1024);
    if (rc) return rc;
    dim3 grid(nsplit, L.nprob, L.B);
    if ((rc = launch_k(gram_kernel, grid, dim3(kGramThreads), smem, stream, a))) return rc;
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

// ------------------------------------------------- slab reduction + fold ----
// Sum `n` float4 / float values spaced `stride` floats apart, in index order.  The loads of a batch are issued back to
// back BEFORE the first addition (volatile asm keeps them ahead of the `pin` statements every addition depends on): left
// to itself the compiler interleaved each L2 load with the additions of the previous one -- three loads in flight, one L2
// round trip per three slab entries, 23 us for the L1 fold (SASS / ncu, profiles/r02_summary.md).  The additions stay
// strictly sequential -> the result depends on nothing but the data.
__device__ __forceinline__ float4 ldcg4_issue(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void pin4(float4& a, float4& b, float4& c, float4& d) {
    asm volatile("" : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w),
                      "+f"(c.x), "+f"(c.y), "+f"(c.z), "+f"(c.w), "+f"(d.x), "+f"(d.y), "+f"(d.z), "+f"(d.w));
}
__device__ __forceinline__ float4 ordered_sum4(const float* p, long long stride, int n) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 16 <= n; s += 16) {
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = ldcg4_issue(p + (long long)(s + u) * stride);
#pragma unroll
        for (int u = 0; u < 16; u += 4) pin4(v[u], v[u + 1], v[u + 2], v[u + 3]);
#pragma unroll
        for (int u = 0; u < 16; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (s < n) {                              // ragged tail: still one batch, missing entries contribute +0
        float4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (s + u < n) ? ldcg4_issue(p + (long long)(s + u) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 16; u += 4) pin4(v[u], v[u + 1], v[u + 2], v[u + 3]);
#pragma unroll
        for (int u = 0; u < 16; ++u) if (s + u < n) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    return acc;
}
__device__ __forceinline__ float ordered_sum1(const float* p, long long stride, int n) {
    float acc = 0.f;
    for (int s = 0; s < n; s += 16) {
        float4 v[4];                          // 16 scalars of a batch, packed so that pin4 covers them
        float* f = reinterpret_cast<float*>(v);
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            f[u] = 0.f;
            if (s + u < n) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(f[u]) : "l"(p + (long long)(s + u) * stride));
        }
        pin4(v[0], v[1], v[2], v[3]);
#pragma unroll
        for (int u = 0; u < 16; ++u) if (s + u < n) acc += f[u];
    }
    return acc;
}

static constexpr int kFoldThreads = 512;
static constexpr int kFoldLanes = 4;           // split lanes per CTA (128 threads each): one 16-deep batch of loads per lane
static constexpr int kFoldCluster = 4;         // CTAs per (problem, head, image): thread-block cluster along grid.z

// grid (nprob * heads, B, kFoldCluster), cluster (1, 1, kFoldCluster).  The kernel is a chain of dependent latencies (slab
// loads -> softmax -> W_o -> store), not work; ncu on the first one-CTA-per-head version: 40 cycles of warp latency per
// issued instruction, ~1000 dependent instructions per warp, 23 us.  So the chain is cut four ways twice:
//   * the CTAs of a cluster each reduce a QUARTER of the split-K slab entries of their head, with four 128-thread lanes
//     taking an eighth of that each -- nsplit <= 256 is one batch of <= 16 independent float4 loads per thread;
//   * the four per-CTA partials are exchanged through distributed shared memory (one cluster barrier) and summed by
//     every CTA in CTA-rank order; each CTA then runs the (tiny) softmax and folds its own quarter of the output rows.
// All summation orders are fixed (split order inside a lane, lane order, CTA-rank order): bit-reproducible.
__global__ void __launch_bounds__(kFoldThreads, 1)
cab_fold_kernel(const CabFoldArgs a) {
    __shared__ __align__(16) float s_lane[kFoldLanes][368];    // per lane: [324 Gram | 18 sq | 18 sk] (+ pad)
    __shared__ __align__(16) float s_cta[368];                 // this CTA's partial (read by the whole cluster)
    __shared__ float s_tot[368];
    __shared__ float s_attn[18 * 18];
    __shared__ float s_wo[36 * 18];                            // W_o[o][head * 18 + c] of this CTA's output rows
    const int prob = blockIdx.x / a.heads, head = blockIdx.x - prob * a.heads, b = blockIdx.y;
    const int rank = blockIdx.z;                               // == %cluster_ctarank: the cluster spans grid.z
    const int C = a.C, tid = threadIdx.x;
    const long long E = (long long)a.heads * 324 + 2 * a.Cp;
    const float* slab = a.slab + ((long long)prob * a.B + b) * a.nsplit * E;
    const float* wo = prob ? a.wo[1] : a.wo[0];
    const int rows_per = (a.n_rows + kFoldCluster - 1) / kFoldCluster;         // <= 36
    const int o_begin = rank * rows_per, o_end = min(o_begin + rows_per, a.n_rows);
    const int n_el = max(o_end - o_begin, 0) * 18;             // M elements of this CTA: [rows of the slice][18 columns]
    // W_o is a model constant: staged before the programmatic-dependency wait (it arrives while the Gram kernel drains)
    for (int i = tid; i < n_el; i += kFoldThreads) {
        const int o = o_begin + i / 18;
        s_wo[i] = o < C ? __ldg(wo + o * C + head * 18 + i % 18) : 0.f;
    }
    ptx::pdl_wait();
    ptx::pdl_trigger();
    // ---- this CTA's quarter of the slab entries, four lanes
    {
        const int lane4 = tid >> 7, t = tid & 127;
        const int c0 = (a.nsplit * rank) / kFoldCluster, c1 = (a.nsplit * (rank + 1)) / kFoldCluster;
        const int n = c1 - c0;
        const int s0 = c0 + (n * lane4) / kFoldLanes, s1 = c0 + (n * (lane4 + 1)) / kFoldLanes;
        const float* base = slab + (long long)s0 * E;
        if (t < 81) {
            const float4 v = ordered_sum4(base + head * 324 + 4 * t, E, s1 - s0);
            *reinterpret_cast<float4*>(&s_lane[lane4][4 * t]) = v;
        } else if (t < 81 + 36) {
            const int i = t - 81, which = i / 18, c = i - which * 18;
            s_lane[lane4][324 + i] = ordered_sum1(base + a.heads * 324 + which * a.Cp + head * 18 + c, E, s1 - s0);
        }
    }
    __syncthreads();
    if (tid < 360) s_cta[tid] = ((s_lane[0][tid] + s_lane[1][tid]) + s_lane[2][tid]) + s_lane[3][tid];
    // ---- exchange the per-CTA partials through distributed shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (tid < 360) {
        const uint32_t local = ptx::smem_u32(&s_cta[tid]);
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < kFoldCluster; ++r) {
            uint32_t remote; float v;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
            tot += v;
        }
        s_tot[tid] = tot;
        if (a.raw_out && rank == 0) {          // tests / sharding: the reduced raw statistics of this head
            float* ro = a.raw_out + ((long long)prob * a.B + b) * E;
            if (tid < 324) ro[head * 324 + tid] = tot;
            else { const int i = tid - 324, which = i / 18; ro[a.heads * 324 + which * a.Cp + head * 18 + (i - which * 18)] = tot; }
        }
    }
    // nobody may exit (and release its shared memory) before every CTA of the cluster has read it
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    __syncthreads();
    if (tid < 18) {
        const float inv_nq = rsqrtf(fmaxf(s_tot[324 + tid], 1e-24f));      // 1 / max(sqrt(sum q^2), 1e-12)
        const float temp = __ldg((prob ? a.temp[1] : a.temp[0]) + head);
        float logit[18], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 18; ++j) {
            const float inv_nk = rsqrtf(fmaxf(s_tot[342 + j], 1e-24f));
            logit[j] = s_tot[tid * 18 + j] * (inv_nq * inv_nk) * temp;
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
    // M[o][head*18 + j] = sum_c Wo[o][head*18 + c] * attn[c][j]; rows o >= C (N padding) are zero
    act_t* m = (prob ? a.m_out[1] : a.m_out[0]) + (long long)b * a.n_rows * a.kt + head * 18;
    for (int i = tid; i < n_el; i += kFoldThreads) {
        const int r = i / 18, j = i - r * 18;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 18; ++c) acc = fmaf(s_wo[r * 18 + c], s_attn[c * 18 + j], acc);
        m[(long long)(o_begin + r) * a.kt + j] = f2act(acc);
    }
    // the K padding columns [C, kt) of every row: written (as zeros) by the last head's CTAs
    if (head == a.heads - 1) {
        const int padw = a.kt - C;
        act_t* mp = (prob ? a.m_out[1] : a.m_out[0]) + (long long)b * a.n_rows * a.kt + C;
        for (int i = tid; i < max(o_end - o_begin, 0) * padw; i += kFoldThreads) {
            const int o = o_begin + i / padw, j = i % padw;
            mp[(long long)o * a.kt + j] = f2act(0.f);
        }
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int launch_cab_fold(const CabFoldArgs& a, cudaStream_t stream) {
    CIDNET_CHECK(a.C <= 144 && a.C == a.heads * 18, CIDNET_ERR_INVALID, "fold: C must be heads * 18 <= 144");
    CIDNET_CHECK(a.slab != nullptr && a.nsplit >= 1, CIDNET_ERR_INVALID, "fold: null slab");
    CIDNET_CHECK(ceil_div(a.n_rows, kFoldCluster) <= 36, CIDNET_ERR_INVALID, "fold: too many output rows per CTA");
    dim3 grid(a.nprob * a.heads, a.B, kFoldCluster);
    return launch_k_cluster(cab_fold_kernel, grid, dim3(kFoldThreads), 0, stream, dim3(1, 1, kFoldCluster), a);
}

// Row-strip sharding: slab -> this rank's raw partial [Gram | sum q^2 | sum k^2] per (problem, image), laid out
// exactly like ONE slab entry, so that after the host's all-reduce the fold runs on it with nsplit = 1.
__global__ void __launch_bounds__(256, 1)
cab_reduce_kernel(const float* __restrict__ slab, float* __restrict__ out, int nsplit, int E4) {
    const long long E = 4ll * E4;
    ptx::pdl_wait();
    ptx::pdl_trigger();
    const float* src = slab + (long long)blockIdx.y * nsplit * E;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < E4; i += gridDim.x * 256) {
        const float4 v = ordered_sum4(src + 4 * i, E, nsplit);
        *reinterpret_cast<float4*>(out + (long long)blockIdx.y * E + 4 * i) = v;
    }
}

int launch_cab_reduce(const float* slab, float* out, int nsplit, int E, int nvec, cudaStream_t stream) {
    CIDNET_CHECK(E % 4 == 0, CIDNET_ERR_INVALID, "reduce: E % 4");
    dim3 grid(ceil_div(E / 4, 256), nvec);
    int rc = launch_k(cab_reduce_kernel, grid, dim3(256), 0, stream, slab, out, nsplit, E / 4);
    if (rc) return rc;
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
