// IEL gate: the whole depthwise / tanh / product chain of the IEL block in ONE kernel
// (replaces 3 depthwise convs, 2 tanh, 2 adds, 1 mul and the chunk() views of net/LCA.py:61-65):
//
//   d  = dwconv(t)            (3x3 depthwise on [x1 | x2], zero pad)
//   x1 = tanh(dwconv1(d1)) + d1 ;  x2 = tanh(dwconv2(d2)) + d2 ;  g = x1 * x2
//
// (fp16 build: iel_gate_v6_kernel, packed HFMA2; bf16 build: iel_gate_v4_kernel, fp32 accumulation.  Both are fed by a
// TMA ring of t; the register-shuffle v3, the smem-weight v5 and the cp.async-producer experiments of round 1 were
// measured slower (profiles/r01_summary.md) and are gone.)
#include "iel.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>
#include <cstring>

namespace cidnet {

static constexpr int kCols = 32;       // columns per warp (outputs: lanes 1..30)

#ifdef CIDNET_ACT_BF16
#define CIDNET_FHFMA "fma.rn.f32.bf16"
#else
#define CIDNET_FHFMA "fma.rn.f32.f16"
#endif

// acc[0..1] += lo/hi(a) * lo/hi(b)   (two mixed-precision FMAs, operands stay packed)
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
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));    // one MUFU op, abs error < 2^-10.9
    return y;
}
__device__ __forceinline__ uint4 shfl_up4(const uint4& v) {
    return make_uint4(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1),
                      __shfl_up_sync(0xffffffffu, v.z, 1), __shfl_up_sync(0xffffffffu, v.w, 1));
}
__device__ __forceinline__ uint4 shfl_down4(const uint4& v) {
    return make_uint4(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1),
                      __shfl_down_sync(0xffffffffu, v.z, 1), __shfl_down_sync(0xffffffffu, v.w, 1));
}

struct TRow { uint4 l, c, r; };          // raw 16-bit t values: left / centre / right column
struct DRow { uint4 l, c, r; float cf[8]; };   // d: 16-bit packed left / centre / right (dwconv1/2 operands) + fp32 centre (residual)

// ================================================================================================
// v4: TMA-fed.  A producer warp streams row blocks of t -- {16 channels, 34 columns (1-column halo),
// 4 rows} per half, zero-filled outside the image -- into a SWIZZLE_32B shared-memory ring; the four
// compute warps (x1/x2 x 2 channel vectors, lane = column) read their own column and both neighbours
// with conflict-free LDS.128.  No global loads, edge-lane special cases or prefetch registers in the
// compute warps; d still moves between lanes by shuffle and between the x1/x2 warps by 1 KB of smem.
// ================================================================================================
static constexpr int kRB = 4;                         // rows per TMA box
static constexpr int kV4Stages = 6;
static constexpr int kBoxCols = 40;                   // 34 needed; 40 makes the row pitch (1280 B) a multiple of the 256-byte swizzle period
static constexpr uint32_t kHalfBoxBytes = kRB * kBoxCols * 32;      // 5120
static constexpr uint32_t kV4StageBytes = 2 * kHalfBoxBytes;
static constexpr int kV4Threads = 160;

struct IelV4Args {
    CUtensorMap tmT[2];         // per problem: t viewed {2*hp channels, W, H, B}, box {16, 34, 4, 1}, SWIZZLE_32B
    IelGateArgs g;
};

#ifdef CIDNET_ACT_BF16
__global__ void __launch_bounds__(kV4Threads, 2)
iel_gate_v4_kernel(const __grid_constant__ IelV4Args A) {
    const IelGateArgs& a = A.g;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* ring = smem;                                                  // kV4Stages x kV4StageBytes
    act_t* s_w0 = reinterpret_cast<act_t*>(ring + kV4Stages * kV4StageBytes);   // [9][2][16]
    act_t* s_w12 = s_w0 + 9 * 2 * 16;
    float4* s_x = reinterpret_cast<float4*>(s_w12 + 9 * 2 * 16);          // [vec][parity][plane][lane]
    uint64_t* full = reinterpret_cast<uint64_t*>(s_x + 2 * 2 * 2 * kCols);
    uint64_t* empty = full + kV4Stages;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hp = a.hp, ngroups = hp / 16;
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int cg = blockIdx.x % ngroups, strip = blockIdx.x / ngroups;
    const int c0 = cg * 16;
    const int X0 = strip * (kCols - 2) - 1;               // image column of lane 0
    const int y0 = blockIdx.y * a.rows_per_cta;
    const int y1 = min(y0 + a.rows_per_cta, a.H);
    const int nrows = (y1 - y0) + 4;                      // t rows y0-2 .. y1+1
    const int nblocks = (nrows + kRB - 1) / kRB;

    for (int i = tid; i < 9 * 2 * 16; i += kV4Threads) {
        const int c = i & 15, hf = (i >> 4) & 1, tap = i >> 5;
        s_w0[i] = f2act(a.w0[prob][tap * 2 * hp + hf * hp + c0 + c]);
        s_w12[i] = f2act((hf == 0 ? a.w1[prob] : a.w2[prob])[tap * hp + c0 + c]);
    }
    if (tid == 0) {
        for (int s = 0; s < kV4Stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 4); }
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&A.tmT[prob]);
    }
    __syncthreads();
    ptx::pdl_wait();          // weights (above) are constants; t is the previous kernel's output
    ptx::pdl_trigger();

    if (warp == 4) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int k = 0; k < nblocks; ++k) {
                const int s = k % kV4Stages;
                ptx::mbar_wait(&empty[s], ((k / kV4Stages) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&full[s], 2 * kHalfBoxBytes);
                uint8_t* dst = ring + (size_t)s * kV4StageBytes;
                const int yb = y0 - 2 + k * kRB;
                ptx::tma_load_4d(dst, &A.tmT[prob], &full[s], c0, X0 - 1, yb, b);
                ptx::tma_load_4d(dst + kHalfBoxBytes, &A.tmT[prob], &full[s], hp + c0, X0 - 1, yb, b);
            }
        }
        return;
    }
    // ---------------------------------------------------------------- compute warps
    const int half = warp >> 1, vec = warp & 1;
    const int x = X0 + lane;
    const bool col_in = x >= 0 && x < a.W;
    const long long hw = (long long)a.H * a.W;
    act_t* gdst = a.g[prob] + (long long)b * hw * hp + c0 + vec * 8;
    const bool writer = half == 0 && lane >= 1 && lane <= kCols - 2 && col_in;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);

    // SWIZZLE_32B: byte offset o -> o ^ (bit 7 of o) << 4.  Row pitch and stage size are multiples of 256 B,
    // so the swizzled offsets of the three columns this lane reads are constants
    auto swz = [](uint32_t o) { return o ^ (((o >> 7) & 1u) << 4); };
    const uint32_t off_l = swz((uint32_t)lane * 32u + (uint32_t)vec * 16u);
    const uint32_t off_c = swz((uint32_t)(lane + 1) * 32u + (uint32_t)vec * 16u);
    const uint32_t off_r = swz((uint32_t)(lane + 2) * 32u + (uint32_t)vec * 16u);
    // t row (relative index j = row - (y0-2)) -> left / centre / right 16-byte vectors from the ring
    auto lds_trow = [&](int j, TRow& t) {
        const int k = j / kRB, rr = j - k * kRB;
        const int s = k % kV4Stages;
        if (rr == 0) ptx::mbar_wait(&full[s], (k / kV4Stages) & 1u);       // first row of a block: wait for the TMA
        const uint8_t* base = ring + (size_t)s * kV4StageBytes + half * kHalfBoxBytes + rr * (kBoxCols * 32);
        t.l = *reinterpret_cast<const uint4*>(base + off_l);
        t.c = *reinterpret_cast<const uint4*>(base + off_c);
        t.r = *reinterpret_cast<const uint4*>(base + off_r);
        if (rr == kRB - 1 || j == nrows - 1) {        // last row of the block consumed: release the stage
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[s]);
        }
    };
    auto dw0_row = [&](const TRow& t, int tap0, float* acc) {
        const uint4* w = reinterpret_cast<const uint4*>(s_w0) + (half * 2 + vec);
        fhfma8(acc, t.l, w[(tap0 + 0) * 4]);
        fhfma8(acc, t.c, w[(tap0 + 1) * 4]);
        fhfma8(acc, t.r, w[(tap0 + 2) * 4]);
    };
    auto dw12_row = [&](const DRow& d, int tap0, float* o) {
        const uint4* w = reinterpret_cast<const uint4*>(s_w12) + (half * 2 + vec);
        fhfma8(o, d.l, w[(tap0 + 0) * 4]);
        fhfma8(o, d.c, w[(tap0 + 1) * 4]);
        fhfma8(o, d.r, w[(tap0 + 2) * 4]);
    };
    auto pack8h = [](const float* f) -> uint4 {
        uint4 raw;
        act_t* o = reinterpret_cast<act_t*>(&raw);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = f2act(f[e]);
        return raw;
    };
    auto iter = [&](int r, const TRow& t0, const TRow& t1, TRow& t2, const DRow& d0, const DRow& d1, DRow& d2) {
        lds_trow(r + 1 - (y0 - 2), t2);
#pragma unroll
        for (int e = 0; e < 8; ++e) d2.cf[e] = 0.f;
        if (r >= 0 && r < a.H && col_in) {             // d is zero outside the image
            dw0_row(t0, 0, d2.cf);
            dw0_row(t1, 3, d2.cf);
            dw0_row(t2, 6, d2.cf);
        }
        d2.c = pack8h(d2.cf);
        d2.l = shfl_up4(d2.c);
        d2.r = shfl_down4(d2.c);
        const int yo = r - 1;
        float xs[8];
        {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = 0.f;
            dw12_row(d0, 0, o);
            dw12_row(d1, 3, o);
            dw12_row(d2, 6, o);
#pragma unroll
            for (int e = 0; e < 8; ++e) xs[e] = tanh_approx(o[e]) + d1.cf[e];
        }
        float4* slot = s_x + ((vec * 2 + (yo & 1)) * 2) * kCols + lane;
        if (half == 1) {
            slot[0] = make_float4(xs[0], xs[1], xs[2], xs[3]);
            slot[kCols] = make_float4(xs[4], xs[5], xs[6], xs[7]);
        }
        asm volatile("bar.sync %0, 64;" :: "r"(1 + vec) : "memory");
        if (writer && yo >= y0) {
            const float4 pa = slot[0], pb = slot[kCols];
            uint4 raw;
            act_t* ov = reinterpret_cast<act_t*>(&raw);
            ov[0] = f2act(xs[0] * pa.x); ov[1] = f2act(xs[1] * pa.y); ov[2] = f2act(xs[2] * pa.z); ov[3] = f2act(xs[3] * pa.w);
            ov[4] = f2act(xs[4] * pb.x); ov[5] = f2act(xs[5] * pb.y); ov[6] = f2act(xs[6] * pb.z); ov[7] = f2act(xs[7] * pb.w);
            *reinterpret_cast<uint4*>(gdst + ((long long)yo * a.W + x) * hp) = raw;
        }
    };
    TRow tA, tB, tC;
    DRow dA, dB, dC;
    dA.l = dA.c = dA.r = zero4; dB.l = dB.c = dB.r = zero4;
#pragma unroll
    for (int e = 0; e < 8; ++e) { dA.cf[e] = 0.f; dB.cf[e] = 0.f; }
    lds_trow(0, tA);
    lds_trow(1, tB);
    for (int r = y0 - 1; r <= y1; r += 3) {
        iter(r, tA, tB, tC, dA, dB, dC);
        if (r + 1 <= y1) iter(r + 1, tB, tC, tA, dB, dC, dA);
        if (r + 2 <= y1) iter(r + 2, tC, tA, tB, dC, dA, dB);
    }
}

#else
// ================================================================================================
// v5 (fp16 build): the v4 data path (TMA-fed SWIZZLE_32B ring, lane = column, LDS neighbours) with the
// whole chain in PACKED fp16: HFMA2 does two multiply-adds per issue slot (the v4 kernel is
// issue-bound: 144 FHFMA + ~150 other instructions per warp-row), accumulators, d, tanh (one MUFU per
// two values, tanh.approx.f16x2) and the product stay packed halves, so there are no pack / unpack or
// zeroing instructions and the x2 -> x1 hand-over is one 16-byte vector.  Precision: the 9-tap sums are
// rounded to fp16 after every tap (two partial chains); measured end to end against the fp32 oracle
// this moves max-abs from ~1.0e-4 to ~1.2e-4 and PSNR from ~100 to ~97 dB (contract: 2e-3 / 50 dB).
// ================================================================================================
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
__device__ __forceinline__ uint32_t htanh2u(uint32_t a) {
    uint32_t d;
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
__device__ __forceinline__ uint4 hmul8(const uint4& t, const uint4& w) {
    return make_uint4(hmul2u(t.x, w.x), hmul2u(t.y, w.y), hmul2u(t.z, w.z), hmul2u(t.w, w.w));
}
__device__ __forceinline__ void hfma8(uint4& acc, const uint4& t, const uint4& w) {
    acc.x = hfma2u(t.x, w.x, acc.x); acc.y = hfma2u(t.y, w.y, acc.y);
    acc.z = hfma2u(t.z, w.z, acc.z); acc.w = hfma2u(t.w, w.w, acc.w);
}
__device__ __forceinline__ uint4 hadd8(const uint4& a, const uint4& b) {
    return make_uint4(hadd2u(a.x, b.x), hadd2u(a.y, b.y), hadd2u(a.z, b.z), hadd2u(a.w, b.w));
}

static constexpr int kV5Stages = 4;

// ================================================================================================
// v6 (fp16 build): v5's data path, but the 18 weight vectors (dwconv + dwconv1/2, 8 channels) live in
// REGISTERS -- ncu on v5 showed the shared-memory data pipe as the top unit (66 %: two wavefronts per
// warp-uniform LDS.128, 36 of the 52 wavefronts per warp-row were weight re-loads).  To pay for the
// 72 weight registers the two 3x3 stages are evaluated in SCATTER form: an arriving row updates the
// three output rows it touches (three running partial sums) instead of three input rows being kept, so
// only the arriving (l, c, r) vectors are live.  Same arithmetic, same fp16 rounding points as v5 up to
// the summation order (row-major chains).
// ================================================================================================
__global__ void __launch_bounds__(kV4Threads, 3)
iel_gate_v6_kernel(const __grid_constant__ IelV4Args A) {
    const IelGateArgs& a = A.g;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* ring = smem;                                                  // kV5Stages x kV4StageBytes
    uint4* s_x = reinterpret_cast<uint4*>(ring + kV5Stages * kV4StageBytes);   // [vec][slot 0..3][lane] packed x2
    uint64_t* full = reinterpret_cast<uint64_t*>(s_x + 2 * 4 * kCols);
    uint64_t* empty = full + kV5Stages;
    act_t* s_w12 = reinterpret_cast<act_t*>(empty + kV5Stages);               // [9][2][16] dwconv1/2 weights

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int hp = a.hp, ngroups = hp / 16;
    const int prob = blockIdx.z % a.nprob, b = blockIdx.z / a.nprob;
    const int cg = blockIdx.x % ngroups, strip = blockIdx.x / ngroups;
    const int c0 = cg * 16;
    const int X0 = strip * (kCols - 2) - 1;               // image column of lane 0
    const int y0 = blockIdx.y * a.rows_per_cta;
    const int y1 = min(y0 + a.rows_per_cta, a.H);
    const int nrows = (y1 - y0) + 4;                      // t rows y0-2 .. y1+1
    const int nblocks = (nrows + kRB - 1) / kRB;

    for (int i = tid; i < 9 * 2 * 16; i += kV4Threads) {
        const int c = i & 15, hf = (i >> 4) & 1, tap = i >> 5;
        s_w12[i] = f2act((hf == 0 ? a.w1[prob] : a.w2[prob])[tap * hp + c0 + c]);
    }
    if (tid == 0) {
        for (int s = 0; s < kV5Stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 4); }
        ptx::fence_barrier_init();
        ptx::prefetch_tensormap(&A.tmT[prob]);
    }
    __syncthreads();
    ptx::pdl_trigger();

    if (warp == 4) {
        if (lane == 0) {
            ptx::pdl_wait();          // t is the previous kernel's output (the weights are constants)
            for (int k = 0; k < nblocks; ++k) {
                const int s = k % kV5Stages;
                ptx::mbar_wait(&empty[s], ((k / kV5Stages) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&full[s], 2 * kHalfBoxBytes);
                uint8_t* dst = ring + (size_t)s * kV4StageBytes;
                const int yb = y0 - 2 + k * kRB;
                ptx::tma_load_4d(dst, &A.tmT[prob], &full[s], c0, X0 - 1, yb, b);
                ptx::tma_load_4d(dst + kHalfBoxBytes, &A.tmT[prob], &full[s], hp + c0, X0 - 1, yb, b);
            }
        }
        return;
    }
    const int half = warp >> 1, vec = warp & 1;
    const int x = X0 + lane;
    const bool col_in = x >= 0 && x < a.W;
    const long long hw = (long long)a.H * a.W;
    act_t* gdst = a.g[prob] + (long long)b * hw * hp + c0 + vec * 8;
    const bool writer = half == 0 && lane >= 1 && lane <= kCols - 2 && col_in;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);

    // dwconv's 9 weight vectors of this thread's 8 channels: packed fp16, in registers for the whole strip;
    // dwconv1/2's are re-read from shared memory (36 registers less -> 128 registers, 3 CTAs / SM)
    uint4 w0[9];
    const uint4* w12p = reinterpret_cast<const uint4*>(s_w12) + (half * 2 + vec);
    {
        const float* p0 = a.w0[prob] + half * hp + c0 + vec * 8;                     // [tap][2*hp]
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float4 a0 = __ldg(reinterpret_cast<const float4*>(p0 + t * 2 * hp));
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(p0 + t * 2 * hp) + 1);
            __half2 h;
            h = __floats2half2_rn(a0.x, a0.y); w0[t].x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(a0.z, a0.w); w0[t].y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(a1.x, a1.y); w0[t].z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2half2_rn(a1.z, a1.w); w0[t].w = *reinterpret_cast<uint32_t*>(&h);
        }
    }

    auto swz = [](uint32_t o) { return o ^ (((o >> 7) & 1u) << 4); };
    const uint32_t off_l = swz((uint32_t)lane * 32u + (uint32_t)vec * 16u);
    const uint32_t off_c = swz((uint32_t)(lane + 1) * 32u + (uint32_t)vec * 16u);
    const uint32_t off_r = swz((uint32_t)(lane + 2) * 32u + (uint32_t)vec * 16u);
    auto lds_trow = [&](int j, TRow& t) {
        const int k = j / kRB, rr = j - k * kRB;
        const int s = k % kV5Stages;
        if (rr == 0) ptx::mbar_wait(&full[s], (k / kV5Stages) & 1u);
        const uint8_t* base = ring + (size_t)s * kV4StageBytes + half * kHalfBoxBytes + rr * (kBoxCols * 32);
        t.l = *reinterpret_cast<const uint4*>(base + off_l);
        t.c = *reinterpret_cast<const uint4*>(base + off_c);
        t.r = *reinterpret_cast<const uint4*>(base + off_r);
        if (rr == kRB - 1 || j == nrows - 1) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&empty[s]);
        }
    };
    // scatter step of one 3x3 stage: the arriving row R is the top row of `n` (new), the middle row of `m`
    // and the bottom row of `f` (finished after this call)
    auto scatter = [&](const uint4* w, const int ws, const TRow& R, uint4& n, uint4& m, uint4& f) {
        n = hmul8(R.l, w[0 * ws]); hfma8(m, R.l, w[3 * ws]); hfma8(f, R.l, w[6 * ws]);
        hfma8(n, R.c, w[1 * ws]);  hfma8(m, R.c, w[4 * ws]); hfma8(f, R.c, w[7 * ws]);
        hfma8(n, R.r, w[2 * ws]);  hfma8(m, R.r, w[5 * ws]); hfma8(f, R.r, w[8 * ws]);
    };
    // iteration j (= arriving t row y0-2+j): finishes d(row y0-3+j) and output row y0-4+j
    //   pn/pm/pf, qn/qm/qf: running sums of the two stages (roles rotate in the caller); dprev = centre of
    //   the previous d row (the residual of the output row finished here)
    auto iter = [&](int j, uint4& pn, uint4& pm, uint4& pf, uint4& qn, uint4& qm, uint4& qf, uint4& dprev) {
        TRow T;
        lds_trow(j, T);
        scatter(w0, 1, T, pn, pm, pf);
        const int rd = y0 - 3 + j;                                    // the d row finished now
        TRow D;
        D.c = (rd >= 0 && rd < a.H && col_in) ? pf : zero4;           // d is zero outside the image
        D.l = shfl_up4(D.c);
        D.r = shfl_down4(D.c);
        scatter(w12p, 4, D, qn, qm, qf);
        const int yo = rd - 1;                                         // the output row finished now
        const uint4 xs = make_uint4(hadd2u(htanh2u(qf.x), dprev.x), hadd2u(htanh2u(qf.y), dprev.y),
                                    hadd2u(htanh2u(qf.z), dprev.z), hadd2u(htanh2u(qf.w), dprev.w));
        dprev = D.c;
        uint4* slot = s_x + (vec * 4 + (j & 3)) * kCols + lane;
        if (half == 1) *slot = xs;
        asm volatile("bar.sync %0, 64;" :: "r"(1 + vec) : "memory");
        if (writer && yo >= y0 && yo < y1)
            *reinterpret_cast<uint4*>(gdst + ((long long)yo * a.W + x) * hp) = hmul8(xs, *slot);
    };
    uint4 pA = zero4, pB = zero4, pC = zero4, qA = zero4, qB = zero4, qC = zero4, dprev = zero4;
    // rows j = 0 .. nrows-1; the running-sum roles rotate with period 3 (no register moves)
    for (int j = 0; j < nrows; j += 3) {
        iter(j, pA, pB, pC, qA, qB, qC, dprev);
        if (j + 1 < nrows) iter(j + 1, pC, pA, pB, qC, qA, qB, dprev);
        if (j + 2 < nrows) iter(j + 2, pB, pC, pA, qB, qC, qA, dprev);
    }
}
#endif  // CIDNET_ACT_BF16

int encode_map_generic_swz(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, int swizzle_bytes);   // conv_gemm.cu

int launch_iel_gate(const IelGateArgs& a_in, cudaStream_t stream) {
    IelGateArgs a = a_in;
    CIDNET_CHECK(a.hp % 16 == 0, CIDNET_ERR_INVALID, "iel: hp % 16");
    const int strips = ceil_div(a.W, kCols - 2);
#ifndef CIDNET_ACT_BF16
    const int ctas_per_sm = 3;
#else
    const int ctas_per_sm = 2;
#endif
    a.rows_per_cta = pick_strip_rows(a.H, (long long)strips * (a.hp / 16) * a.B * a.nprob, ctas_per_sm * device_sm_count(), 4, 13, 12, 128);   // fixed ~ 5 us of prologue + ring fill per CTA = ~13 rows (measured: L1 / L3 CTA times)
    IelV4Args A;
    memset(&A, 0, sizeof A);
    A.g = a;
    const long long hw = (long long)a.H * a.W;
    for (int p = 0; p < a.nprob; ++p) {
        const uint64_t pb = (uint64_t)2 * a.hp * sizeof(act_t);
        const uint64_t dims[4] = {(uint64_t)2 * a.hp, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.B};
        const uint64_t str[3] = {pb, pb * a.W, pb * hw};
        const uint32_t box[4] = {16, (uint32_t)kBoxCols, (uint32_t)kRB, 1};
        int rc = encode_map_generic_swz(&A.tmT[p], a.t[p], 4, dims, str, box, 32);
        if (rc) return rc;
    }
    dim3 grid(strips * (a.hp / 16), ceil_div(a.H, a.rows_per_cta), a.B * a.nprob);
    int rc;
#ifndef CIDNET_ACT_BF16
    const size_t smem6 = 1024 + (size_t)kV5Stages * kV4StageBytes + 2 * 4 * kCols * sizeof(uint4) +
                         2 * kV5Stages * sizeof(uint64_t) + 9 * 2 * 16 * sizeof(act_t) + 64;
    if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(iel_gate_v6_kernel), (int)smem6))) return rc;
    if ((rc = launch_k(iel_gate_v6_kernel, grid, dim3(kV4Threads), smem6, stream, A))) return rc;
#else
    const size_t smem = 1024 + (size_t)kV4Stages * kV4StageBytes + 2 * 9 * 2 * 16 * sizeof(act_t) +
                        2 * 2 * 2 * kCols * sizeof(float4) + 2 * kV4Stages * sizeof(uint64_t) + 64;
    if ((rc = ensure_dynamic_smem(reinterpret_cast<const void*>(iel_gate_v4_kernel), (int)smem))) return rc;
    if ((rc = launch_k(iel_gate_v4_kernel, grid, dim3(kV4Threads), smem, stream, A))) return rc;
#endif
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
