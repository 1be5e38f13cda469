// Persistent TMA-fed tcgen05 implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels, N] += A[128 pixels, K] * B[N, K]^T      (fp16/bf16 operands, fp32 accumulate in TMEM)
//
// One CTA per SM (per N block) loops over tiles of 128 output pixels (a TH x TW rectangle of one
// image, or 128 consecutive pixels for "flat" 1x1 layers).
//   warp 0   : TMA producer.  For every (tap, 64-channel chunk) ONE box load of the NHWC activation
//              (shifted by the tap offset; TMA zero-fills out-of-image pixels -- that IS the conv's zero
//              padding -- and channels >= Cin), landing in the canonical SWIZZLE_128B K-major layout.
//              The packed weights are loaded ONCE per CTA and stay resident in shared memory when they
//              fit; otherwise (and for per-image weights) they stream through the same ring.
//              An optional second K source (in2 / wt2, e.g. identity weights) appends K chunks: that is
//              how residual adds are done inside the MMA instead of in the epilogue.
//   warp 1   : allocates TMEM (two accumulator buffers), a single thread issues tcgen05.mma
//              (M=128, N=block_n, K=16); tcgen05.commit releases smem stages / publishes accumulators.
//   warps 2-5: epilogue, overlapped with the next tile's MMAs through the TMEM double buffer:
//              tcgen05.ld (thread == pixel row) -> fused tail -> 16-bit -> swizzled smem staging ->
//              TMA store (64 channels x tile; channels >= Cout and out-of-image pixels are clipped).
//                EPI_STORE  [-> PReLU]
//                EPI_LN     LayerNorm folded in: rstd*(acc-mean*wsum)+bias, statistics read from the A
//                           tile while it is resident (the epilogue warps co-own the stage's release)
//                EPI_DOWN   conv3x3 -> bilinear x0.5 (align_corners) -> PReLU; the tile is 16 columns x
//                           8 ROW PAIRS, even rows in accumulator 0 and odd rows in accumulator 1 (5-D
//                           TMA view) -> thread-local vertical lerp + one shuffle for the horizontal
//                EPI_UP     1x1 on the skip + bilinear x2 of the (pre-composed) low-res conv -> PReLU
//
// Replaces nn.Conv2d calls at net/LCA.py:13,15,17,51,57 and net/transformer_utils.py:39,58,60
// together with the LayerNorm (:25-28), UpsamplingBilinear2d (:40,:59), cat (:64), PReLU (:43,:66) and
// the residual adds (LCA.py:79,91-92) around them.
#include "conv_gemm.cuh"
#include "ptx_sm100.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace cidnet {

// ------------------------------------------------------------------ params ---
struct ConvGemmArgs {
    CUtensorMap tmA, tmA2, tmB, tmB2, tmOut, tmUp, tmA_t, tmB_t;
    int taps, kchunks, kchunks2, cin, cin2;
    int Hv, Wv, TH, TW, tiles_x, tiles_y, num_tiles;
    int n_out, block_n, stages, per_image_w, b_resident, stg_bufs;
    float2* sa_stats;                // EPI_UP, MSSA variant: per-pixel (mean, max) over the output channels (SpatialAttention input)
    int w_early;                     // resident weights are model constants: request them before the programmatic-dependency wait
    int halo;                        // 3x3: tile + halo loaded once per channel chunk, taps = row-shifted views
    int wring;                       // halo mode with STREAMED weights: stages of the weight ring (one [block_n x 64] chunk per
                                     // (tap, channel chunk)); 0 = weights resident (or riding with the A stages)
    // channels per TMA box (64, or 48 for the C = 36 layers).  Measured on B200: with an inner box narrower than the swizzle
    // span TMA still lays the rows out at the span's pitch (128 B), i.e. exactly the canonical SWIZZLE_128B K-major tile with
    // the unused tail of every row left untouched -- and the cost of a box is per BYTE (profiles/r01_tma_rowrate.txt), so a
    // 48-channel box moves 25 % fewer bytes through the TMA unit than the zero-filled 64-channel one.
    int boxc, boxc2, boxc_out, boxc_up;
    // halo mode, Cin = 72 / 144: the LAST channel chunk holds 8 / 16 real channels -- its halo tile and its weight chunks
    // come through 16-channel boxes (tmA_t / tmB_t; same smem tiles, a quarter of the TMA bytes).  0 = no tail box.
    int tailc;
    int w_real;
    const float* bias; const float* wsum; float ln_eps;
    int up_H, up_W, up_chunks; float up_ry, up_rx;   // EPI_UP: low-res source (staged by TMA next to every A tile)
    float prelu; int use_prelu;
    float down_ry, down_rx; int in_H, in_W;
    // row-strip sharding: global row index of local output row 0 (DOWN: on the half-resolution grid,
    // UP: on the full-resolution grid), global row of local row 0 of the low-res UP source and its
    // global row count.  All zero / == local sizes when the image is not sharded.
    int down_row0, up_row0, up_src_row0, up_Hg;
};

// Epilogue warpgroups (template parameter kEpiGroups): group g owns TMEM accumulator buffer g and drains the tiles
// j with j % kEpiGroups == g.  The per-tile cost of the memory-bound layers is the epilogue's dependent chain
// (tcgen05.ld -> math -> pack -> st.shared -> fence -> bar -> TMA store).  2 groups are launched; see launch_conv_gemm
// for the 3-group experiment.
static constexpr int kMaxEpiGroups = 3;
static constexpr int kMaxWRing = 12;
static constexpr uint32_t kSubTileBytes = 128 * 128;   // 128 rows x 64 elements x 2 B
static constexpr uint32_t kStagingBytes = 128 * 128;   // one 64-channel output block of a tile
static constexpr uint32_t kHaloTileBytes = 11 * 16 * 128;   // halo mode: 11 rows x 16 columns x 64 channels
// EPI_UP: the low-res pixels an 8 x 16 output tile interpolates from (align_corners x2: <= 6 rows x 10 columns) are
// one TMA box per 64 channels, staged behind the A tile of the same pipeline stage
static constexpr int kUpBoxW = 10, kUpBoxH = 6;
static constexpr uint32_t kUpBoxBytes = kUpBoxW * kUpBoxH * 128;   // what the TMA delivers (zero fill included)
static constexpr uint32_t kUpTileBytes = 8192;                      // its slot (1024-byte aligned for the swizzle)

__device__ __forceinline__ float prelu_f(float v, float slope) { return v >= 0.f ? v : v * slope; }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {      // one F2FP instruction
#ifdef CIDNET_ACT_BF16
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
#else
    uint32_t d;                                                       // saturating: +-65504 instead of inf (common.cuh f2act)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
#endif
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
#ifdef CIDNET_ACT_BF16
#define CIDNET_FHFMA "fma.rn.f32.bf16"
#define CIDNET_ONE_ONE 0x3F803F80u
#else
#define CIDNET_FHFMA "fma.rn.f32.f16"
#define CIDNET_ONE_ONE 0x3C003C00u
#endif
// acc0/1 += lo/hi(a) * lo/hi(b): mixed-precision FMA (SASS FHFMA), 16-bit operands stay packed
__device__ __forceinline__ void fhfma2(float& acc0, float& acc1, uint32_t a, uint32_t b) {
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\t"
        "mov.b32 {al, ah}, %2;\n\t"
        "mov.b32 {bl, bh}, %3;\n\t"
        CIDNET_FHFMA " %0, al, bl, %0;\n\t"
        CIDNET_FHFMA " %1, ah, bh, %1;\n\t}"
        : "+f"(acc0), "+f"(acc1) : "r"(a), "r"(b));
}

// Division-free bookkeeping.  sm_100 has no integer divide: every `t / tiles_per_img`, `trem % tiles_x`, `it % stages` was
// a ~20-instruction I2F / MUFU.RCP / F2I sequence with a ~100-cycle dependent latency, on the single producer / issuer
// thread and on every epilogue thread once per tile or per stage (18 of them in the 1x1 kernel: 360 of its 1832 SASS
// instructions).  TileWalk decodes the first tile with real divisions and then steps by gridDim.x with carries; RingPos
// counts a ring position and its mbarrier phase.
struct TileWalk {
    int t, img, ty, tx;                // linear tile index and its (image, tile row, tile column)
    int st, s_img, s_ty, s_tx;         // the step (gridDim.x), decomposed the same way
    int tiles_x, tiles_y;
    __device__ __forceinline__ TileWalk(int t0, int step, int tiles_x_, int tiles_y_) : tiles_x(tiles_x_), tiles_y(tiles_y_) {
        const int per_img = tiles_x * tiles_y;
        t = t0; img = t0 / per_img; { const int r = t0 - img * per_img; ty = r / tiles_x; tx = r - ty * tiles_x; }
        st = step; s_img = step / per_img; { const int r = step - s_img * per_img; s_ty = r / tiles_x; s_tx = r - s_ty * tiles_x; }
    }
    __device__ __forceinline__ void next() {
        t += st;
        tx += s_tx; if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        ty += s_ty; if (ty >= tiles_y) { ty -= tiles_y; ++img; }
        img += s_img;
    }
};
struct RingPos {
    uint32_t s = 0, ph = 0;            // stage index, phase bit
    __device__ __forceinline__ void advance(uint32_t n) { if (++s == n) { s = 0; ph ^= 1u; } }
    __device__ __forceinline__ void advance_by(uint32_t k, uint32_t n) { s += k; while (s >= n) { s -= n; ph ^= 1u; } }
};

// kHalo: 0 = 1x1 / tap-shifted tiles, 1 = 3x3 from one halo tile with resident weights, 2 = halo tile + streamed weight ring.
// A TEMPLATE parameter on purpose: the halo paths unroll 9 taps x 2 accumulators x 4 k-steps of tcgen05.mma issue code; as a
// run-time flag that code sat in every instantiation (the STORE kernel grew from 1800 to 3120 SASS instructions when the weight
// ring was added) and every 1x1 layer lost 2-4 us per launch to it (measured: same-box A/B against the previous library).
template <int kMode, int kEpiGroups, int kHalo>
__global__ void __launch_bounds__(64 + 128 * kEpiGroups, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmArgs a) {
    constexpr bool kHaloOn = kHalo != 0;
    constexpr bool kWRing = kHalo == 2;
    // 1024-byte alignment is required by the SWIZZLE_128B TMA / UMMA tiles; using the array directly
    // (no integer round trip) keeps the accesses in the shared state space (LDS / STS, not generic LD / ST)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    constexpr int kSub = (kMode == EPI_DOWN) ? 2 : 1;
    constexpr bool kUp = (kMode == EPI_UP || kMode == EPI_UP_SA);      // EPI_UP_SA: EPI_UP + per-pixel channel (mean, max)
    constexpr bool kSa = (kMode == EPI_UP_SA);
    const uint32_t a_stage = kSub * (kHaloOn ? kHaloTileBytes : kSubTileBytes) + (kUp ? a.up_chunks * kUpTileBytes : 0u);
    const uint32_t b_chunk = (uint32_t)a.block_n * 128u;
    const int stages = a.stages;
    const int k1 = a.taps * a.kchunks;              // primary K chunks (weight chunks)
    const int nbchunks = k1 + a.kchunks2;           // weight chunks resident / streamed
    const int kiters = kHaloOn ? a.kchunks : nbchunks;   // pipeline iterations (A stages) per tile
    uint8_t* smBres = smem;                                                // resident weights (optional)
    uint8_t* smA = smBres + (a.b_resident ? (size_t)nbchunks * b_chunk : 0);
    uint8_t* smB = smA + (size_t)stages * a_stage;                         // streamed weights (optional)
    uint8_t* smOut = smB + (a.b_resident ? 0 : (size_t)(kWRing ? a.wring : stages) * b_chunk);  // 2 staging buffers per epilogue group
    uint64_t* full = reinterpret_cast<uint64_t*>(smOut + (size_t)a.stg_bufs * kEpiGroups * kStagingBytes);
    uint64_t* empty = full + stages;
    uint64_t* bfull = empty + stages;
    uint64_t* tmem_full = bfull + 1;                       // [kEpiGroups]
    uint64_t* tmem_empty = tmem_full + kEpiGroups;         // [kEpiGroups]
    // a_ready[buf]: "every pipeline stage of the tile that accumulates into TMEM buffer `buf` has landed".  The LN / UP
    // epilogues read the A tile / the low-res box straight from the ring; they wait HERE, never on the ring's `full`
    // barriers.  A parity wait on full[s] is only sound when the PREVIOUS phase of that stage is known to be complete:
    // true for the MMA warp (it consumes the ring in order), not for an epilogue group that skips the other groups' tiles
    // -- with fewer tiles in the ring than groups (or an smem-limited ring) the wait could succeed on the old phase and
    // the LayerNorm statistics were read from the previous tile's data (the round-1 "3 epilogue groups lose parity"
    // failure, profiles/r02_summary.md).  The MMA warp arrives once per tile, so a group sees consecutive phases.
    uint64_t* a_ready = tmem_empty + kEpiGroups;           // [kEpiGroups]
    uint64_t* wfull = a_ready + kEpiGroups + (kEpiGroups & 1);   // [kMaxWRing] weight ring (halo mode with streamed weights)
    uint64_t* wempty = wfull + kMaxWRing;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wempty + kMaxWRing);   // (s_bias stays 16-byte aligned)
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);
    const int bn32 = (a.block_n + 31) & ~31;
    float* s_wsum = s_bias + bn32;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * a.block_n;
    const uint32_t acc_cols = (uint32_t)(kSub * a.block_n);    // columns of one accumulator buffer

    uint32_t ncols = 32;
    while (ncols < kEpiGroups * acc_cols) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&a.tmA);
        ptx::prefetch_tensormap(&a.tmB);
        ptx::prefetch_tensormap(&a.tmOut);
        if (a.kchunks2) { ptx::prefetch_tensormap(&a.tmA2); ptx::prefetch_tensormap(&a.tmB2); }
        if (a.tailc) { ptx::prefetch_tensormap(&a.tmA_t); ptx::prefetch_tensormap(&a.tmB_t); }
        if (kUp) ptx::prefetch_tensormap(&a.tmUp);
    }
    if (warp == 1) {
        if (lane == 0) {
            // MMA commit (+ the 4 epilogue warps of the tile's group for LN / UP, which read the stage themselves)
            const uint32_t empty_count = (kMode == EPI_LN || kUp) ? 5u : 1u;
            for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], empty_count); }
            ptx::mbar_init(bfull, 1);
            for (int i = 0; i < kEpiGroups; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 4); ptx::mbar_init(&a_ready[i], 1); }
            if (kWRing) for (int i = 0; i < a.wring; ++i) { ptx::mbar_init(&wfull[i], 1); ptx::mbar_init(&wempty[i], 1); }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, ncols);
        ptx::tmem_relinquish();
    }
    if (kMode == EPI_LN && warp >= 2) {
        for (int i = threadIdx.x - 64; i < bn32; i += 128 * kEpiGroups) {
            s_bias[i] = i < a.block_n ? __ldg(a.bias + n0 + i) : 0.f;
            s_wsum[i] = i < a.block_n ? __ldg(a.wsum + n0 + i) : 0.f;
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // resident weights that are CONSTANTS of the model are requested before the programmatic-dependency wait: they land
    // while the previous kernel of the stream is still finishing.  The attention's folded weights (written by
    // cab_fold_kernel a moment ago; resident when the batch is 1) must wait like every activation.
    const uint32_t b_tail = (uint32_t)a.block_n * (uint32_t)a.tailc * 2u;        // bytes of a tail weight chunk (tailc != 0)
    auto load_resident_weights = [&]() {
        const uint32_t ntail = a.tailc ? (uint32_t)a.taps : 0u;                    // one tail chunk per tap
        ptx::mbar_expect_tx(bfull, ((uint32_t)nbchunks - ntail) * b_chunk + ntail * b_tail);
        for (int i = 0; i < nbchunks; ++i) {
            if (i < k1) {
                const bool tail = a.tailc && (i % a.kchunks) == a.kchunks - 1;
                ptx::tma_load_3d(smBres + (size_t)i * b_chunk, tail ? &a.tmB_t : &a.tmB, bfull, i * 64, n0, 0);
            } else {
                ptx::tma_load_3d(smBres + (size_t)i * b_chunk, &a.tmB2, bfull, (i - k1) * 64, n0, 0);
            }
        }
    };
    if (warp == 0 && lane == 0 && a.b_resident && a.w_early) load_resident_weights();
    const TileWalk tw_first(blockIdx.x, gridDim.x, a.tiles_x, a.tiles_y);     // (its four divisions run under the predecessor's tail)
    ptx::pdl_wait();
    ptx::pdl_trigger();
    if (warp == 0 && lane == 0 && a.b_resident && !a.w_early) load_resident_weights();

    if (warp == 0) {
        // ------------------------------------------------------- TMA producer
        if (lane == 0) {
            RingPos rp, wp;
            for (TileWalk tw = tw_first; tw.t < a.num_tiles; tw.next()) {
                const int img = tw.img;
                const int y0 = tw.ty * a.TH;
                const int x0 = tw.tx * a.TW;
                int tap_i = 0, kc_i = 0;                       // (tap, channel chunk) of iteration i in the tap-shifted mode
                for (int i = 0; i < kiters; ++i, rp.advance((uint32_t)stages)) {
                    const uint32_t s = rp.s;
                    ptx::mbar_wait(&empty[s], rp.ph ^ 1u);
                    const bool up_here = (kUp) && i == 0;      // the tile's low-res box rides with its first stage
                    const bool a_tail = a.tailc && kHaloOn && i == a.kchunks - 1;
                    const uint32_t a_bytes = (kHaloOn || i < k1) ? kSub * (kHaloOn ? 11u * 16u : 128u) * (uint32_t)(a_tail ? a.tailc : a.boxc) * 2u
                                                                : 128u * (uint32_t)a.boxc2 * 2u;
                    const CUtensorMap* mapA = a_tail ? &a.tmA_t : &a.tmA;
                    ptx::mbar_expect_tx(&full[s], a_bytes + (up_here ? a.up_chunks * (uint32_t)(kUpBoxW * kUpBoxH) * (uint32_t)a.boxc_up * 2u : 0u) +
                                                      ((a.b_resident || kWRing) ? 0u : b_chunk));
                    uint8_t* dstA = smA + (size_t)s * a_stage;
                    if (up_here) {
                        const int by0 = (int)(a.up_ry * (float)(y0 + a.up_row0)) - a.up_src_row0;
                        const int bx0 = (int)(a.up_rx * (float)x0);
                        for (int ch = 0; ch < a.up_chunks; ++ch)
                            ptx::tma_load_4d(dstA + kSubTileBytes + ch * kUpTileBytes, &a.tmUp, &full[s], ch * 64, bx0, by0, img);
                    }
                    if (kHaloOn) {
                        // i == channel chunk; one (DOWN: two, even / odd rows) halo box serves all 9 taps
                        if (kMode == EPI_DOWN) {
                            ptx::tma_load_5d(dstA, mapA, &full[s], i * 64, x0 - 1, 0, y0 - 1, img);
                            ptx::tma_load_5d(dstA + kHaloTileBytes, mapA, &full[s], i * 64, x0 - 1, 1, y0 - 1, img);
                        } else {
                            ptx::tma_load_4d(dstA, mapA, &full[s], i * 64, x0 - 1, y0 - 1, img);
                        }
                        if (kWRing) {
                            // the 9 taps' weight chunks of this channel chunk follow through their own ring: the halo tile is
                            // loaded ONCE per tile and channel chunk (not once per tap, as the tap-shifted tiles were)
                            for (int tap = 0; tap < 9; ++tap, wp.advance((uint32_t)a.wring)) {
                                const uint32_t r = wp.s;
                                ptx::mbar_wait(&wempty[r], wp.ph ^ 1u);
                                ptx::mbar_expect_tx(&wfull[r], a_tail ? b_tail : b_chunk);
                                ptx::tma_load_3d(smB + (size_t)r * b_chunk, a_tail ? &a.tmB_t : &a.tmB, &wfull[r], (tap * a.kchunks + i) * 64, n0, 0);
                            }
                        }
                    } else if (i < k1) {
                        const int tap = tap_i, kc = kc_i;
                        if (++kc_i == a.kchunks) { kc_i = 0; ++tap_i; }
                        int dy = 1, dx = 1;
                        if (a.taps == 9) { dy = (tap * 11) >> 5; dx = tap - dy * 3; }      // tap / 3 for tap < 9
                        if (kMode == EPI_DOWN) {
#pragma unroll
                            for (int sub = 0; sub < 2; ++sub) {
                                const int srow = 2 * y0 + sub + dy - 1;     // first image row of this sub-tile
                                ptx::tma_load_5d(dstA + sub * kSubTileBytes, &a.tmA, &full[s],
                                                 kc * 64, x0 + dx - 1, srow & 1, srow >> 1, img);
                            }
                        } else {
                            ptx::tma_load_4d(dstA, &a.tmA, &full[s], kc * 64, x0 + dx - 1, y0 + dy - 1, img);
                        }
                        if (!a.b_resident)
                            ptx::tma_load_3d(smB + (size_t)s * b_chunk, &a.tmB, &full[s], i * 64, n0,
                                             a.per_image_w ? img : 0);
                    } else {
                        const int kc = i - k1;
                        ptx::tma_load_4d(dstA, &a.tmA2, &full[s], kc * 64, x0, y0, img);
                        if (!a.b_resident)
                            ptx::tma_load_3d(smB + (size_t)s * b_chunk, &a.tmB2, &full[s], kc * 64, n0, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // --------------------------------------------------------- MMA issuer
        const uint32_t idesc = ptx::umma_idesc_f16(CIDNET_UMMA_FMT, (uint32_t)a.block_n);
        if (a.b_resident) { ptx::mbar_wait(bfull, 0); }
        const uint32_t leader = ptx::elect_one() ? 1u : 0u;     // the one lane whose tcgen05.mma / commit take effect
        uint32_t j = 0;
        RingPos rp, wp;
        for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++j) {
            const uint32_t buf = j % kEpiGroups;
            ptx::mbar_wait(&tmem_empty[buf], ((j / kEpiGroups) & 1u) ^ 1u);      // epilogue drained this buffer
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * acc_cols;
            int kc_i = 0;                                                        // channel chunk of iteration i (tap-shifted mode)
            for (int i = 0; i < kiters; ++i, rp.advance((uint32_t)stages)) {
                const uint32_t s = rp.s;
                ptx::mbar_wait(&full[s], rp.ph);
                ptx::tc_fence_after();
                if ((kMode == EPI_LN || kUp) && i == kiters - 1 && lane == 0) ptx::mbar_arrive(&a_ready[buf]);
                if (kHaloOn) {
                    // One (elected) lane issues every tcgen05.mma, the whole warp runs the loop convergently; ncu showed the tensor pipe busy only ~55 % of the time
                    // behind this loop (~150 cycles of uniform-datapath descriptor arithmetic per MMA), so the
                    // loops are fully unrolled: tap / sub-tile offsets are compile-time constants added to two base
                    // descriptors, the accumulate flag is an immediate and only `k < ksteps` stays a uniform branch.
                    int ksteps = (a.cin - i * 64 + 15) >> 4;
                    if (ksteps > 4) ksteps = 4;
                    const uint64_t dA0 = ptx::umma_smem_desc_sw128(ptx::smem_u32(smA + (size_t)s * a_stage));
                    uint64_t descB = ptx::umma_smem_desc_sw128(ptx::smem_u32(smBres + (size_t)i * b_chunk));
                    const uint64_t b_tap = (uint64_t)(((uint32_t)a.kchunks * b_chunk) >> 4);   // next tap's weights
                    const uint32_t d0 = d_tmem, d1 = d_tmem + (uint32_t)a.block_n;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dy = tap / 3, dx = tap - dy * 3;
                        uint32_t wr = 0;
                        if (kWRing) {                      // streamed weights: this tap's chunk comes out of the weight ring
                            wr = wp.s;
                            ptx::mbar_wait(&wfull[wr], wp.ph);
                            ptx::tc_fence_after();
                            descB = ptx::umma_smem_desc_sw128(ptx::smem_u32(smB + (size_t)wr * b_chunk));
                            wp.advance((uint32_t)a.wring);
                        }
#pragma unroll
                        for (int sub = 0; sub < kSub; ++sub) {
                            // which halo tile and which first row this (accumulator, tap) reads
                            uint32_t tile = 0, rowoff = (uint32_t)(dy * 16 + dx);
                            if (kMode == EPI_DOWN) {
                                // tiles hold even (0) / odd (1) image rows of row pairs [y0-1, y0+9]
                                if (sub == 0) { tile = (dy == 1) ? 0u : 1u; rowoff = (uint32_t)((dy == 0 ? 0 : 16) + dx); }
                                else          { tile = (dy == 1) ? 1u : 0u; rowoff = (uint32_t)((dy == 2 ? 32 : 16) + dx); }
                            }
                            const uint64_t descA = dA0 + (uint64_t)((tile * kHaloTileBytes + rowoff * 128u) >> 4);
                            const uint32_t dt = sub == 0 ? d0 : d1;
                            if (tap == 0) {
                                if (i == 0) ptx::umma_f16_lead<false>(leader, dt, descA, descB, idesc);
                                else        ptx::umma_f16_lead<true>(leader, dt, descA, descB, idesc);
                            } else {
                                ptx::umma_f16_lead<true>(leader, dt, descA, descB, idesc);
                            }
                            if (ksteps > 1) ptx::umma_f16_lead<true>(leader, dt, descA + 2, descB + 2, idesc);
                            if (ksteps > 2) ptx::umma_f16_lead<true>(leader, dt, descA + 4, descB + 4, idesc);
                            if (ksteps > 3) ptx::umma_f16_lead<true>(leader, dt, descA + 6, descB + 6, idesc);
                        }
                        if (kWRing) ptx::umma_commit_lead(leader, &wempty[wr]);     // frees the weight stage when its MMAs retire
                        else descB += b_tap;
                    }
                    ptx::umma_commit_lead(leader, &empty[s]);
                    if (i == kiters - 1) ptx::umma_commit_lead(leader, &tmem_full[buf]);
                } else {
                    int crem = (i < k1) ? a.cin - kc_i * 64 : a.cin2 - (i - k1) * 64;
                    if (++kc_i == a.kchunks) kc_i = 0;
                    int ksteps = (crem + 15) >> 4;
                    if (ksteps > 4) ksteps = 4;
                    const uint8_t* bsrc = a.b_resident ? smBres + (size_t)i * b_chunk : smB + (size_t)s * b_chunk;
                    const uint64_t descB = ptx::umma_smem_desc_sw128(ptx::smem_u32(bsrc));
#pragma unroll
                    for (int sub = 0; sub < kSub; ++sub) {
                        const uint64_t descA = ptx::umma_smem_desc_sw128(
                            ptx::smem_u32(smA + (size_t)s * a_stage + sub * kSubTileBytes));
                        const uint32_t dt = d_tmem + sub * a.block_n;
                        if (i == 0) ptx::umma_f16_lead<false>(leader, dt, descA, descB, idesc);
                        else        ptx::umma_f16_lead<true>(leader, dt, descA, descB, idesc);
                        if (ksteps > 1) ptx::umma_f16_lead<true>(leader, dt, descA + 2, descB + 2, idesc);
                        if (ksteps > 2) ptx::umma_f16_lead<true>(leader, dt, descA + 4, descB + 4, idesc);
                        if (ksteps > 3) ptx::umma_f16_lead<true>(leader, dt, descA + 6, descB + 6, idesc);
                    }
                    ptx::umma_commit_lead(leader, &empty[s]);                           // frees the smem stage when the MMAs retire
                    if (i == kiters - 1) ptx::umma_commit_lead(leader, &tmem_full[buf]); // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ----------------------------------------------------------- epilogue
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;        // epilogue group == TMEM buffer it drains
        const int row = q * 32 + lane;          // accumulator row == pixel within the tile
        const int pw = kHaloOn ? 16 : a.TW;      // halo mode: rows are 16 wide, the last 2 columns are wrap-around garbage
        const int ty = row / pw;
        const int tx = row - ty * pw;
        const bool col_ok = tx < a.TW;
        const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
        uint8_t* const stg_base = smOut + (size_t)grp * a.stg_bufs * kStagingBytes;
        const uint32_t sb_toggle = a.stg_bufs - 1;
        const uint32_t bar_id = 1 + grp;
        const int nblk64 = (min(a.block_n, a.n_out - n0) + 63) >> 6;     // 64-channel output blocks
        uint32_t j = 0, sb = 0;
        RingPos rp;                                   // ring position of the current tile's FIRST pipeline stage (LN / UP read it)
        for (TileWalk tw = tw_first; tw.t < a.num_tiles;
             tw.next(), ++j, rp.advance_by((uint32_t)kiters, (uint32_t)stages)) {
            if ((int)(j % kEpiGroups) != grp) continue;                        // another group's tile
            const int img = tw.img;
            const int y0 = tw.ty * a.TH;
            const int x0 = tw.tx * a.TW;
            const int y = y0 + ty, x = x0 + tx;
            const bool valid = col_ok && (y < a.Hv) && (x < a.Wv);

            float mean = 0.f, rstd = 1.f;
            if (kMode == EPI_LN) {
                // per-pixel LayerNorm statistics from the A tile (all K chunks of this tile are resident:
                // the host guarantees stages >= kchunks + 1); then co-release the stages
                ptx::mbar_wait(&a_ready[j % kEpiGroups], (j / kEpiGroups) & 1u);
                // one pass: sum and sum of squares with FHFMA (x*1 and x*x, exact 16-bit products, fp32
                // accumulation; channels >= Cin are TMA zero fill and contribute nothing)
                float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                uint32_t st = rp.s;
                for (int kc = 0; kc < a.kchunks; ++kc, st = (st + 1 == (uint32_t)stages) ? 0u : st + 1) {
                    const uint8_t* rowp = smA + (size_t)st * a_stage + row * 128;
                    const int nvec = min(8, (a.cin - kc * 64 + 7) >> 3);
                    for (int jv = 0; jv < nvec; ++jv) {
                        const uint4 raw = *reinterpret_cast<const uint4*>(rowp + ((jv ^ (row & 7)) << 4));
                        fhfma2(s0, s1, raw.x, CIDNET_ONE_ONE); fhfma2(q0, q1, raw.x, raw.x);
                        fhfma2(s0, s1, raw.y, CIDNET_ONE_ONE); fhfma2(q0, q1, raw.y, raw.y);
                        fhfma2(s0, s1, raw.z, CIDNET_ONE_ONE); fhfma2(q0, q1, raw.z, raw.z);
                        fhfma2(s0, s1, raw.w, CIDNET_ONE_ONE); fhfma2(q0, q1, raw.w, raw.w);
                    }
                }
                const float inv_c = 1.0f / (float)a.cin;
                mean = (s0 + s1) * inv_c;
                const float var = fmaxf((q0 + q1) * inv_c - mean * mean, 0.f);
                rstd = 1.0f / sqrtf(var + a.ln_eps);
                __syncwarp();
                if (lane == 0) {
                    uint32_t sr = rp.s;
                    for (int kc = 0; kc < a.kchunks; ++kc, sr = (sr + 1 == (uint32_t)stages) ? 0u : sr + 1) ptx::mbar_arrive(&empty[sr]);
                }
            }

            // EPI_UP: bilinear x2 taps of the low-res tensor (align_corners=True), read from the box the producer
            // staged with this tile's first pipeline stage (rows of 64 channels, SWIZZLE_128B)
            const uint8_t* up_tile = nullptr;
            uint32_t u00 = 0, u01 = 0, u10 = 0, u11 = 0;      // box row index (128-byte rows) of the four neighbours
            float uly = 0.f, ulx = 0.f;
            uint32_t uw00 = 0, uw01 = 0, uw10 = 0, uw11 = 0;  // the four bilinear weights as packed (w, w) 16-bit pairs
            if (kUp) {
                const uint32_t s0 = rp.s;
                ptx::mbar_wait(&a_ready[j % kEpiGroups], (j / kEpiGroups) & 1u);
                up_tile = smA + (size_t)s0 * a_stage + kSubTileBytes;
                const int yr = y, xr = x;                                    // UP tiles are 8 x 16 rectangles of one image
                const float sy = a.up_ry * (float)(yr + a.up_row0);          // GLOBAL source row
                const int i0g = (int)sy; uly = sy - (float)i0g;
                const int i1g = i0g + (i0g < a.up_Hg - 1 ? 1 : 0);
                const float sx = a.up_rx * (float)xr;
                const int j0 = (int)sx; ulx = sx - (float)j0;
                const int j1 = j0 + (j0 < a.up_W - 1 ? 1 : 0);
                // box origin: the same arithmetic as the producer's.  Rows / columns clamp into the box: only pixels
                // outside the image (never stored) or invalid halo rows of a row strip can fall outside it
                const int by0 = (int)(a.up_ry * (float)(y0 + a.up_row0));
                const int bx0 = (int)(a.up_rx * (float)x0);
                const int r0 = min(max(i0g - by0, 0), kUpBoxH - 1), r1 = min(max(i1g - by0, 0), kUpBoxH - 1);
                const int c0 = min(max(j0 - bx0, 0), kUpBoxW - 1), c1 = min(max(j1 - bx0, 0), kUpBoxW - 1);
                u00 = r0 * kUpBoxW + c0; u01 = r0 * kUpBoxW + c1; u10 = r1 * kUpBoxW + c0; u11 = r1 * kUpBoxW + c1;
#ifndef CIDNET_ACT_BF16
                const float w11 = uly * ulx, w10 = uly - w11, w01 = ulx - w11, w00 = (1.f - uly) - w01;
                uw00 = pack2(w00, w00); uw01 = pack2(w01, w01); uw10 = pack2(w10, w10); uw11 = pack2(w11, w11);
#endif
            }
            // EPI_DOWN geometry: y counts ROW PAIRS == output rows; even lanes own an output pixel
            float dly = 0.f, dlx = 0.f;
            bool top_is_odd = false, left_is_right = false;
            if (kMode == EPI_DOWN) {
                const int oy = y + a.down_row0, ox = x >> 1;                 // GLOBAL output row
                const float sy = a.down_ry * (float)oy;
                const int i0 = (int)sy;
                dly = sy - (float)i0;
                top_is_odd = (i0 - 2 * oy) != 0;          // only at the clamped last row
                const float sx = a.down_rx * (float)ox;
                const int j0 = (int)sx;
                dlx = sx - (float)j0;
                left_is_right = (j0 - 2 * ox) != 0;       // only at the clamped last column
            }

            const uint32_t buf = j % kEpiGroups;
            ptx::mbar_wait(&tmem_full[buf], (j / kEpiGroups) & 1u);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + buf * acc_cols + ((uint32_t)(q * 32) << 16);

            float sa_sum = 0.f, sa_max = -INFINITY;               // EPI_UP + sa_stats: channel statistics of this thread's pixel
            for (int cb = 0; cb < nblk64; ++cb, sb ^= sb_toggle) {
                uint8_t* stg = stg_base + sb * kStagingBytes;
                // the TMA store that used this staging buffer two blocks ago must have read it
                if (issuer) { if (a.stg_bufs == 2) ptx::tma_store_wait_read<1>(); else ptx::tma_store_wait_read<0>(); }
                asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int c = cb * 64 + hh * 32;
                    if (c >= a.block_n) break;                      // warp-uniform
                    float v[32];
                    if (kMode == EPI_DOWN) {
                        float o[32];
                        ptx::tmem_ld32(taddr + c, v);
                        ptx::tmem_ld32(taddr + a.block_n + c, o);
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            const float top = top_is_odd ? o[e] : v[e];
                            const float vv = (1.f - dly) * top + dly * o[e];
                            const float vr = __shfl_down_sync(0xffffffffu, vv, 1);
                            const float left = left_is_right ? vr : vv;
                            v[e] = prelu_f((1.f - dlx) * left + dlx * vr, a.prelu);
                        }
                    } else {
                        ptx::tmem_ld32(taddr + c, v);
                        if (kMode == EPI_LN) {
                            const float nm = -mean * rstd;                  // rstd*(v - mean*ws) + b = rstd*v + nm*ws + b
#pragma unroll
                            for (int e4 = 0; e4 < 8; ++e4) {
                                const float4 ws = *reinterpret_cast<const float4*>(s_wsum + c + 4 * e4);
                                const float4 bs = *reinterpret_cast<const float4*>(s_bias + c + 4 * e4);
                                v[4 * e4 + 0] = fmaf(rstd, v[4 * e4 + 0], fmaf(nm, ws.x, bs.x));
                                v[4 * e4 + 1] = fmaf(rstd, v[4 * e4 + 1], fmaf(nm, ws.y, bs.y));
                                v[4 * e4 + 2] = fmaf(rstd, v[4 * e4 + 2], fmaf(nm, ws.z, bs.z));
                                v[4 * e4 + 3] = fmaf(rstd, v[4 * e4 + 3], fmaf(nm, ws.w, bs.w));
                            }
                        } else if (kMode == EPI_STORE) {
                            if (a.use_prelu) {
#pragma unroll
                                for (int e = 0; e < 32; ++e) v[e] = prelu_f(v[e], a.prelu);
                            }
                        } else if (kUp) {
                            const int nrem = a.n_out - (n0 + c);
                            const uint8_t* ub = up_tile + cb * kUpTileBytes;
#pragma unroll
                            for (int h = 0; h < 4; ++h) {
                                if (nrem > 8 * h) {                            // warp-uniform
                                    const uint32_t ck = (uint32_t)(hh * 4 + h);   // 16-byte chunk of the 128-byte row
#ifndef CIDNET_ACT_BF16
                                    // The four neighbours stay PACKED: acc += w_tap * p_tap as mixed-precision FMAs (FHFMA: exact
                                    // fp16 x fp16 product, fp32 accumulation straight into the GEMM accumulator) -- 4 instructions
                                    // per channel instead of 4 conversions + 6 fp32 FMAs.  ncu on the fp32 version: 24.5 M warp
                                    // instructions per up1 launch (1094 per pixel) on the 8 epilogue warps, issue slots 52 % busy with
                                    // two warps per scheduler: the layer was bound by this loop, not by HBM (22 %).  The bilinear
                                    // weights are rounded to fp16 (relative 2^-11, the size of the storage rounding of the result).
                                    const uint4 q00 = *reinterpret_cast<const uint4*>(ub + u00 * 128 + ((ck ^ (u00 & 7)) << 4));
                                    const uint4 q01 = *reinterpret_cast<const uint4*>(ub + u01 * 128 + ((ck ^ (u01 & 7)) << 4));
                                    const uint4 q10 = *reinterpret_cast<const uint4*>(ub + u10 * 128 + ((ck ^ (u10 & 7)) << 4));
                                    const uint4 q11 = *reinterpret_cast<const uint4*>(ub + u11 * 128 + ((ck ^ (u11 & 7)) << 4));
                                    float* vv = v + 8 * h;
                                    fhfma2(vv[0], vv[1], q00.x, uw00); fhfma2(vv[2], vv[3], q00.y, uw00);
                                    fhfma2(vv[4], vv[5], q00.z, uw00); fhfma2(vv[6], vv[7], q00.w, uw00);
                                    fhfma2(vv[0], vv[1], q01.x, uw01); fhfma2(vv[2], vv[3], q01.y, uw01);
                                    fhfma2(vv[4], vv[5], q01.z, uw01); fhfma2(vv[6], vv[7], q01.w, uw01);
                                    fhfma2(vv[0], vv[1], q10.x, uw10); fhfma2(vv[2], vv[3], q10.y, uw10);
                                    fhfma2(vv[4], vv[5], q10.z, uw10); fhfma2(vv[6], vv[7], q10.w, uw10);
                                    fhfma2(vv[0], vv[1], q11.x, uw11); fhfma2(vv[2], vv[3], q11.y, uw11);
                                    fhfma2(vv[4], vv[5], q11.z, uw11); fhfma2(vv[6], vv[7], q11.w, uw11);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        vv[e] = prelu_f(vv[e], a.prelu);
                                        if (kSa && 8 * h + e < nrem) {
                                            const float r16 = act2f(f2act(vv[e]));             // the value as it is stored
                                            sa_sum += r16; sa_max = fmaxf(sa_max, r16);
                                        }
                                    }
#else
                                    float p00[8], p01[8], p10[8], p11[8];
                                    load8(reinterpret_cast<const act_t*>(ub + u00 * 128 + ((ck ^ (u00 & 7)) << 4)), p00);
                                    load8(reinterpret_cast<const act_t*>(ub + u01 * 128 + ((ck ^ (u01 & 7)) << 4)), p01);
                                    load8(reinterpret_cast<const act_t*>(ub + u10 * 128 + ((ck ^ (u10 & 7)) << 4)), p10);
                                    load8(reinterpret_cast<const act_t*>(ub + u11 * 128 + ((ck ^ (u11 & 7)) << 4)), p11);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float up = (1.f - uly) * ((1.f - ulx) * p00[e] + ulx * p01[e]) +
                                                         uly * ((1.f - ulx) * p10[e] + ulx * p11[e]);
                                        v[8 * h + e] = prelu_f(v[8 * h + e] + up, a.prelu);
                                        if (kSa && 8 * h + e < nrem) {
                                            const float r16 = act2f(f2act(v[8 * h + e]));      // the value as it is stored
                                            sa_sum += r16; sa_max = fmaxf(sa_max, r16);
                                        }
                                    }
#endif
                                }
                            }
                        }
                    }
                    // 16-bit, swizzled (SWIZZLE_128B) staging row; DOWN: only even lanes own an output pixel
                    int srow = kHaloOn ? ty * a.TW + tx : row;      // halo mode: drop the 2 garbage columns
                    bool wr = col_ok;
                    if (kMode == EPI_DOWN) { srow = ty * (a.TW >> 1) + (tx >> 1); wr = col_ok && (tx & 1) == 0; }
                    if (wr) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int chunk = hh * 4 + g;       // 16-byte chunk within the 128-byte row
                            *reinterpret_cast<uint4*>(stg + srow * 128 + ((chunk ^ (srow & 7)) << 4)) = pack8(v + 8 * g);
                        }
                    }
                }
                ptx::fence_proxy_async_smem();
                asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
                if (issuer) {
                    const int c0 = n0 + cb * 64;
                    if (kMode == EPI_DOWN) ptx::tma_store_4d(&a.tmOut, stg, c0, x0 >> 1, y0, img);
                    else                   ptx::tma_store_4d(&a.tmOut, stg, c0, x0, y0, img);
                    ptx::tma_store_commit();
                }
            }
            if (kSa && valid)
                a.sa_stats[((long long)img * a.Hv + y) * a.Wv + x] = make_float2(sa_sum / (float)a.n_out, sa_max);
            // all TMEM reads of this buffer are done -> hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&tmem_empty[buf]);
                if (kUp) {    // the low-res box has been read: co-release the tile's stages
                    uint32_t sr = rp.s;
                    for (int i = 0; i < kiters; ++i, sr = (sr + 1 == (uint32_t)stages) ? 0u : sr + 1) ptx::mbar_arrive(&empty[sr]);
                }
            }
        }
        if (issuer) ptx::tma_store_wait_read<0>();      // the staging buffers have been READ; the global writes complete with the grid
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------- tensor map encode ---
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    return fn;
}

static int encode_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes = 128) {
    EncodeTiledFn fn = get_encode_fn();
    CIDNET_CHECK(fn != nullptr, CIDNET_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
#ifdef CIDNET_ACT_BF16
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
#else
    const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
#endif
    CUresult r = fn(m, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[256];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) rank=%d dims=%llu,%llu,%llu stride0=%llu", (int)r,
                 rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                 (unsigned long long)strides_bytes[0]);
        return fail(CIDNET_ERR_CUDA, buf);
    }
    return CIDNET_OK;
}

int encode_map_generic(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box) {
    return encode_map(m, base, rank, dims, strides_bytes, box);
}

int encode_map_generic_swz(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, int swizzle_bytes) {
    return encode_map(m, base, rank, dims, strides_bytes, box, swizzle_bytes);
}

static inline float ac_scale(int n_in, int n_out) {
    // torch area_pixel_compute_scale<float>(align_corners=True)
    return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
}

template <int kMode, int kEpiGroups, int kHalo>
static int launch_mode_g(const ConvGemmArgs& args, dim3 grid, size_t smem, cudaStream_t stream) {
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(conv_gemm_kernel<kMode, kEpiGroups, kHalo>), 227 * 1024);
    if (rc) return rc;
    return launch_k(conv_gemm_kernel<kMode, kEpiGroups, kHalo>, grid, dim3(64 + 128 * kEpiGroups), smem, stream, args);
}
template <int kMode>
static int launch_mode(const ConvGemmArgs& args, dim3 grid, size_t smem, cudaStream_t stream, int groups) {
    const int hm = args.halo ? (args.wring ? 2 : 1) : 0;
#ifdef CIDNET_GEMM_GROUPS3
    if (groups == 3 && hm == 0) return launch_mode_g<kMode, 3, 0>(args, grid, smem, stream);
#endif
    (void)groups;
    if constexpr (kMode == EPI_STORE || kMode == EPI_DOWN) {      // the only modes with 3x3 layers
        if (hm == 1) return launch_mode_g<kMode, 2, 1>(args, grid, smem, stream);
        if (hm == 2) return launch_mode_g<kMode, 2, 2>(args, grid, smem, stream);
    } else {
        CIDNET_CHECK(hm == 0, CIDNET_ERR_INVALID, "conv_gemm: halo tiles are built for the STORE / DOWN epilogues only");
    }
    return launch_mode_g<kMode, 2, 0>(args, grid, smem, stream);
}

int launch_conv_gemm(const ConvGemmLaunch& L, cudaStream_t stream) {
    const PackedWeights& wt = *L.wt;
    CIDNET_CHECK(L.in && L.out && wt.w, CIDNET_ERR_INVALID, "conv_gemm: null pointer");
    CIDNET_CHECK(L.B > 0 && L.H > 0 && L.W > 0, CIDNET_ERR_INVALID, "conv_gemm: empty problem");
    CIDNET_CHECK(wt.block_n % 16 == 0 && wt.block_n >= 16 && wt.block_n <= 256, CIDNET_ERR_INVALID,
                 "conv_gemm: block_n must be a multiple of 16 in [16,256]");
    CIDNET_CHECK(L.in_pitch % 8 == 0 && L.out_pitch % 8 == 0, CIDNET_ERR_INVALID, "conv_gemm: pitch % 8");
    CIDNET_CHECK(wt.taps == 1 || wt.taps == 9, CIDNET_ERR_INVALID, "conv_gemm: taps must be 1 or 9");
    CIDNET_CHECK(!(L.flat && wt.taps != 1), CIDNET_ERR_INVALID, "conv_gemm: flat tiling is for 1x1 only");
    const int ksub = L.mode == EPI_DOWN ? 2 : 1;
    CIDNET_CHECK(2 * ksub * wt.block_n <= 512, CIDNET_ERR_INVALID, "conv_gemm: accumulators exceed TMEM");
    // epilogue groups: 2.  A third group for the 1x1 layers was measured on B200 in round 1: no layer got faster at cfg 2
    // or 16x400x600, and the 400x600 forward lost parity (2.9e-2).  Root cause, found in round 2 with the
    // `build.py --groups3` experiment (profiles/r02_summary.md): (1) TMEM was sized as the next power of two >= 3 * block_n
    // columns without checking the 512 an SM has (guard below); (2) the LN / UP epilogues waited on the ring's `full`
    // barriers by parity although an epilogue group does not see every phase of a stage -- with 2 tiles in the ring and
    // 3 groups the wait could succeed on the previous phase (see `a_ready` in the kernel).  The same hole was latent in
    // the shipped 2-group kernels whenever shared memory limited the ring to fewer than 2 tiles; the a_ready barrier closes
    // it for every configuration.  3 groups are still not faster, so the library instantiates the 2-group kernels only.
    int kEpiGroups = 2;
#ifdef CIDNET_GEMM_GROUPS3
    if (wt.taps == 1 && L.mode != EPI_DOWN && 3 * wt.block_n <= 512) kEpiGroups = 3;
#endif

    ConvGemmArgs a;
    memset(&a, 0, sizeof a);
    a.taps = wt.taps; a.kchunks = wt.kchunks; a.cin = wt.cin;
    a.n_out = wt.n_out; a.block_n = wt.block_n; a.per_image_w = wt.n_img > 1 ? 1 : 0;
    a.bias = wt.bias; a.wsum = wt.wsum; a.ln_eps = L.ln_eps;
    a.prelu = L.prelu; a.use_prelu = L.use_prelu ? 1 : 0;
    a.w_real = L.W;
    a.w_early = L.dynamic_weights ? 0 : 1;
    a.sa_stats = L.mode == EPI_UP ? L.sa_stats : nullptr;
    const EpiMode kmode = (L.mode == EPI_UP && L.sa_stats) ? EPI_UP_SA : L.mode;
    const long long hw = (long long)L.H * L.W;
    int rc;

    // 3x3 halo mode: possible when all 9 taps' weights stay resident next to >= 2 halo stages
    {
        // (measured on B200: the SWIZZLE_128B XOR is applied to ABSOLUTE shared-memory address bits, so a
        // descriptor that starts on an arbitrary 128-byte row of a 1024-byte aligned TMA tile reads the
        // right data with base_offset = 0; setting (addr >> 7) & 7 there gives wrong results)
        const size_t fixed1 = 1024 + (size_t)kEpiGroups * kStagingBytes + 512 + 2 * round_up(wt.block_n, 32) * sizeof(float);
        const size_t need = fixed1 + (size_t)9 * wt.kchunks * wt.block_n * 128 + (size_t)2 * ksub * kHaloTileBytes;
        const bool halo_mode_ok = L.mode == EPI_STORE || L.mode == EPI_DOWN;
        a.halo = (halo_mode_ok && wt.taps == 9 && !L.in2 && !a.per_image_w && need <= 226 * 1024) ? 1 : 0;
        // weights too large to stay resident (72 -> 144, 144 -> 72): keep the halo tile anyway and STREAM the weight chunks
        // through a ring of their own.  The tap-shifted mode re-loaded the A tile for every tap: 18 x (32 KB + 9 KB) = 742 KB
        // of TMA traffic per 256-pixel tile of down3 (x 2 N blocks), which at ~25 B / cycle / SM IS the layer's time
        // (38.7 us per launch at cfg 2, 277 TFLOP/s at cfg 5); with the halo tile resident it is 2 x 45 KB + 18 x 9 KB.
        const size_t need_ring = fixed1 + (size_t)2 * ksub * kHaloTileBytes + (size_t)4 * wt.block_n * 128;
        if (halo_mode_ok && !a.halo && wt.taps == 9 && !L.in2 && !a.per_image_w && need_ring <= 226 * 1024) { a.halo = 1; a.wring = 4; }
    }
    {   // tail box: halo mode, several channel chunks, the last one at most a quarter full
        const int rem = wt.cin - (wt.kchunks - 1) * 64;
        a.tailc = (a.halo && wt.kchunks > 1 && rem <= 16) ? 16 : 0;
    }
    auto box_channels = [](int c, int chunks) { return chunks == 1 ? (c <= 48 ? (unsigned)round_up(c, 16) : 64u) : 64u; };
    a.boxc = (int)box_channels(wt.cin, wt.kchunks);
    a.boxc2 = L.in2 && L.wt2 ? (int)box_channels(L.wt2->cin, L.wt2->kchunks) : 64;
    a.boxc_out = wt.n_out <= 48 ? round_up(wt.n_out, 16) : 64;
    a.boxc_up = a.boxc_out;
    const int tw = a.halo ? 14 : 16;                            // halo mode: 16-wide smem rows, 14 valid columns
    const uint64_t pb = (uint64_t)L.in_pitch * sizeof(act_t);   // bytes per pixel row
    const uint64_t ob = (uint64_t)L.out_pitch * sizeof(act_t);
    if (L.mode == EPI_DOWN) {
        CIDNET_CHECK(L.H % 2 == 0 && L.W % 2 == 0 && wt.taps == 9, CIDNET_ERR_INVALID, "conv_gemm: DOWN needs even H,W, 3x3");
        a.Hv = L.H / 2; a.Wv = L.W; a.TH = 8; a.TW = tw;
        a.in_H = L.H; a.in_W = L.W;
        const int gH = L.gH ? L.gH : L.H;
        CIDNET_CHECK(gH % 2 == 0 && L.grow % 2 == 0 && L.grow + L.H <= gH, CIDNET_ERR_INVALID, "conv_gemm: bad row strip");
        a.down_ry = ac_scale(gH, gH / 2); a.down_rx = ac_scale(L.W, L.W / 2);
        a.down_row0 = L.grow / 2;
        const uint64_t dims[5] = {(uint64_t)wt.cin, (uint64_t)L.W, 2, (uint64_t)L.H / 2, (uint64_t)L.B};
        const uint64_t str[4] = {pb, pb * L.W, pb * L.W * 2, pb * hw};
        const uint32_t box[5] = {(uint32_t)a.boxc, 16, 1, (uint32_t)(a.halo ? 11 : 8), 1};
        if ((rc = encode_map(&a.tmA, L.in, 5, dims, str, box))) return rc;
        if (a.tailc) {
            const uint32_t tbox[5] = {(uint32_t)a.tailc, 16, 1, 11, 1};
            if ((rc = encode_map(&a.tmA_t, L.in, 5, dims, str, tbox))) return rc;
        }
        const uint64_t od[4] = {(uint64_t)wt.n_out, (uint64_t)L.W / 2, (uint64_t)L.H / 2, (uint64_t)L.B};
        const uint64_t os[3] = {ob, ob * (L.W / 2), ob * (hw / 4)};
        const uint32_t obox[4] = {(uint32_t)a.boxc_out, (uint32_t)(tw / 2), 8, 1};
        if ((rc = encode_map(&a.tmOut, L.out, 4, od, os, obox))) return rc;
    } else {
        if (L.flat && L.mode != EPI_UP) { a.Hv = 1; a.Wv = (int)hw; a.TH = 1; a.TW = 128; }   // UP tiles are always 8 x 16 rectangles
        else        { a.Hv = L.H; a.Wv = L.W; a.TH = 8; a.TW = tw; }
        const uint64_t dims[4] = {(uint64_t)wt.cin, (uint64_t)a.Wv, (uint64_t)a.Hv, (uint64_t)L.B};
        const uint64_t str[3] = {pb, pb * a.Wv, pb * hw};
        const uint32_t box[4] = {(uint32_t)a.boxc, (uint32_t)a.TW, (uint32_t)a.TH, 1};
        const uint32_t hbox[4] = {(uint32_t)a.boxc, 16, 11, 1};
        if ((rc = encode_map(&a.tmA, L.in, 4, dims, str, a.halo ? hbox : box))) return rc;
        if (a.tailc) {
            const uint32_t tbox[4] = {(uint32_t)a.tailc, 16, 11, 1};
            if ((rc = encode_map(&a.tmA_t, L.in, 4, dims, str, tbox))) return rc;
        }
        const uint64_t od[4] = {(uint64_t)wt.n_out, (uint64_t)a.Wv, (uint64_t)a.Hv, (uint64_t)L.B};
        const uint64_t os[3] = {ob, ob * a.Wv, ob * hw};
        const uint32_t obox[4] = {(uint32_t)a.boxc_out, (uint32_t)a.TW, (uint32_t)a.TH, 1};
        if ((rc = encode_map(&a.tmOut, L.out, 4, od, os, obox))) return rc;
        if (L.in2) {
            CIDNET_CHECK(L.wt2 && L.wt2->w && L.wt2->taps == 1 && L.wt2->block_n == wt.block_n &&
                             L.wt2->n_blocks == wt.n_blocks && L.wt2->n_img == 1 && L.in2_pitch % 8 == 0,
                         CIDNET_ERR_INVALID, "conv_gemm: second K source must be a 1x1 with the same N blocking");
            a.kchunks2 = L.wt2->kchunks; a.cin2 = L.wt2->cin;
            const uint64_t pb2 = (uint64_t)L.in2_pitch * sizeof(act_t);
            const uint64_t d2[4] = {(uint64_t)L.wt2->cin, (uint64_t)a.Wv, (uint64_t)a.Hv, (uint64_t)L.B};
            const uint64_t s2[3] = {pb2, pb2 * a.Wv, pb2 * hw};
            const uint32_t box2[4] = {(uint32_t)a.boxc2, (uint32_t)a.TW, (uint32_t)a.TH, 1};
            if ((rc = encode_map(&a.tmA2, L.in2, 4, d2, s2, box2))) return rc;
            const uint64_t kt2 = (uint64_t)L.wt2->ktot();
            const uint64_t bd[3] = {kt2, (uint64_t)L.wt2->n_rows, 1};
            const uint64_t bs[2] = {kt2 * sizeof(act_t), kt2 * sizeof(act_t) * L.wt2->n_rows};
            const uint32_t bbox[3] = {64, (uint32_t)wt.block_n, 1};
            if ((rc = encode_map(&a.tmB2, L.wt2->w, 3, bd, bs, bbox))) return rc;
        }
        if (L.mode == EPI_UP) {
            CIDNET_CHECK(L.up != nullptr && L.H % 2 == 0 && L.W % 2 == 0, CIDNET_ERR_INVALID, "conv_gemm: UP needs t");
            CIDNET_CHECK(wt.n_blocks == 1 && L.up_pitch % 8 == 0 && !L.in2, CIDNET_ERR_INVALID, "conv_gemm: UP needs a single N block");
            a.up_H = L.H / 2; a.up_W = L.W / 2; a.up_chunks = ceil_div(wt.n_out, 64);
            {   // low-res tensor t [B][H/2][W/2][up_pitch]: one {64 ch, 10, 6} box per output tile and 64 channels
                const uint64_t ub = (uint64_t)L.up_pitch * sizeof(act_t);
                const uint64_t ud[4] = {(uint64_t)wt.n_out, (uint64_t)a.up_W, (uint64_t)a.up_H, (uint64_t)L.B};
                const uint64_t us[3] = {ub, ub * a.up_W, ub * (uint64_t)a.up_W * a.up_H};
                const uint32_t ubox[4] = {(uint32_t)a.boxc_up, (uint32_t)kUpBoxW, (uint32_t)kUpBoxH, 1};
                if ((rc = encode_map(&a.tmUp, L.up, 4, ud, us, ubox))) return rc;
            }
            const int gH = L.gH ? L.gH : L.H;
            CIDNET_CHECK(gH % 2 == 0 && L.grow % 2 == 0 && L.grow + L.H <= gH, CIDNET_ERR_INVALID, "conv_gemm: bad row strip");
            a.up_ry = ac_scale(gH / 2, gH); a.up_rx = ac_scale(L.W / 2, L.W);
            a.up_row0 = L.grow; a.up_src_row0 = L.grow / 2; a.up_Hg = gH / 2;
        }
    }
    a.tiles_x = ceil_div(a.Wv, a.TW);
    a.tiles_y = ceil_div(a.Hv, a.TH);
    a.num_tiles = a.tiles_x * a.tiles_y * L.B;
    {
        const uint64_t kt = (uint64_t)wt.ktot();
        const uint64_t dims[3] = {kt, (uint64_t)wt.n_rows, (uint64_t)wt.n_img};
        const uint64_t str[2] = {kt * sizeof(act_t), kt * sizeof(act_t) * wt.n_rows};
        const uint32_t box[3] = {64, (uint32_t)wt.block_n, 1};
        if ((rc = encode_map(&a.tmB, wt.w, 3, dims, str, box))) return rc;
        if (a.tailc) {
            const uint32_t tbox[3] = {(uint32_t)a.tailc, (uint32_t)wt.block_n, 1};
            if ((rc = encode_map(&a.tmB_t, wt.w, 3, dims, str, tbox))) return rc;
        }
    }

    // shared-memory plan: [resident weights] [A ring] [streamed-weight ring] [2 staging] [barriers, bias]
    const int nbchunks = wt.taps * wt.kchunks + a.kchunks2;
    const int kiters = a.halo ? wt.kchunks : nbchunks;
    const size_t a_stage = (size_t)ksub * (a.halo ? kHaloTileBytes : kSubTileBytes) +
                           (L.mode == EPI_UP ? (size_t)a.up_chunks * kUpTileBytes : 0);
    const size_t b_chunk = (size_t)wt.block_n * 128;
    const size_t budget = 226 * 1024;
    int min_stages = 2;
    if (L.mode == EPI_LN) {
        CIDNET_CHECK(wt.taps == 1 && wt.kchunks <= 4 && wt.bias && wt.wsum && !L.in2, CIDNET_ERR_INVALID,
                     "conv_gemm: LN needs a single-source 1x1 with K<=256");
        min_stages = wt.kchunks + 1;
    }
    const size_t bres = (size_t)nbchunks * b_chunk;
    // A tiles in flight.  Measured on B200 (cfg2 / cfg4): 2, 4 and 6 tiles give the same step time -- the
    // per-tile cost (~3000-3900 cycles whatever N and K) is the epilogue's dependent-latency chain, not the
    // load pipeline -- so the ring stays at 2 tiles.
    const int depth = 2;
    // UP: each epilogue group holds its tile's stages (the low-res box) until it is done -> one more tile of slack
    const int tiles_in_ring = L.mode == EPI_UP ? depth + kEpiGroups : depth;
    const int want = tiles_in_ring * kiters > min_stages ? tiles_in_ring * kiters : min_stages;
    int stages = 0;
    size_t fixed = 0;
    // preference: resident weights (2 staging buffers, then 1), else streamed weights (2, then 1)
    for (int pass = 0; pass < 4 && stages < min_stages; ++pass) {
        const int bufs = (pass & 1) ? 1 : 2;
        const bool resident = pass < 2;
        fixed = 1024 + (size_t)bufs * kEpiGroups * kStagingBytes + 512 + 2 * round_up(wt.block_n, 32) * sizeof(float);
        if (resident) {
            if (a.per_image_w || fixed + bres + (size_t)min_stages * a_stage > budget) continue;
            a.b_resident = 1; a.stg_bufs = bufs;
            stages = (int)((budget - fixed - bres) / a_stage);
        } else {
            if (a.halo || fixed + (size_t)min_stages * (a_stage + b_chunk) > budget) continue;
            a.b_resident = 0; a.stg_bufs = bufs;
            stages = (int)((budget - fixed) / (a_stage + b_chunk));
        }
    }
    if (a.wring) {
        // halo tiles + weight ring: as many halo stages as a second tile needs while >= 6 weight stages remain, else 2
        a.b_resident = 0; a.stg_bufs = 2;
        fixed = 1024 + (size_t)2 * kEpiGroups * kStagingBytes + 512 + 2 * round_up(wt.block_n, 32) * sizeof(float);
        // (measured: 3 halo stages + 6 weight stages and 2 + 11 give the same layer time -- neither ring is the bound)
        stages = want;
        while (stages > 2 && fixed + (size_t)stages * a_stage + 6 * b_chunk > budget) --stages;
        int ring = (int)((budget - fixed - (size_t)stages * a_stage) / b_chunk);
        a.wring = ring > kMaxWRing ? kMaxWRing : ring;
        CIDNET_CHECK(a.wring >= 2, CIDNET_ERR_INVALID, "conv_gemm: shared memory budget exceeded (weight ring)");
    }
    if (stages > 12) stages = 12;
    if (stages > want) stages = want;
    CIDNET_CHECK(stages >= min_stages, CIDNET_ERR_INVALID, "conv_gemm: shared memory budget exceeded");
    a.stages = stages;
    const size_t smem = a.wring ? fixed + (size_t)stages * a_stage + (size_t)a.wring * b_chunk
                                : fixed + (a.b_resident ? bres : 0) + (size_t)stages * (a_stage + (a.b_resident ? 0 : b_chunk));

    int gx = (L.max_ctas > 0 ? L.max_ctas : device_sm_count()) / wt.n_blocks;
    if (gx < 1) gx = 1;
    if (gx > a.num_tiles) gx = a.num_tiles;
    dim3 grid((unsigned)gx, (unsigned)wt.n_blocks, 1);
    switch (kmode) {
        case EPI_STORE: return launch_mode<EPI_STORE>(a, grid, smem, stream, kEpiGroups);
        case EPI_LN:    return launch_mode<EPI_LN>(a, grid, smem, stream, kEpiGroups);
        case EPI_DOWN:  return launch_mode<EPI_DOWN>(a, grid, smem, stream, kEpiGroups);
        case EPI_UP:    return launch_mode<EPI_UP>(a, grid, smem, stream, kEpiGroups);
        case EPI_UP_SA: return launch_mode<EPI_UP_SA>(a, grid, smem, stream, kEpiGroups);
    }
    return fail(CIDNET_ERR_INVALID, "conv_gemm: bad mode");
}

// ----------------------------------------------------------- layout helpers ---
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, act_t* __restrict__ dst, int C, long long hw,
                                    int pitch, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int c = (int)(i % pitch);
        const long long p = i / pitch;            // b*hw + pixel
        const long long b = p / hw, px = p - b * hw;
        dst[i] = f2act(c < C ? src[(b * C + c) * hw + px] : 0.f);
    }
}
__global__ void nhwc_to_nchw_kernel(const act_t* __restrict__ src, float* __restrict__ dst, int C, long long hw,
                                    int pitch, long long total) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long px = i % hw;
        const long long bc = i / hw;
        const long long b = bc / C;
        const int c = (int)(bc - b * C);
        dst[i] = act2f(src[(b * hw + px) * pitch + c]);
    }
}

int launch_nchw_to_nhwc(const float* src, act_t* dst, int B, int C, int H, int W, int pitch, cudaStream_t s) {
    const long long hw = (long long)H * W, total = (long long)B * hw * pitch;
    if (total == 0) return CIDNET_OK;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    nchw_to_nhwc_kernel<<<blocks, 256, 0, s>>>(src, dst, C, hw, pitch, total);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}
int launch_nhwc_to_nchw(const act_t* src, float* dst, int B, int C, int H, int W, int pitch, cudaStream_t s) {
    const long long hw = (long long)H * W, total = (long long)B * C * hw;
    if (total == 0) return CIDNET_OK;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    nhwc_to_nchw_kernel<<<blocks, 256, 0, s>>>(src, dst, C, hw, pitch, total);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
