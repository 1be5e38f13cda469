// Depthwise 3x3 convolution ON THE TENSOR CORES (sm_100a, TMA + tcgen05), NHWC 16-bit in / out.
//
// A depthwise conv has no channel mixing, but the tensor pipe is idle in this network while the
// CUDA-core version is issue-bound (9 FMAs + loads per output).  Per 16-channel slice the conv is
//     D[128 px, 16] += A_tap[128 px, 16] * diag(w_tap[16])          (M=128, N=16, K=16 tcgen05.mma)
// summed over the 9 taps, where A_tap is a ROW-SHIFTED VIEW of one shared-memory halo tile:
//   * TMA loads the tile + halo once: box {64 ch, 16 columns, 11 rows} -> [176 rows][128 B],
//     SWIZZLE_128B (out-of-image pixels zero-filled = the conv's zero padding);
//   * tap (dy,dx) = the 128 rows starting at row dy*16+dx (the swizzle XOR works on absolute smem
//     address bits, so a descriptor may start on any 128-byte row);  accumulator row m = ty*16+tx,
//     columns tx = 14,15 are wrap-around garbage and are dropped by the epilogue (tile = 8 x 14);
//   * the 4 slices of a 64-channel block are the 4 K-steps of the swizzle atom (descriptor +2 each),
//     their diagonal weight tiles [16 n][16 k] sit side by side in one [16][64] SWIZZLE_128B tile per
//     tap (18 KB per 64-channel block, resident in shared memory for the CTA's lifetime).
// 36 MMAs (8 cycles each) produce 128 x 64 outputs; the epilogue (TMEM -> fp16 -> swizzled staging ->
// TMA store, plus per-channel sum of squares for F.normalize) is the only CUDA-core work left.
//
// Used for q_dwconv / kv_dwconv of the CAB (net/LCA.py:14,16,22-23).
#include "dwtc.cuh"
#include "ptx_sm100.cuh"

#include <cstring>
#include <vector>

namespace cidnet {

int encode_map_generic(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box);   // conv_gemm.cu

static constexpr int kDwtcThreads = 64 + 128 * 2;          // TMA warp, MMA warp, 2 epilogue warpgroups
static constexpr uint32_t kHaloBytes = 11 * 16 * 128;      // 22.5 KB
static constexpr uint32_t kWTapBytes = 16 * 128;           // one tap's [16][64] weight tile
static constexpr uint32_t kStgBytes = 112 * 128;           // 8 x 14 pixels x 64 channels

struct DwtcSegDev {
    CUtensorMap tmIn, tmOut, tmW;
    float* ssq;            // [B][ssq_pitch] sum of squares per channel (nullptr: not needed)
    int channels, nblk, ssq_pitch, ssq_channels, blk0;   // blk0 = first blockIdx.y of this segment
};
struct DwtcArgs {
    DwtcSegDev seg[6];
    int nseg, H, W, tiles_x, tiles_y, num_tiles, stages;
};

__global__ void __launch_bounds__(kDwtcThreads, 1)
dwtc_kernel(const __grid_constant__ DwtcArgs a) {
    // 1024-byte alignment is required by the SWIZZLE_128B TMA / UMMA tiles; using the array directly
    // (no integer round trip) keeps the accesses in the shared state space (LDS / STS, not generic LD / ST)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
    const int stages = a.stages;
    uint8_t* smW = smem;                                        // 9 taps x 2 KB (padded to 18 KB)
    uint8_t* smA = smW + 9 * kWTapBytes;                        // halo ring
    uint8_t* smOut = smA + (size_t)stages * kHaloBytes;         // 2 groups x 2 staging buffers
    uint64_t* full = reinterpret_cast<uint64_t*>(smOut + 4 * kStgBytes);
    uint64_t* empty = full + stages;
    uint64_t* wfull = empty + stages;
    uint64_t* tmem_full = wfull + 1;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* s_ssq = reinterpret_cast<float*>(tmem_slot + 2);     // [64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // which (segment, 64-channel block) this CTA owns
    int si = 0;
    for (int i = 1; i < a.nseg; ++i) if ((int)blockIdx.y >= a.seg[i].blk0) si = i;
    const DwtcSegDev& sg = a.seg[si];
    const int blk = blockIdx.y - sg.blk0;
    const int ch_valid = min(64, sg.channels - blk * 64);
    const int nslices = (ch_valid + 15) >> 4;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&sg.tmIn); ptx::prefetch_tensormap(&sg.tmOut); ptx::prefetch_tensormap(&sg.tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
            ptx::mbar_init(wfull, 1);
            for (int i = 0; i < 2; ++i) { ptx::mbar_init(&tmem_full[i], 1); ptx::mbar_init(&tmem_empty[i], 4); }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(tmem_slot, 128);       // 2 accumulators x 64 fp32 columns
        ptx::tmem_relinquish();
    }
    if (threadIdx.x < 64) s_ssq[threadIdx.x] = 0.f;
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_expect_tx(wfull, 9 * kWTapBytes);
            for (int tap = 0; tap < 9; ++tap)
                ptx::tma_load_3d(smW + tap * kWTapBytes, &sg.tmW, wfull, 0, (blk * 9 + tap) * 16, 0);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
                const int img = t / tiles_per_img, trem = t - img * tiles_per_img;
                const int y0 = (trem / a.tiles_x) * 8, x0 = (trem % a.tiles_x) * 14;
                const int s = it % stages;
                ptx::mbar_wait(&empty[s], ((it / stages) & 1u) ^ 1u);
                ptx::mbar_expect_tx(&full[s], kHaloBytes);
                ptx::tma_load_4d(smA + (size_t)s * kHaloBytes, &sg.tmIn, &full[s], blk * 64, x0 - 1, y0 - 1, img);
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::umma_idesc_f16(CIDNET_UMMA_FMT, 16u);
        ptx::mbar_wait(wfull, 0);
        uint32_t it = 0;
        for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
            const uint32_t buf = it & 1u, s = it % stages;
            ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1u) ^ 1u);
            ptx::mbar_wait(&full[s], (it / stages) & 1u);
            ptx::tc_fence_after();
            if (lane == 0) {
                const uint32_t abase = ptx::smem_u32(smA + (size_t)s * kHaloBytes);
                const uint32_t wbase = ptx::smem_u32(smW);
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint64_t dA = ptx::umma_smem_desc_sw128(abase + (uint32_t)(dy * 16 + dx) * 128u);
                    const uint64_t dB = ptx::umma_smem_desc_sw128(wbase + tap * kWTapBytes);
                    for (int j = 0; j < nslices; ++j)
                        ptx::umma_f16(tmem_base + buf * 64 + j * 16, dA + 2 * j, dB + 2 * j, idesc, (uint32_t)(tap != 0));
                }
                ptx::umma_commit(&empty[s]);
                ptx::umma_commit(&tmem_full[buf]);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3, grp = (warp - 2) >> 2;
        const int row = q * 32 + lane, ty = row >> 4, tx = row & 15;
        const bool col_ok = tx < 14;
        const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
        uint8_t* const stg_base = smOut + (size_t)grp * 2 * kStgBytes;
        const uint32_t bar_id = 1 + grp;
        const int srow = ty * 14 + tx;
        float ssq[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) ssq[e] = 0.f;
        int cur_img = -1;
        uint32_t it = 0, sb = 0;
        // flush this thread's per-channel partial sums of squares (warp reduce -> smem -> global)
        auto flush = [&](int img) {
            if (sg.ssq == nullptr || img < 0) return;
#pragma unroll
            for (int e = 0; e < 64; ++e) {
                float v = ssq[e];
                v += __shfl_xor_sync(0xffffffffu, v, 16); v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 4);  v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (lane == 0 && blk * 64 + e < sg.ssq_channels) atomicAdd(sg.ssq + (long long)img * sg.ssq_pitch + blk * 64 + e, v);
                ssq[e] = 0.f;
            }
        };
        for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
            if ((int)(it & 1u) != grp) continue;
            const int img = t / tiles_per_img, trem = t - img * tiles_per_img;
            const int y0 = (trem / a.tiles_x) * 8, x0 = (trem % a.tiles_x) * 14;
            const bool valid = col_ok && (y0 + ty < a.H) && (x0 + tx < a.W);
            if (img != cur_img) { flush(cur_img); cur_img = img; }
            const uint32_t buf = it & 1u;
            ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1u);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + buf * 64 + ((uint32_t)(q * 32) << 16);
            uint8_t* stg = stg_base + sb * kStgBytes;
            if (issuer) ptx::tma_store_wait_read<1>();
            asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (hh * 32 >= nslices * 16) break;          // nothing valid in this half (warp-uniform)
                float v[32];
                ptx::tmem_ld32(taddr + hh * 32, v);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 raw;
                    act_t* o = reinterpret_cast<act_t*>(&raw);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        o[e] = f2act(v[8 * g + e]);
                        const float r = valid ? act2f(o[e]) : 0.f;      // squares of the ROUNDED values
                        ssq[hh * 32 + g * 8 + e] = fmaf(r, r, ssq[hh * 32 + g * 8 + e]);
                    }
                    if (col_ok) *reinterpret_cast<uint4*>(stg + srow * 128 + (((hh * 4 + g) ^ (srow & 7)) << 4)) = raw;
                }
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[buf]);
            if (issuer) {
                ptx::tma_store_4d(&sg.tmOut, stg, blk * 64, x0, y0, img);
                ptx::tma_store_commit();
            }
            sb ^= 1u;
        }
        flush(cur_img);
        if (issuer) ptx::tma_store_wait_all();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------ host -----
int pack_dwtc_weights(DwtcWeights* out, const float* w_tapmajor /*[9][pitch]*/, int pitch, int c_begin, int channels) {
    // [nblk * 9 * 16 rows][64]: row (blk*9+tap)*16 + n, column k:  w[tap][c] at n == k%16, c = blk*64 + (k/16)*16 + n
    const int nblk = ceil_div(channels, 64);
    std::vector<act_t> h((size_t)nblk * 9 * 16 * 64, f2act(0.f));
    for (int blk = 0; blk < nblk; ++blk)
        for (int tap = 0; tap < 9; ++tap)
            for (int k = 0; k < 64; ++k) {
                const int n = k & 15, c = blk * 64 + (k >> 4) * 16 + n;
                if (c < channels)
                    h[((size_t)(blk * 9 + tap) * 16 + n) * 64 + k] = f2act(w_tapmajor[(size_t)tap * pitch + c_begin + c]);
            }
    out->channels = channels; out->nblk = nblk; out->w = nullptr;
    CIDNET_CUDA_OK(cudaMalloc(&out->w, h.size() * sizeof(act_t)));
    CIDNET_CUDA_OK(cudaMemcpy(out->w, h.data(), h.size() * sizeof(act_t), cudaMemcpyHostToDevice));
    return CIDNET_OK;
}

int launch_dwtc(const DwtcLaunch& L, cudaStream_t stream) {
    CIDNET_CHECK(L.nseg >= 1 && L.nseg <= 6, CIDNET_ERR_INVALID, "dwtc: 1..6 segments");
    DwtcArgs a;
    memset(&a, 0, sizeof a);
    a.nseg = L.nseg; a.H = L.H; a.W = L.W;
    a.tiles_x = ceil_div(L.W, 14); a.tiles_y = ceil_div(L.H, 8);
    a.num_tiles = a.tiles_x * a.tiles_y * L.B;
    const long long hw = (long long)L.H * L.W;
    int blocks = 0, rc;
    for (int i = 0; i < L.nseg; ++i) {
        const DwtcSeg& s = L.seg[i];
        DwtcSegDev& d = a.seg[i];
        CIDNET_CHECK(s.in && s.out && s.wt && s.wt->w && s.in_pitch % 8 == 0 && s.out_pitch % 8 == 0, CIDNET_ERR_INVALID, "dwtc: bad segment");
        const uint64_t pi = (uint64_t)s.in_pitch * sizeof(act_t), po = (uint64_t)s.out_pitch * sizeof(act_t);
        const uint64_t dims[4] = {(uint64_t)s.wt->channels, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.B};
        const uint64_t si[3] = {pi, pi * L.W, pi * hw}, so[3] = {po, po * L.W, po * hw};
        const uint32_t ibox[4] = {64, 16, 11, 1}, obox[4] = {64, 14, 8, 1};
        if ((rc = encode_map_generic(&d.tmIn, s.in, 4, dims, si, ibox))) return rc;
        if ((rc = encode_map_generic(&d.tmOut, s.out, 4, dims, so, obox))) return rc;
        const uint64_t wd[3] = {64, (uint64_t)s.wt->nblk * 9 * 16, 1};
        const uint64_t ws[2] = {128, 128ull * s.wt->nblk * 9 * 16};
        const uint32_t wbox[3] = {64, 16, 1};
        if ((rc = encode_map_generic(&d.tmW, s.wt->w, 3, wd, ws, wbox))) return rc;
        d.ssq = s.ssq; d.ssq_pitch = s.ssq_pitch; d.ssq_channels = s.ssq_channels; d.channels = s.wt->channels; d.nblk = s.wt->nblk; d.blk0 = blocks;
        blocks += s.wt->nblk;
    }
    const size_t fixed = 1024 + 9 * kWTapBytes + 4 * kStgBytes + 1024;
    int stages = (int)((200 * 1024 - fixed) / kHaloBytes);
    if (stages > 5) stages = 5;
    a.stages = stages;
    const size_t smem = fixed + (size_t)stages * kHaloBytes;
    static bool configured = false;
    if (!configured) {
        CIDNET_CUDA_OK(cudaFuncSetAttribute(dwtc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int gx = 148 / blocks;
    if (gx < 1) gx = 1;
    if (gx > a.num_tiles) gx = a.num_tiles;
    dim3 grid(gx, blocks, 1);
    dwtc_kernel<<<grid, kDwtcThreads, smem, stream>>>(a);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
}

}  // namespace cidnet
