// TMA-fed tcgen05 implicit-GEMM convolution for sm_100a.
//
//   D[128 pixels, N] += A[128 pixels, K] * B[N, K]^T      (fp16/bf16 operands, fp32 accumulate in TMEM)
//
// One CTA = one tile of 128 output pixels (a TH x TW rectangle of one image, or 128
// consecutive pixels for "flat" 1x1 layers) x one block of <=256 output channels.
//   warp 0   : TMA producer.  For every (tap, 64-channel chunk) it issues ONE box load of the
//              NHWC activation (shifted by the tap offset; TMA zero-fills out-of-image pixels,
//              which IS the conv's zero padding, and channels >= Cin) and ONE box load of the
//              packed weights, both landing in the canonical SWIZZLE_128B K-major layout.
//   warp 1   : allocates TMEM, then a single thread issues tcgen05.mma (M=128, N=block_n, K=16).
//   warps 2-5: epilogue.  tcgen05.ld the accumulator rows (thread == pixel), apply the fused
//              tail and store NHWC act_t:
//                EPI_STORE  (+residual, PReLU)                      CAB fold GEMM, IEL project_out, ...
//                EPI_LN     LayerNorm folded in: rstd*(acc-mean*wsum)+bias, stats read from the
//                           A tile that is still resident in shared memory
//                EPI_DOWN   conv3x3 -> bilinear x0.5 (align_corners) -> PReLU; the tile is 16
//                           columns x 8 ROW PAIRS, even rows in accumulator 0 and odd rows in
//                           accumulator 1 (5-D TMA view), so the vertical lerp is thread-local
//                           and the horizontal one is a single warp shuffle
//                EPI_UP     1x1 on the skip + bilinear x2 of the (pre-composed) low-res conv -> PReLU
//
// Replaces nn.Conv2d calls at net/LCA.py:13,15,17,51,57 and net/transformer_utils.py:39,58,60
// together with the LayerNorm (:25-28), UpsamplingBilinear2d (:40,:59), cat (:64) and PReLU
// (:43,:66) around them.
#include "conv_gemm.cuh"
#include "ptx_sm100.cuh"

#include <mutex>

namespace cidnet {

// ------------------------------------------------------------------ params ---
struct ConvGemmArgs {
    CUtensorMap tmA;
    CUtensorMap tmB;
    int taps, kchunks, cin;
    int Hv, Wv, TH, TW, tiles_x, tiles_y;
    int n_out, block_n, stages, per_image_w;
    int w_real;                      // real image width (pixels) of the OUTPUT grid of this launch
    act_t* out; int out_pitch; long long out_img_stride;
    const act_t* res; int res_pitch; long long res_img_stride;
    const float* bias; const float* wsum; float ln_eps;
    const act_t* up; int up_H, up_W, up_pitch; long long up_img_stride; float up_ry, up_rx;
    float prelu; int use_prelu;
    float down_ry, down_rx; int in_H, in_W;
};

static constexpr int kThreads = 192;
static constexpr uint32_t kSubTileBytes = 128 * 128;   // 128 rows x 64 elements x 2 B

__device__ __forceinline__ float prelu_f(float v, float slope) { return v >= 0.f ? v : v * slope; }

// store `n` (<=32) channels starting at p; n is rounded up to a multiple of 8 (the pitch is)
__device__ __forceinline__ void store_chunk(act_t* p, const float* v, int n) {
    if (n > 0) store8(p, v);
    if (n > 8) store8(p + 8, v + 8);
    if (n > 16) store8(p + 16, v + 16);
    if (n > 24) store8(p + 24, v + 24);
}

template <int kMode>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    constexpr int kSub = (kMode == EPI_DOWN) ? 2 : 1;
    const uint32_t a_stage = kSub * kSubTileBytes;
    const uint32_t b_stage = (uint32_t)a.block_n * 128u;
    const int stages = a.stages;
    uint8_t* smA = smem;
    uint8_t* smB = smem + (size_t)stages * a_stage;
    uint64_t* full = reinterpret_cast<uint64_t*>(smB + (size_t)stages * b_stage);
    uint64_t* empty = full + stages;
    uint64_t* tmem_full = empty + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);     // [block_n] bias, then [block_n] wsum (EPI_LN)
    float* s_wsum = s_bias + a.block_n;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile decode
    const int tiles_per_img = a.tiles_x * a.tiles_y;
    const int img = blockIdx.x / tiles_per_img;
    const int trem = blockIdx.x - img * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * a.TH;
    const int x0 = (trem % a.tiles_x) * a.TW;
    const int n0 = blockIdx.y * a.block_n;
    const int kiters = a.taps * a.kchunks;

    uint32_t ncols = 32;
    while (ncols < (uint32_t)(kSub * a.block_n)) ncols <<= 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&a.tmA);
        ptx::prefetch_tensormap(&a.tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < stages; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 1); }
            ptx::mbar_init(tmem_full, 1);
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
        // ------------------------------------------------------- TMA producer
        if (lane == 0) {
            for (int i = 0; i < kiters; ++i) {
                const int s = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                ptx::mbar_wait(&empty[s], ph ^ 1u);
                ptx::mbar_expect_tx(&full[s], a_stage + b_stage);
                const int tap = i / a.kchunks;
                const int kc = i - tap * a.kchunks;
                int dy = 1, dx = 1;
                if (a.taps == 9) { dy = tap / 3; dx = tap - dy * 3; }
                uint8_t* dstA = smA + (size_t)s * a_stage;
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
                ptx::tma_load_3d(smB + (size_t)s * b_stage, &a.tmB, &full[s], i * 64, n0,
                                 a.per_image_w ? img : 0);
            }
        }
    } else if (warp == 1) {
        // --------------------------------------------------------- MMA issuer
        const uint32_t idesc = ptx::umma_idesc_f16(CIDNET_UMMA_FMT, (uint32_t)a.block_n);
        for (int i = 0; i < kiters; ++i) {
            const int s = i % stages;
            const uint32_t ph = (uint32_t)(i / stages) & 1u;
            ptx::mbar_wait(&full[s], ph);
            ptx::tc_fence_after();
            if (lane == 0) {
                const int kc = i % a.kchunks;
                int ksteps = (a.cin - kc * 64 + 15) >> 4;
                if (ksteps > 4) ksteps = 4;
                const uint64_t descB = ptx::umma_smem_desc_sw128(ptx::smem_u32(smB + (size_t)s * b_stage));
#pragma unroll
                for (int sub = 0; sub < kSub; ++sub) {
                    const uint64_t descA = ptx::umma_smem_desc_sw128(
                        ptx::smem_u32(smA + (size_t)s * a_stage + sub * kSubTileBytes));
                    for (int k = 0; k < ksteps; ++k) {
                        ptx::umma_f16(tmem_base + sub * a.block_n, descA + 2 * k, descB + 2 * k, idesc,
                                      (uint32_t)((i | k) != 0));
                    }
                }
                ptx::umma_commit(&empty[s]);                       // frees the smem slot when the MMAs retire
                if (i == kiters - 1) ptx::umma_commit(tmem_full);  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ----------------------------------------------------------- epilogue
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;          // accumulator row == pixel within the tile
        const int ty = row / a.TW;
        const int tx = row - ty * a.TW;
        const int y = y0 + ty;
        const int x = x0 + tx;
        const bool valid = (y < a.Hv) && (x < a.Wv);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);

        float mean = 0.f, rstd = 1.f;
        if (kMode == EPI_LN) {
            // bias / wsum of this channel block -> shared memory (read once per CTA instead of once
            // per 16-column chunk from global); named barrier 1 = the 128 epilogue threads only
            for (int i = row; i < a.block_n; i += 128) {
                s_bias[i] = __ldg(a.bias + n0 + i);
                s_wsum[i] = __ldg(a.wsum + n0 + i);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // per-pixel LayerNorm statistics from the A tile (all K chunks are still resident:
            // the host guarantees stages >= kchunks and taps == 1 for this mode)
            for (int kc = 0; kc < a.kchunks; ++kc) ptx::mbar_wait(&full[kc], 0);
            float sum = 0.f;
            for (int pass = 0; pass < 2; ++pass) {
                float acc = 0.f;
                for (int kc = 0; kc < a.kchunks; ++kc) {
                    const uint8_t* rowp = smA + (size_t)kc * a_stage + row * 128;
                    const int cbase = kc * 64;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (cbase + j * 8 < a.cin) {
                            float f[8];
                            load8(reinterpret_cast<const act_t*>(rowp + ((j ^ (row & 7)) << 4)), f);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                if (cbase + j * 8 + e < a.cin) {
                                    if (pass == 0) acc += f[e];
                                    else { const float d = f[e] - mean; acc += d * d; }
                                }
                            }
                        }
                    }
                }
                if (pass == 0) { sum = acc; mean = sum / (float)a.cin; }
                else rstd = 1.0f / sqrtf(acc / (float)a.cin + a.ln_eps);
            }
        }

        // pull the rows this thread will add (residual / low-res upsample taps) towards L1 while
        // the MMAs are still running
        if (kMode == EPI_STORE && a.res != nullptr && valid) {
            const char* rp = reinterpret_cast<const char*>(a.res + (long long)img * a.res_img_stride +
                                                           ((long long)y * a.Wv + x) * a.res_pitch + n0);
            const int bytes = min(a.block_n, a.n_out - n0) * 2;
            for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L1 [%0];" :: "l"(rp + o));
        }
        if (kMode == EPI_UP && valid) {
            const long long pix = (long long)y * a.Wv + x;
            const int yr = (int)(pix / a.w_real), xr = (int)(pix - (long long)yr * a.w_real);
            const int i0 = (int)(a.up_ry * (float)yr), j0 = (int)(a.up_rx * (float)xr);
            const int i1 = i0 + (i0 < a.up_H - 1 ? 1 : 0), j1 = j0 + (j0 < a.up_W - 1 ? 1 : 0);
            const char* tb = reinterpret_cast<const char*>(a.up + (long long)img * a.up_img_stride + n0);
            const int bytes = min(a.block_n, a.n_out - n0) * 2;
            for (int o = 0; o < bytes; o += 128) {
                asm volatile("prefetch.global.L1 [%0];" :: "l"(tb + ((long long)i0 * a.up_W + j0) * a.up_pitch * 2 + o));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(tb + ((long long)i0 * a.up_W + j1) * a.up_pitch * 2 + o));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(tb + ((long long)i1 * a.up_W + j0) * a.up_pitch * 2 + o));
                asm volatile("prefetch.global.L1 [%0];" :: "l"(tb + ((long long)i1 * a.up_W + j1) * a.up_pitch * 2 + o));
            }
        }

        ptx::mbar_wait(tmem_full, 0);
        ptx::tc_fence_after();

        if (kMode == EPI_DOWN) {
            // y counts ROW PAIRS == output rows; x is the input column (even lanes own an output pixel)
            const int oy = y, ox = x >> 1;
            const float sy = a.down_ry * (float)oy;
            const int i0 = (int)sy;
            const float ly = sy - (float)i0;
            const bool top_is_odd = (i0 - 2 * oy) != 0;      // only at the clamped last row
            const float sx = a.down_rx * (float)ox;
            const int j0 = (int)sx;
            const float lx = sx - (float)j0;
            const bool left_is_right = (j0 - 2 * ox) != 0;   // only at the clamped last column
            const bool writer = valid && ((x & 1) == 0);
            act_t* outp = a.out + (long long)img * a.out_img_stride +
                          ((long long)oy * (a.Wv >> 1) + ox) * a.out_pitch + n0;
            for (int c = 0; c < a.block_n; c += 32) {
                if (n0 + c >= a.n_out) break;
                float e[32], o[32];
                ptx::tmem_ld32(taddr + c, e);
                ptx::tmem_ld32(taddr + a.block_n + c, o);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float top = top_is_odd ? o[j] : e[j];
                    const float v = (1.f - ly) * top + ly * o[j];
                    const float vr = __shfl_down_sync(0xffffffffu, v, 1);
                    const float left = left_is_right ? vr : v;
                    e[j] = prelu_f((1.f - lx) * left + lx * vr, a.prelu);
                }
                if (writer) store_chunk(outp + c, e, a.n_out - (n0 + c));
            }
        } else {
            const long long pix = (long long)y * a.Wv + x;
            act_t* outp = a.out + (long long)img * a.out_img_stride + pix * a.out_pitch + n0;
            const act_t* resp = (kMode == EPI_STORE && a.res != nullptr)
                                    ? a.res + (long long)img * a.res_img_stride + pix * a.res_pitch + n0 : nullptr;
            // EPI_UP: bilinear x2 taps of the low-res tensor (align_corners=True)
            const act_t *t00 = nullptr, *t01 = nullptr, *t10 = nullptr, *t11 = nullptr;
            float uly = 0.f, ulx = 0.f;
            if (kMode == EPI_UP && valid) {
                const int yr = (int)(pix / a.w_real), xr = (int)(pix - (long long)yr * a.w_real);
                const float sy = a.up_ry * (float)yr;
                const int i0 = (int)sy; uly = sy - (float)i0;
                const int i1 = i0 + (i0 < a.up_H - 1 ? 1 : 0);
                const float sx = a.up_rx * (float)xr;
                const int j0 = (int)sx; ulx = sx - (float)j0;
                const int j1 = j0 + (j0 < a.up_W - 1 ? 1 : 0);
                const act_t* tb = a.up + (long long)img * a.up_img_stride + n0;
                t00 = tb + ((long long)i0 * a.up_W + j0) * a.up_pitch;
                t01 = tb + ((long long)i0 * a.up_W + j1) * a.up_pitch;
                t10 = tb + ((long long)i1 * a.up_W + j0) * a.up_pitch;
                t11 = tb + ((long long)i1 * a.up_W + j1) * a.up_pitch;
            }
            for (int c = 0; c < a.block_n; c += 32) {
                if (n0 + c >= a.n_out) break;
                float v[32];
                ptx::tmem_ld32(taddr + c, v);
                const int nrem = a.n_out - (n0 + c);
                if (kMode == EPI_LN) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = (c + j < a.block_n) ? c + j : 0;      // bias / wsum staged for block_n channels
                        v[j] = rstd * (v[j] - mean * s_wsum[n]) + s_bias[n];
                    }
                } else if (kMode == EPI_STORE) {
                    if (resp != nullptr && valid) {
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            if (nrem > 8 * h) {
                                float r[8];
                                load8(resp + c + 8 * h, r);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[8 * h + j] += r[j];
                            }
                        }
                    }
                    if (a.use_prelu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = prelu_f(v[j], a.prelu);
                    }
                } else if (kMode == EPI_UP) {
                    if (valid) {
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            if (nrem > 8 * h) {
                                float p00[8], p01[8], p10[8], p11[8];
                                load8(t00 + c + 8 * h, p00); load8(t01 + c + 8 * h, p01);
                                load8(t10 + c + 8 * h, p10); load8(t11 + c + 8 * h, p11);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float up = (1.f - uly) * ((1.f - ulx) * p00[j] + ulx * p01[j]) +
                                                     uly * ((1.f - ulx) * p10[j] + ulx * p11[j]);
                                    v[8 * h + j] = prelu_f(v[8 * h + j] + up, a.prelu);
                                }
                            }
                        }
                    }
                }
                if (valid) store_chunk(outp + c, v, nrem);
            }
        }
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
                      const uint32_t* box) {
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
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

static inline float ac_scale(int n_in, int n_out) {
    // torch area_pixel_compute_scale<float>(align_corners=True)
    return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
}

template <int kMode>
static int launch_mode(const ConvGemmArgs& args, dim3 grid, size_t smem, cudaStream_t stream) {
    static bool configured = false;   // per mode; attribute is sticky per function
    if (!configured) {
        CIDNET_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            227 * 1024));
        configured = true;
    }
    conv_gemm_kernel<kMode><<<grid, kThreads, smem, stream>>>(args);
    CIDNET_CUDA_OK(cudaGetLastError());
    return CIDNET_OK;
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

    ConvGemmArgs a;
    memset(&a, 0, sizeof a);
    a.taps = wt.taps; a.kchunks = wt.kchunks; a.cin = wt.cin;
    a.n_out = wt.n_out; a.block_n = wt.block_n; a.per_image_w = wt.n_img > 1 ? 1 : 0;
    a.out = L.out; a.out_pitch = L.out_pitch;
    a.res = L.res; a.res_pitch = L.res_pitch;
    a.bias = wt.bias; a.wsum = wt.wsum; a.ln_eps = L.ln_eps;
    a.prelu = L.prelu; a.use_prelu = L.use_prelu ? 1 : 0;
    a.w_real = L.W;
    const long long hw = (long long)L.H * L.W;
    a.res_img_stride = hw * L.res_pitch;

    const uint64_t pb = (uint64_t)L.in_pitch * sizeof(act_t);   // bytes per pixel row
    if (L.mode == EPI_DOWN) {
        CIDNET_CHECK(L.H % 2 == 0 && L.W % 2 == 0 && wt.taps == 9, CIDNET_ERR_INVALID, "conv_gemm: DOWN needs even H,W, 3x3");
        a.Hv = L.H / 2; a.Wv = L.W; a.TH = 8; a.TW = 16;
        a.in_H = L.H; a.in_W = L.W;
        a.down_ry = ac_scale(L.H, L.H / 2); a.down_rx = ac_scale(L.W, L.W / 2);
        a.out_img_stride = (hw / 4) * L.out_pitch;
        const uint64_t dims[5] = {(uint64_t)wt.cin, (uint64_t)L.W, 2, (uint64_t)L.H / 2, (uint64_t)L.B};
        const uint64_t str[4] = {pb, pb * L.W, pb * L.W * 2, pb * hw};
        const uint32_t box[5] = {64, 16, 1, 8, 1};
        int rc = encode_map(&a.tmA, L.in, 5, dims, str, box);
        if (rc) return rc;
    } else {
        if (L.flat) { a.Hv = 1; a.Wv = (int)hw; a.TH = 1; a.TW = 128; }
        else        { a.Hv = L.H; a.Wv = L.W; a.TH = 8; a.TW = 16; }
        a.out_img_stride = hw * L.out_pitch;
        const uint64_t dims[4] = {(uint64_t)wt.cin, (uint64_t)a.Wv, (uint64_t)a.Hv, (uint64_t)L.B};
        const uint64_t str[3] = {pb, pb * a.Wv, pb * hw};
        const uint32_t box[4] = {64, (uint32_t)a.TW, (uint32_t)a.TH, 1};
        int rc = encode_map(&a.tmA, L.in, 4, dims, str, box);
        if (rc) return rc;
        if (L.mode == EPI_UP) {
            CIDNET_CHECK(L.up != nullptr && L.H % 2 == 0 && L.W % 2 == 0, CIDNET_ERR_INVALID, "conv_gemm: UP needs t");
            a.up = L.up; a.up_H = L.H / 2; a.up_W = L.W / 2; a.up_pitch = L.up_pitch;
            a.up_img_stride = (hw / 4) * L.up_pitch;
            a.up_ry = ac_scale(L.H / 2, L.H); a.up_rx = ac_scale(L.W / 2, L.W);
        }
    }
    a.tiles_x = ceil_div(a.Wv, a.TW);
    a.tiles_y = ceil_div(a.Hv, a.TH);
    {
        const uint64_t kt = (uint64_t)wt.ktot();
        const uint64_t dims[3] = {kt, (uint64_t)wt.n_rows, (uint64_t)wt.n_img};
        const uint64_t str[2] = {kt * sizeof(act_t), kt * sizeof(act_t) * wt.n_rows};
        const uint32_t box[3] = {64, (uint32_t)wt.block_n, 1};
        int rc = encode_map(&a.tmB, wt.w, 3, dims, str, box);
        if (rc) return rc;
    }

    const int kiters = wt.taps * wt.kchunks;
    const int ksub = L.mode == EPI_DOWN ? 2 : 1;
    const size_t stage_bytes = (size_t)ksub * kSubTileBytes + (size_t)wt.block_n * 128;
    int stages = (int)((96 * 1024) / stage_bytes);
    if (stages < 2) stages = 2;
    if (L.mode == EPI_LN) {
        CIDNET_CHECK(wt.taps == 1 && wt.kchunks <= 4 && wt.bias && wt.wsum, CIDNET_ERR_INVALID, "conv_gemm: LN needs 1x1, K<=256");
        if (stages < wt.kchunks) stages = wt.kchunks;
    }
    if (stages > kiters) stages = kiters;
    if (stages > 8) stages = 8;
    a.stages = stages;
    const size_t smem = 1024 + stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 2 * wt.block_n * sizeof(float);
    CIDNET_CHECK(smem <= 227 * 1024, CIDNET_ERR_INVALID, "conv_gemm: shared memory budget exceeded");

    dim3 grid((unsigned)(a.tiles_x * a.tiles_y * L.B), (unsigned)wt.n_blocks, 1);
    switch (L.mode) {
        case EPI_STORE: return launch_mode<EPI_STORE>(a, grid, smem, stream);
        case EPI_LN:    return launch_mode<EPI_LN>(a, grid, smem, stream);
        case EPI_DOWN:  return launch_mode<EPI_DOWN>(a, grid, smem, stream);
        case EPI_UP:    return launch_mode<EPI_UP>(a, grid, smem, stream);
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
