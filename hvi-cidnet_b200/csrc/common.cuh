// Shared declarations for libcidnet_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>

#include "../../include/cidnet_b200.h"

namespace cidnet {

// Storage type of the internal NHWC activations and of the tensor-core operands.
// fp16 (10-bit mantissa) keeps the end-to-end error ~10x below the 2e-3 contract
// (SURVEY App. E); build with -DCIDNET_ACT_BF16 for bf16 operands instead.
#ifdef CIDNET_ACT_BF16
typedef __nv_bfloat16 act_t;
#define CIDNET_UMMA_FMT 1u
__host__ __device__ __forceinline__ float act2f(act_t v) { return __bfloat162float(v); }
__host__ __device__ __forceinline__ act_t f2act(float v) { return __float2bfloat16_rn(v); }
#else
typedef __half act_t;
#define CIDNET_UMMA_FMT 0u
__host__ __device__ __forceinline__ float act2f(act_t v) { return __half2float(v); }
// SATURATING on the device: a value beyond +-65504 becomes +-65504 instead of inf (an inf would turn into NaN in the
// next LayerNorm / softmax); F2FP.SATFINITE, the same single instruction as the plain conversion
__host__ __device__ __forceinline__ act_t f2act(float v) {
#ifdef __CUDA_ARCH__
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return __ushort_as_half(h);
#else
    return __float2half_rn(v);
#endif
}
#endif

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: raise it once per (current
// device, kernel) pair (a process may hold contexts on several GPUs; a process-wide `static bool` would leave the
// second device at the 48 KB default).  Returns CIDNET_OK or an error code.
int ensure_dynamic_smem(const void* kernel, int bytes);
// SM count of the CURRENT device (cached per device ordinal)
int device_sm_count();

#define CIDNET_CUDA_OK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            return ::cidnet::fail(CIDNET_ERR_CUDA, std::string(#expr) + ": " +            \
                                                       cudaGetErrorString(_e));           \
        }                                                                                 \
    } while (0)

#define CIDNET_CHECK(cond, code, msg)                                                     \
    do {                                                                                  \
        if (!(cond)) return ::cidnet::fail((code), (msg));                                \
    } while (0)

// Kernel launch with (optionally) programmatic stream serialization: see ptx_sm100.cuh pdl_wait / pdl_trigger.  Only
// kernels that execute pdl_wait() before touching non-constant global memory may be launched through this helper.
bool pdl_enabled();            // CIDNET_PDL=0 turns the attribute off (A/B runs)
// `cluster`: thread-block cluster dimensions (grid must be divisible by them); {1,1,1} = no cluster attribute.
template <typename... KArgs, typename... Args>
inline int launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, dim3 cluster,
                            Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster.x * cluster.y * cluster.z > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster.x; attr[n].val.clusterDim.y = cluster.y; attr[n].val.clusterDim.z = cluster.z;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    if (e != cudaSuccess) return fail(CIDNET_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    return CIDNET_OK;
}
template <typename... KArgs, typename... Args>
inline int launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_k_cluster(kernel, grid, block, smem, stream, dim3(1, 1, 1), std::forward<Args>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Rows per CTA of a row-walking kernel (dw3x3, IEL gate): every CTA walks `rows` output rows plus `halo` extra input rows, the
// grid is ctas_per_strip * ceil(H / rows) CTAs on `slots` concurrently resident CTAs.  A fixed strip height leaves a nearly
// empty last wave (cfg 2, L1 dw3x3: 660 CTAs on 592 slots = two waves of 34 rows; 40-row strips = ONE wave of 42 rows), so the
// height is chosen per launch to minimise waves * (rows + halo + fixed), `fixed` ~ the prologue in row-equivalents.
static inline int pick_strip_rows(int H, long long ctas_per_strip, int slots, int halo, int fixed, int lo, int hi) {
    int best = lo > H ? H : lo;
    long long best_cost = -1;
    for (int r = lo; r <= hi; ++r) {
        const int rows = r > H ? H : r;
        const long long ctas = ctas_per_strip * ((H + rows - 1) / rows);
        const long long waves = (ctas + slots - 1) / slots;
        const long long cost = waves * (rows + halo + fixed);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
        if (rows == H) break;
    }
    return best;
}
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// pitch (in elements) of an NHWC activation with C channels: multiple of 8 so that
// every pixel row is 16-byte aligned (TMA global strides must be multiples of 16 B).
static inline int act_pitch(int C) { return round_up(C, 8); }

// 8 packed activations <-> 8 floats (one 16-byte vector)
struct alignas(16) act8 { act_t v[8]; };

__device__ __forceinline__ void load8(const act_t* p, float* f) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const act_t* a = reinterpret_cast<const act_t*>(&raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = act2f(a[i]);
}
__device__ __forceinline__ void store8(act_t* p, const float* f) {
    uint4 raw;
    act_t* a = reinterpret_cast<act_t*>(&raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = f2act(f[i]);
    *reinterpret_cast<uint4*>(p) = raw;
}

}  // namespace cidnet
