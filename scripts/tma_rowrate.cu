// Micro-benchmark: what does a TMA box cost -- bytes, or rows?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/tma_rowrate scripts/tma_rowrate.cu
//   ./scripts/tma_rowrate
//
// One CTA per SM; one thread streams 16 KiB boxes of `inner` 16-bit elements per row between global memory
// ([rows][inner], row pitch = inner*2 bytes rounded up to what the layout under test uses) and shared memory:
//   load : cp.async.bulk.tensor.2d global -> smem, 4 boxes in flight (mbarrier ring)
//   store: cp.async.bulk.tensor.2d smem -> global, <= 4 bulk groups in flight
// for inner = 16 / 32 (32 / 64-byte rows: the IEL gate's 16-channel SWIZZLE_32B ring), 64 (128-byte rows,
// SWIZZLE_128B: the conv GEMM's operand / staging layout), and 128 / 192 / 256 elements (256 / 384 / 512-byte
// rows, no swizzle).  Prints cycles per row and bytes per cycle per SM.
// Motivation (profiles/r01_summary.md): every 1x1 conv GEMM of the forward costs ~7 cycles per 128-byte row
// request, loads and stores alike; if the cost is per ROW, wide un-swizzled store boxes cut the store side 2-4x.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
}

constexpr int kDepth = 4;

__global__ void __launch_bounds__(128, 1)
tma_kernel(const __grid_constant__ CUtensorMap tm, int tiles_per_cta, int box_bytes, int kRows, int is_store, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[kDepth];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kDepth; ++i) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kDepth * box_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    const int tile0 = blockIdx.x * tiles_per_cta;
    if (is_store) {
        for (int i = 0; i < tiles_per_cta; ++i) {
            asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(kDepth - 1) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                         :: "l"((uint64_t)&tm), "r"(s32(smem + (size_t)(i % kDepth) * box_bytes)), "r"(0), "r"((tile0 + i) * kRows) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        for (int i = 0; i < tiles_per_cta + kDepth; ++i) {
            if (i >= kDepth) mbar_wait(&bar[i % kDepth], ((i / kDepth) - 1) & 1);
            if (i < tiles_per_cta) {
                mbar_expect(&bar[i % kDepth], box_bytes);
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                             :: "r"(s32(smem + (size_t)(i % kDepth) * box_bytes)), "l"((uint64_t)&tm), "r"(s32(&bar[i % kDepth])),
                                "r"(0), "r"((tile0 + i) * kRows) : "memory");
            }
        }
    }
    cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fnp;
    const int bytes_per_cta = 3 << 20;      // every configuration moves the same 3 MiB per SM in 16 KiB boxes (8 KiB for 32-byte rows)
    long long* d_cyc; CK(cudaMalloc(&d_cyc, sms * sizeof(long long)));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    printf("%d SMs, %d KiB per SM, %d boxes in flight\n", sms, bytes_per_cta >> 10, kDepth);
    printf("%-6s %-6s %-9s %10s %9s %12s %14s %12s\n", "op", "inner", "swizzle", "row bytes", "box rows", "cycles/row", "B/cycle/SM", "chip GB/s");
    for (int is_store = 0; is_store < 2; ++is_store) {
        for (int inner : {16, 32, 64, 64, 128, 192, 256}) {
            static int first64 = 1;
            const bool swz = (inner == 64) && first64; if (inner == 64) first64 = !first64;
            const bool swz32 = inner == 16;      // the IEL gate's / dw3x3 v2's ring: 16-channel rows, SWIZZLE_32B
            const size_t row_bytes = (size_t)inner * 2;
            const int kRows = (int)(16384 / row_bytes > 256 ? 256 : 16384 / row_bytes);     // box rows (<= 256 per box dimension)
            const int tiles_per_cta = bytes_per_cta / (kRows * (int)row_bytes);
            const size_t rows = (size_t)sms * tiles_per_cta * kRows;
            void* buf; CK(cudaMalloc(&buf, rows * row_bytes)); CK(cudaMemset(buf, 1, rows * row_bytes));
            CUtensorMap tm;
            cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)row_bytes};
            cuuint32_t box[2] = {(cuuint32_t)inner, (cuuint32_t)kRows}, estr[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             swz ? CU_TENSOR_MAP_SWIZZLE_128B : (swz32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d (inner %d)\n", (int)r, inner); return 1; }
            const int box_bytes = kRows * (int)row_bytes;
            cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            float best = 1e30f; std::vector<long long> cyc(sms);
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                tma_kernel<<<sms, 128, kDepth * box_bytes>>>(tm, tiles_per_cta, box_bytes, kRows, is_store, d_cyc);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) { best = ms; CK(cudaMemcpy(cyc.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost)); }
            }
            double avg = 0; for (long long c : cyc) avg += (double)c; avg /= sms;
            printf("%-6s %-6d %-9s %10zu %9d %12.2f %14.1f %12.0f\n", is_store ? "store" : "load", inner, swz ? "128B" : (swz32 ? "32B" : "none"), row_bytes, kRows,
                   avg / ((double)tiles_per_cta * kRows), (double)tiles_per_cta * box_bytes / avg, rows * row_bytes / (best * 1e-3) / 1e9);
            CK(cudaFree(buf));
        }
    }
    return 0;
}
