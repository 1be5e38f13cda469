// Host-side description of one implicit-GEMM convolution launch (conv_gemm.cu).
#pragma once
#include "common.cuh"

namespace cidnet {

enum EpiMode {
    EPI_STORE = 0,  // out = acc [-> PReLU]   (a residual is folded into the MMA: K-extension with identity weights)
    EPI_LN    = 1,  // out = rstd[p] * (acc - mean[p] * wsum[n]) + bias[n]   (LayerNorm folded into the 1x1)
    EPI_DOWN  = 2,  // out = PReLU(bilinear_x0.5(acc))                        (NormDownsample)
    EPI_UP    = 3,  // out = PReLU(acc + bilinear_x2(t)[p, n])                (NormUpsample tail)
    EPI_UP_SA = 4,  // EPI_UP that also writes the per-pixel channel (mean, max) of its output (chosen by the launcher when
                    // ConvGemmLaunch::sa_stats is set; MSSA variant)
};

// Weights of one GEMM, packed on the device as act_t [n_img][n_rows][taps*kchunks*64]
// (K-major, every tap's Cin zero-padded to kchunks*64 so each TMA box is one swizzle atom).
struct PackedWeights {
    act_t* w = nullptr;
    int n_rows = 0;      // rows present in memory (>= n_blocks*block_n, zero padded)
    int n_out = 0;       // output channels actually stored (row index == output channel)
    int cin = 0;
    int taps = 1;
    int kchunks = 1;
    int block_n = 16;
    int n_blocks = 1;
    int n_img = 1;       // >1: per-image weights (CAB fold), indexed by batch
    float* bias = nullptr;   // [n_rows]  (EPI_LN)
    float* wsum = nullptr;   // [n_rows]  (EPI_LN)
    int ktot() const { return taps * kchunks * 64; }
};

// N blocking: <=256 accumulator columns per block (two TMEM buffers of block_n columns must fit the
// 512-column TMEM; EPI_DOWN keeps two accumulators per buffer -> max_block 128)
static inline void choose_blocking(int n_out, int* block_n, int* n_blocks, int max_block = 256) {
    int nb = ceil_div(n_out, max_block);
    // with several N blocks every block must end on a 64-channel boundary: the epilogue stores whole
    // 64-channel boxes and only the tensor extent (not the neighbouring block) clips them
    int bn = nb > 1 ? round_up(ceil_div(n_out, nb), 64) : round_up(n_out, 16);
    *block_n = bn;
    *n_blocks = nb;
}

struct ConvGemmLaunch {
    EpiMode mode = EPI_STORE;
    // input activation, NHWC act_t
    const act_t* in = nullptr;
    int B = 0, H = 0, W = 0, in_pitch = 0;
    bool flat = false;            // 1x1 only: tile over the flattened H*W axis (128 consecutive pixels)
    const PackedWeights* wt = nullptr;
    bool dynamic_weights = false; // wt->w is written by an earlier kernel of the same forward (the CAB fold), not a model constant
    // output
    act_t* out = nullptr;
    int out_pitch = 0;
    // optional second K source (same pixel grid, 1x1): acc += in2 * wt2^T.  Used with identity
    // weights for residual adds (x + f(x)) so the epilogue never reads the residual from global.
    const act_t* in2 = nullptr;
    int in2_pitch = 0;
    const PackedWeights* wt2 = nullptr;
    bool use_prelu = false;
    float prelu = 0.f;
    // EPI_LN
    float ln_eps = 1e-6f;
    // EPI_UP: low-resolution tensor t [B, H/2, W/2, up_pitch] to be upsampled x2 and added
    const act_t* up = nullptr;
    int up_pitch = 0;
    float2* sa_stats = nullptr;   // EPI_UP (MSSA variant): also write the per-pixel channel (mean, max) of the output, [B][H][W]
    // Row-strip sharding of ONE image (cidnet_forward_sharded): this launch covers local rows
    // [0, H) = global rows [grow, grow + H) of an image with gH rows on this launch's INPUT pixel
    // grid (gH == 0: not sharded).  Only the align_corners bilinear weights of EPI_DOWN / EPI_UP
    // depend on it (they are functions of the GLOBAL row index and size); grow must be even.
    int gH = 0, grow = 0;
    // persistent grid size limit (0 = one CTA per SM): two independent launches that run concurrently on
    // two streams take half the SMs each, so their fixed prologue / tail costs overlap
    int max_ctas = 0;
};

int launch_conv_gemm(const ConvGemmLaunch& L, cudaStream_t stream);

// layout helpers (tests / taps only)
int launch_nchw_to_nhwc(const float* src, act_t* dst, int B, int C, int H, int W, int pitch, cudaStream_t s);
int launch_nhwc_to_nchw(const act_t* src, float* dst, int B, int C, int H, int W, int pitch, cudaStream_t s);

}  // namespace cidnet
