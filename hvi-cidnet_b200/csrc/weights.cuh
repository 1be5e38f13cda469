// Host-side packing of fp32 state_dict tensors into the device layouts of the kernels.
#pragma once
#include "conv_gemm.cuh"
#include <vector>

namespace cidnet {

// Pack a conv weight [n_src][cin][taps] (PyTorch [Cout][Cin][kh][kw], kh*kw == taps) into a
// PackedWeights.  row_of_src maps source row -> destination row (nullptr = identity);
// destination rows that receive nothing are zero.  With ln_w/ln_b (both [cin]) the
// LayerNorm affine is folded in:  W' = W * diag(ln_w)  (then rounded to act_t),
// wsum[n] = sum_c W'[n][c] (of the ROUNDED values, so the epilogue's mean correction
// cancels exactly), bias[n] = sum_c W[n][c] * ln_b[c].
struct WeightSegment {
    const float* w;      // [n_src][cin][taps]
    int n_src;
    int dst_row0;        // destination row of source row 0 (rows are contiguous)
    const float* ln_w;   // [cin] or nullptr
    const float* ln_b;   // [cin] or nullptr
};
// Several row blocks (each with its own LayerNorm fold) into one GEMM weight, e.g.
// [q of this LCA | k,v of the sibling LCA] sharing one A operand.
int pack_conv_segments(PackedWeights* out, const std::vector<WeightSegment>& segs, int cin, int taps, int n_out,
                       bool with_ln, int max_block = 256);
int pack_conv_weights(PackedWeights* out, const float* w, int n_src, int cin, int taps, const int* row_of_src,
                      int n_out, const float* ln_w, const float* ln_b, int max_block = 256);
// identity [c][c] as a 1x1 weight: second K source of a GEMM == residual add inside the MMA
int pack_identity(PackedWeights* out, int c);
void free_packed(PackedWeights* p);

// device buffer helpers
int upload_f32(float** dst, const std::vector<float>& v);
int upload_act(act_t** dst, const std::vector<float>& v);   // converts to act_t

}  // namespace cidnet
