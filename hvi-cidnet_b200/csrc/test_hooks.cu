// Unit-test entry points (parity tests call single kernels through the C ABI).
// Unlike the production entry points these allocate scratch memory and synchronise.
#include "conv_gemm.cuh"
#include "weights.cuh"

using namespace cidnet;

// x: dev fp32 [B,Cin,H,W]; w_host: host fp32 [Cout,Cin,k,k]; out: dev fp32
//   mode 0 (STORE): out [B,Cout,H,W] = conv(x) (+ aux [B,Cout,H,W]) (-> PReLU if prelu != 0)
//   mode 1 (LN)   : out = conv1x1(LayerNorm(x)); ln_host = [ln_w(Cin), ln_b(Cin)]
//   mode 2 (DOWN) : out [B,Cout,H/2,W/2] = PReLU(bilinear_half(conv3x3(x)))
//   mode 3 (UP)   : out [B,Cout,H,W] = PReLU(conv1x1(x) + bilinear_x2(aux [B,Cout,H/2,W/2]))
extern "C" CIDNET_API int cidnet_test_conv(const float* x, const float* w_host, const float* aux,
                                           const float* ln_host, float* out, int B, int Cin, int H, int W,
                                           int Cout, int ksize, int mode, int flat, float prelu, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CIDNET_CHECK(ksize == 1 || ksize == 3, CIDNET_ERR_INVALID, "ksize must be 1 or 3");
    const int taps = ksize * ksize;
    PackedWeights pw, eye;
    int rc = pack_conv_weights(&pw, w_host, Cout, Cin, taps, nullptr, Cout,
                               mode == EPI_LN ? ln_host : nullptr, mode == EPI_LN ? ln_host + Cin : nullptr,
                               mode == EPI_DOWN ? 128 : 256);
    if (rc) return rc;
    const int pin = act_pitch(Cin), pout = act_pitch(Cout);
    const int Ho = mode == EPI_DOWN ? H / 2 : H, Wo = mode == EPI_DOWN ? W / 2 : W;
    act_t *xin = nullptr, *xout = nullptr, *xaux = nullptr;
    CIDNET_CUDA_OK(cudaMalloc(&xin, (size_t)B * H * W * pin * sizeof(act_t)));
    CIDNET_CUDA_OK(cudaMalloc(&xout, (size_t)B * Ho * Wo * pout * sizeof(act_t)));
    // poison the output so unwritten pixels are detected
    CIDNET_CUDA_OK(cudaMemsetAsync(xout, 0x7f, (size_t)B * Ho * Wo * pout * sizeof(act_t), stream));
    rc = launch_nchw_to_nhwc(x, xin, B, Cin, H, W, pin, stream);
    ConvGemmLaunch L;
    L.mode = (EpiMode)mode; L.in = xin; L.B = B; L.H = H; L.W = W; L.in_pitch = pin; L.flat = flat != 0;
    L.wt = &pw; L.out = xout; L.out_pitch = pout; L.prelu = prelu; L.use_prelu = prelu != 0.f;
    if (!rc && aux) {
        const int ah = mode == EPI_UP ? H / 2 : H, aw = mode == EPI_UP ? W / 2 : W;
        CIDNET_CUDA_OK(cudaMalloc(&xaux, (size_t)B * ah * aw * pout * sizeof(act_t)));
        rc = launch_nchw_to_nhwc(aux, xaux, B, Cout, ah, aw, pout, stream);
        if (mode == EPI_UP) { L.up = xaux; L.up_pitch = pout; }
        else {      // residual: second K source with identity weights
            if ((rc = pack_identity(&eye, Cout))) return rc;
            L.in2 = xaux; L.in2_pitch = pout; L.wt2 = &eye;
        }
    }
    if (!rc) rc = launch_conv_gemm(L, stream);
    if (!rc) rc = launch_nhwc_to_nchw(xout, out, B, Cout, Ho, Wo, pout, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    cudaFree(xin); cudaFree(xout); if (xaux) cudaFree(xaux);
    free_packed(&pw);
    free_packed(&eye);
    if (rc) return rc;
    CIDNET_CHECK(e == cudaSuccess, CIDNET_ERR_CUDA, std::string("test_conv: ") + cudaGetErrorString(e));
    return CIDNET_OK;
}
