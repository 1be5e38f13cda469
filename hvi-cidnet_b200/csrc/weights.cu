#include "weights.cuh"

namespace cidnet {

int upload_f32(float** dst, const std::vector<float>& v) {
    *dst = nullptr;
    if (v.empty()) return CIDNET_OK;
    CIDNET_CUDA_OK(cudaMalloc(dst, v.size() * sizeof(float)));
    CIDNET_CUDA_OK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return CIDNET_OK;
}

int upload_act(act_t** dst, const std::vector<float>& v) {
    *dst = nullptr;
    if (v.empty()) return CIDNET_OK;
    std::vector<act_t> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = f2act(v[i]);
    CIDNET_CUDA_OK(cudaMalloc(dst, h.size() * sizeof(act_t)));
    CIDNET_CUDA_OK(cudaMemcpy(*dst, h.data(), h.size() * sizeof(act_t), cudaMemcpyHostToDevice));
    return CIDNET_OK;
}

int pack_conv_segments(PackedWeights* out, const std::vector<WeightSegment>& segs, int cin, int taps, int n_out,
                       bool with_ln, int max_block) {
    PackedWeights p;
    p.cin = cin; p.taps = taps; p.kchunks = ceil_div(cin, 64); p.n_out = n_out; p.n_img = 1;
    choose_blocking(n_out, &p.block_n, &p.n_blocks, max_block);
    p.n_rows = p.block_n * p.n_blocks;
    const int kt = p.ktot();
    std::vector<float> packed((size_t)p.n_rows * kt, 0.f);
    std::vector<float> bias(p.n_rows, 0.f), wsum(p.n_rows, 0.f);
    for (const WeightSegment& sg : segs) {
        for (int s = 0; s < sg.n_src; ++s) {
            const int r = sg.dst_row0 + s;
            if (r < 0 || r >= n_out) return fail(CIDNET_ERR_INVALID, "pack_conv_segments: row out of range");
            double b = 0.0, ws = 0.0;
            for (int c = 0; c < cin; ++c) {
                for (int t = 0; t < taps; ++t) {
                    float v = sg.w[((size_t)s * cin + c) * taps + t];
                    if (sg.ln_w) v *= sg.ln_w[c];
                    const float vr = act2f(f2act(v));
                    packed[(size_t)r * kt + (size_t)t * p.kchunks * 64 + c] = vr;
                    ws += vr;
                }
                if (sg.ln_b) b += (double)sg.w[(size_t)s * cin * taps + (size_t)c * taps] * sg.ln_b[c];
            }
            bias[r] = (float)b;
            wsum[r] = (float)ws;
        }
    }
    int rc = upload_act(&p.w, packed);
    if (rc) return rc;
    if (with_ln) {
        if ((rc = upload_f32(&p.bias, bias))) return rc;
        if ((rc = upload_f32(&p.wsum, wsum))) return rc;
    }
    *out = p;
    return CIDNET_OK;
}

int pack_conv_weights(PackedWeights* out, const float* w, int n_src, int cin, int taps, const int* row_of_src,
                      int n_out, const float* ln_w, const float* ln_b, int max_block) {
    if (row_of_src != nullptr) return fail(CIDNET_ERR_INVALID, "pack_conv_weights: row maps go through segments");
    std::vector<WeightSegment> segs{{w, n_src, 0, ln_w, ln_b}};
    return pack_conv_segments(out, segs, cin, taps, n_out, ln_w != nullptr, max_block);
}

int pack_identity(PackedWeights* out, int c) {
    std::vector<float> eye((size_t)c * c, 0.f);
    for (int i = 0; i < c; ++i) eye[(size_t)i * c + i] = 1.f;
    return pack_conv_weights(out, eye.data(), c, c, 1, nullptr, c, nullptr, nullptr);
}

void free_packed(PackedWeights* p) {
    if (p->w) cudaFree(p->w);
    if (p->bias) cudaFree(p->bias);
    if (p->wsum) cudaFree(p->wsum);
    *p = PackedWeights();
}

}  // namespace cidnet
