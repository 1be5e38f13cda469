// Model context, weight folding/packing and the forward schedule behind the C ABI.
//
// Graph = /root/reference/net/CIDNet.py:71-122 with the dead I_LCA5 (:105) skipped.
// Every activation is NHWC, 16-bit (act_t), channel pitch rounded up to 8; the HVI image
// is kept as fp32 NCHW for the global residual (:119).
#include "cab.cuh"
#include "conv_gemm.cuh"
#include "iel.cuh"
#include "peer.cuh"
#include "sa.cuh"
#include "stem_head.cuh"
#include "weights.cuh"

#include <cstdlib>
#include <map>
#include <string>
#include <vector>

using namespace cidnet;

namespace {

// makes ctx->device current for the duration of an entry point (a process may drive several GPUs)
struct DeviceGuard {
    int prev = -1; bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

const int kCh[4] = {36, 36, 72, 144};
const int kHeads[4] = {1, 2, 4, 8};

struct LcaWeights {
    bool live = true;
    int C = 0, Cp = 0, heads = 0, h = 0, hp = 0;
    float* wqkv = nullptr;                               // depthwise weights [9][3*Cp]: q_dwconv | kv_dwconv (k) | (v)
    float* temp = nullptr;                               // [heads]
    float* wo = nullptr;                                 // [C][C]
    PackedWeights fold_tmpl;                             // geometry of the per-image folded weights
    PackedWeights w_in;                                  // project_in (LN folded), rows [x1 hp | x2 hp]
    float *dw0 = nullptr, *dw1 = nullptr, *dw2 = nullptr;
    PackedWeights w_out;                                 // project_out  h -> C
};

struct StageWeights {      // the I_/HV_ pair of one of the 6 LCA stages
    LcaWeights lca[2];     // 0 = I_LCA, 1 = HV_LCA
    PackedWeights qkv[2];  // GEMM on the I tensor: [q_I | kv_HV];  on the HV tensor: [q_HV | kv_I]
};

struct DownWeights { PackedWeights w; float prelu = 0.25f; };
struct UpWeights { PackedWeights w3; PackedWeights w1; float prelu = 0.25f; };

struct Tap { const void* ptr; int C, H, W, pitch; bool f32_nchw; };

}  // namespace

struct cidnet_ctx {
    int device = 0;
    bool finalized = false;
    std::map<std::string, std::vector<float>> raw;
    float k_host = 0.2f;
    float* k_dev = nullptr;
    float *stem_whv = nullptr, *stem_wi = nullptr, *head_wi = nullptr, *head_whv = nullptr;
    const uint2 *stem_bfrag = nullptr, *head_bfrag = nullptr;
    DownWeights down[2][3];    // [branch 0=I,1=HV][block1..3]
    UpWeights up[2][3];        // [branch][block3, block2, block1]  (index 0 = block3)
    StageWeights stage[6];
    PackedWeights eye[4];      // identity weights per level (residual adds inside the MMA)
    int variant = CIDNET_VARIANT_BASE;   // CIDNET_VARIANT_MSSA: net/CIDNet_MSSA.py (spatial-attention gates, live I_LCA5)
    float* sa_w[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // [branch][output level]: sa_{i,hv}{1,2,3}.conv1.weight
    std::vector<void*> owned;  // every cudaMalloc'ed pointer (freed in destroy)
    std::map<std::string, Tap> taps;
    int last_B = 0;
    int launches = 0;
    // optional per-launch profiling (bench.py roofline): event i is recorded BEFORE launch i,
    // one more after the last launch
    // CUDA-graph replay of the forward: one entry per (shape, workspace, flags); the input / output
    // image pointers are patched into the stem / head kernel nodes when they change
    struct GraphEntry {
        int B = 0, H = 0, W = 0, gated = 0, gated2 = 0; float alpha_s = 0, alpha = 0;
        const void* ws = nullptr; const float* k_dev = nullptr;
        uint64_t extra = 0;                   // sharded forwards: signature of the shard geometry and the peer mappings
        int seen = 0;                         // eager runs with this key before capturing
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t stem_node = nullptr, head_node = nullptr;
        cudaKernelNodeParams stem_p{}, head_p{};
        void* stem_args[16]; void* head_args[16];
        const void* cur_in = nullptr; void* cur_out = nullptr;
        std::map<std::string, Tap> taps; int launches = 0;
        uint64_t last_use = 0;
    };
    std::vector<GraphEntry*> graphs;
    uint64_t use_clock = 0;
    bool use_graphs = true;
    cudaStream_t cap_stream = nullptr;   // capture happens on a private stream (the legacy default stream cannot be captured)
    cudaStream_t cap_stream2 = nullptr;  // second branch of the captured graph: the I and HV halves of a stage run concurrently
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool profiling = false;
    std::vector<cudaEvent_t> events;
    struct Rec { std::string name; double bytes; double flops; };
    std::vector<Rec> recs;
};

namespace {

// ---------------------------------------------------------------- weights ---
std::vector<std::string> all_keys(int variant) {
    std::vector<std::string> keys;
    auto lca = [&](const std::string& p) {
        for (const char* s : {".norm.weight", ".norm.bias", ".gdfn.project_in.weight", ".gdfn.dwconv.weight",
                              ".gdfn.dwconv1.weight", ".gdfn.dwconv2.weight", ".gdfn.project_out.weight",
                              ".ffn.temperature", ".ffn.q.weight", ".ffn.q_dwconv.weight", ".ffn.kv.weight",
                              ".ffn.kv_dwconv.weight", ".ffn.project_out.weight"})
            keys.push_back(p + s);
    };
    for (const char* br : {"HVE", "IE"}) {
        keys.push_back(std::string(br) + "_block0.1.weight");
        for (int n = 1; n <= 3; ++n) {
            keys.push_back(std::string(br) + "_block" + std::to_string(n) + ".prelu.weight");
            keys.push_back(std::string(br) + "_block" + std::to_string(n) + ".down.0.weight");
        }
    }
    for (const char* br : {"HVD", "ID"}) {
        keys.push_back(std::string(br) + "_block0.1.weight");
        for (int n = 1; n <= 3; ++n) {
            const std::string p = std::string(br) + "_block" + std::to_string(n);
            keys.push_back(p + ".prelu.weight");
            keys.push_back(p + ".up_scale.0.weight");
            keys.push_back(p + ".up.weight");
        }
    }
    for (const char* br : {"HV", "I"})
        for (int n = 1; n <= 6; ++n) lca(std::string(br) + "_LCA" + std::to_string(n));
    keys.push_back("trans.density_k");
    if (variant == CIDNET_VARIANT_MSSA)
        for (const char* br : {"hv", "i"})
            for (int n = 1; n <= 3; ++n) keys.push_back(std::string("sa_") + br + std::to_string(n) + ".conv1.weight");
    return keys;
}

int64_t expected_numel(const std::string& key) {
    // shapes of SURVEY App. B, derived from the key
    auto lvl_of_lca = [](int n) { return n <= 3 ? n : 7 - n; };   // LCA1,6 -> 1; 2,5 -> 2; 3,4 -> 3
    if (key == "trans.density_k") return 1;
    if (key.rfind("sa_", 0) == 0 && key.find(".conv1.weight") != std::string::npos) return 2 * 49;   // CIDNet_MSSA.py:17
    if (key.find("prelu.weight") != std::string::npos) return 1;
    if (key == "HVE_block0.1.weight") return 36 * 3 * 9;
    if (key == "IE_block0.1.weight") return 36 * 9;
    if (key == "HVD_block0.1.weight") return 2 * 36 * 9;
    if (key == "ID_block0.1.weight") return 36 * 9;
    const size_t bpos = key.find("_block");
    if (bpos != std::string::npos) {
        const int n = key[bpos + 6] - '0';
        if (key.find(".down.0.weight") != std::string::npos) return (int64_t)kCh[n] * kCh[n - 1] * 9;
        if (key.find(".up_scale.0.weight") != std::string::npos) return (int64_t)kCh[n - 1] * kCh[n] * 9;
        if (key.find(".up.weight") != std::string::npos) return (int64_t)kCh[n - 1] * 2 * kCh[n - 1];
    }
    const size_t lpos = key.find("_LCA");
    if (lpos != std::string::npos) {
        const int n = key[lpos + 4] - '0';
        const int C = kCh[lvl_of_lca(n)], heads = kHeads[lvl_of_lca(n)], h = (int)(C * 2.66);
        if (key.find(".norm.") != std::string::npos) return C;
        if (key.find("project_in") != std::string::npos) return (int64_t)2 * h * C;
        if (key.find(".gdfn.dwconv.weight") != std::string::npos) return (int64_t)2 * h * 9;
        if (key.find(".gdfn.dwconv1") != std::string::npos || key.find(".gdfn.dwconv2") != std::string::npos) return (int64_t)h * 9;
        if (key.find(".gdfn.project_out") != std::string::npos) return (int64_t)C * h;
        if (key.find("temperature") != std::string::npos) return heads;
        if (key.find(".ffn.q.weight") != std::string::npos) return (int64_t)C * C;
        if (key.find(".ffn.q_dwconv") != std::string::npos) return (int64_t)C * 9;
        if (key.find(".ffn.kv.weight") != std::string::npos) return (int64_t)2 * C * C;
        if (key.find(".ffn.kv_dwconv") != std::string::npos) return (int64_t)2 * C * 9;
        if (key.find(".ffn.project_out") != std::string::npos) return (int64_t)C * C;
    }
    return -1;
}

int dev_f32(cidnet_ctx* ctx, float** dst, const std::vector<float>& v) {
    int rc = upload_f32(dst, v);
    if (!rc && *dst) ctx->owned.push_back(*dst);
    return rc;
}
void own(cidnet_ctx* ctx, PackedWeights& p) {
    if (p.w) ctx->owned.push_back(p.w);
    if (p.bias) ctx->owned.push_back(p.bias);
    if (p.wsum) ctx->owned.push_back(p.wsum);
}

// depthwise weight [n][1][3][3] rows [r0, r0+n) -> fp32 [9][pitch] tap-major, zero padded
std::vector<float> dw_tapmajor(const std::vector<float>& w, int r0, int n, int pitch, int dst0 = 0,
                               std::vector<float>* into = nullptr) {
    std::vector<float> local;
    std::vector<float>& out = into ? *into : local;
    if (!into) out.assign((size_t)9 * pitch, 0.f);
    for (int c = 0; c < n; ++c)
        for (int t = 0; t < 9; ++t) out[(size_t)t * pitch + dst0 + c] = w[(size_t)(r0 + c) * 9 + t];
    return out;
}

int build_lca(cidnet_ctx* ctx, const std::string& pfx, int level, LcaWeights* L) {
    auto& R = ctx->raw;
    const int C = kCh[level], heads = kHeads[level], h = (int)(C * 2.66), hp = round_up(h, 16), Cp = act_pitch(C);
    L->C = C; L->Cp = Cp; L->heads = heads; L->h = h; L->hp = hp;
    int rc;
    {
        std::vector<float> wqkv((size_t)9 * 3 * Cp, 0.f);
        dw_tapmajor(R[pfx + ".ffn.q_dwconv.weight"], 0, C, 3 * Cp, 0, &wqkv);
        dw_tapmajor(R[pfx + ".ffn.kv_dwconv.weight"], 0, C, 3 * Cp, Cp, &wqkv);
        dw_tapmajor(R[pfx + ".ffn.kv_dwconv.weight"], C, C, 3 * Cp, 2 * Cp, &wqkv);
        if ((rc = dev_f32(ctx, &L->wqkv, wqkv))) return rc;
    }
    if ((rc = dev_f32(ctx, &L->temp, R[pfx + ".ffn.temperature"]))) return rc;
    if ((rc = dev_f32(ctx, &L->wo, R[pfx + ".ffn.project_out.weight"]))) return rc;
    // geometry of the folded per-image weights (C x C, 1x1)
    PackedWeights& f = L->fold_tmpl;
    f.cin = C; f.taps = 1; f.kchunks = ceil_div(C, 64); f.n_out = C; f.n_img = 1;
    choose_blocking(C, &f.block_n, &f.n_blocks);
    f.n_rows = f.block_n * f.n_blocks;
    // IEL
    const float* lnw = R[pfx + ".norm.weight"].data();
    const float* lnb = R[pfx + ".norm.bias"].data();
    const std::vector<float>& pin = R[pfx + ".gdfn.project_in.weight"];
    std::vector<WeightSegment> segs{{pin.data(), h, 0, lnw, lnb}, {pin.data() + (size_t)h * C, h, hp, lnw, lnb}};
    if ((rc = pack_conv_segments(&L->w_in, segs, C, 1, 2 * hp, true))) return rc;
    own(ctx, L->w_in);
    std::vector<float> dw0((size_t)9 * 2 * hp, 0.f);
    dw_tapmajor(R[pfx + ".gdfn.dwconv.weight"], 0, h, 2 * hp, 0, &dw0);
    dw_tapmajor(R[pfx + ".gdfn.dwconv.weight"], h, h, 2 * hp, hp, &dw0);
    if ((rc = dev_f32(ctx, &L->dw0, dw0))) return rc;
    if ((rc = dev_f32(ctx, &L->dw1, dw_tapmajor(R[pfx + ".gdfn.dwconv1.weight"], 0, h, hp)))) return rc;
    if ((rc = dev_f32(ctx, &L->dw2, dw_tapmajor(R[pfx + ".gdfn.dwconv2.weight"], 0, h, hp)))) return rc;
    if ((rc = pack_conv_weights(&L->w_out, R[pfx + ".gdfn.project_out.weight"].data(), C, h, 1, nullptr, C, nullptr, nullptr))) return rc;
    own(ctx, L->w_out);
    return CIDNET_OK;
}

int build_weights(cidnet_ctx* ctx) {
    auto& R = ctx->raw;
    for (const std::string& k : all_keys(ctx->variant))
        CIDNET_CHECK(R.count(k) && (int64_t)R[k].size() == expected_numel(k), CIDNET_ERR_STATE,
                     "finalize_weights: missing or mis-sized state_dict tensor '" + k + "'");
    int rc;
    ctx->k_host = R["trans.density_k"][0];
    if ((rc = dev_f32(ctx, &ctx->k_dev, R["trans.density_k"]))) return rc;
    {   // stem / head weights, tap-input major fp32
        const auto& whv = R["HVE_block0.1.weight"];   // [36][3][9]
        std::vector<float> a(27 * 36), b(9 * 36), c(9 * 36), d(2 * 9 * 36);
        for (int o = 0; o < 36; ++o)
            for (int i = 0; i < 27; ++i) a[i * 36 + o] = whv[o * 27 + i];
        const auto& wi = R["IE_block0.1.weight"];     // [36][1][9]
        for (int o = 0; o < 36; ++o)
            for (int t = 0; t < 9; ++t) b[t * 36 + o] = wi[o * 9 + t];
        const auto& wid = R["ID_block0.1.weight"];    // [1][36][9]
        for (int ci = 0; ci < 36; ++ci)
            for (int t = 0; t < 9; ++t) c[t * 36 + ci] = wid[ci * 9 + t];
        const auto& whd = R["HVD_block0.1.weight"];   // [2][36][9]
        for (int o = 0; o < 2; ++o)
            for (int ci = 0; ci < 36; ++ci)
                for (int t = 0; t < 9; ++t) d[(o * 9 + t) * 36 + ci] = whd[(o * 36 + ci) * 9 + t];
        if ((rc = dev_f32(ctx, &ctx->stem_whv, a))) return rc;
        if ((rc = dev_f32(ctx, &ctx->stem_wi, b))) return rc;
        if ((rc = dev_f32(ctx, &ctx->head_wi, c))) return rc;
        if ((rc = dev_f32(ctx, &ctx->head_whv, d))) return rc;
        // per-lane B fragments of the stem / head tensor-core kernels (uint2 viewed as 2 floats for the upload helper)
        std::vector<float> fs(3 * 5 * 32 * 2), fh(5 * 3 * 32 * 2);
        pack_stem_bfrag(a.data(), b.data(), reinterpret_cast<uint2*>(fs.data()));
        pack_head_bfrag(c.data(), d.data(), reinterpret_cast<uint2*>(fh.data()));
        float *dfs = nullptr, *dfh = nullptr;
        if ((rc = dev_f32(ctx, &dfs, fs))) return rc;
        if ((rc = dev_f32(ctx, &dfh, fh))) return rc;
        ctx->stem_bfrag = reinterpret_cast<const uint2*>(dfs);
        ctx->head_bfrag = reinterpret_cast<const uint2*>(dfh);
    }
    const char* enc[2] = {"IE", "HVE"};
    const char* dec[2] = {"ID", "HVD"};
    for (int br = 0; br < 2; ++br) {
        for (int n = 1; n <= 3; ++n) {
            const std::string p = std::string(enc[br]) + "_block" + std::to_string(n);
            DownWeights& D = ctx->down[br][n - 1];
            if ((rc = pack_conv_weights(&D.w, R[p + ".down.0.weight"].data(), kCh[n], kCh[n - 1], 9, nullptr, kCh[n], nullptr, nullptr, 128))) return rc;
            own(ctx, D.w);
            D.prelu = R[p + ".prelu.weight"][0];
        }
        for (int n = 3; n >= 1; --n) {
            // NormUpsample(in = kCh[n], out = kCh[n-1]):  1x1(cat[up(conv3(x)), skip]) =
            //   up((Wa o conv3)(x)) + Wb * skip   -- the 1x1 commutes with the bilinear resampling.
            const std::string p = std::string(dec[br]) + "_block" + std::to_string(n);
            UpWeights& U = ctx->up[br][3 - n];
            const int cin = kCh[n], co = kCh[n - 1];
            const auto& w3 = R[p + ".up_scale.0.weight"];   // [co][cin][9]
            const auto& wu = R[p + ".up.weight"];           // [co][2co]
            std::vector<float> comp((size_t)co * cin * 9), wb((size_t)co * co);
            for (int o = 0; o < co; ++o) {
                for (int i = 0; i < cin * 9; ++i) {
                    double s = 0.0;
                    for (int m = 0; m < co; ++m) s += (double)wu[(size_t)o * 2 * co + m] * w3[(size_t)m * cin * 9 + i];
                    comp[(size_t)o * cin * 9 + i] = (float)s;
                }
                for (int m = 0; m < co; ++m) wb[(size_t)o * co + m] = wu[(size_t)o * 2 * co + co + m];
            }
            if ((rc = pack_conv_weights(&U.w3, comp.data(), co, cin, 9, nullptr, co, nullptr, nullptr))) return rc;
            own(ctx, U.w3);
            if ((rc = pack_conv_weights(&U.w1, wb.data(), co, co, 1, nullptr, co, nullptr, nullptr))) return rc;
            own(ctx, U.w1);
            U.prelu = R[p + ".prelu.weight"][0];
        }
    }
    if (ctx->variant == CIDNET_VARIANT_MSSA) {
        const char* sa_br[2] = {"i", "hv"};
        for (int br = 0; br < 2; ++br)
            for (int n = 1; n <= 3; ++n)      // sa_x{n} gates the output of up block n, i.e. a tensor of level n-1
                if ((rc = dev_f32(ctx, &ctx->sa_w[br][n - 1], R[std::string("sa_") + sa_br[br] + std::to_string(n) + ".conv1.weight"]))) return rc;
    }
    for (int l = 1; l <= 3; ++l) {
        if ((rc = pack_identity(&ctx->eye[l], kCh[l]))) return rc;
        own(ctx, ctx->eye[l]);
    }
    for (int n = 1; n <= 6; ++n) {
        const int level = n <= 3 ? n : 7 - n;
        StageWeights& S = ctx->stage[n - 1];
        const std::string pi = "I_LCA" + std::to_string(n), ph = "HV_LCA" + std::to_string(n);
        // I_LCA5 is dead in the reference graph (CIDNet.py:105 vs :109) but live in the MSSA variant (CIDNet_MSSA.py:139,144)
        S.lca[0].live = (n != 5) || ctx->variant == CIDNET_VARIANT_MSSA;
        S.lca[1].live = true;
        if (S.lca[0].live && (rc = build_lca(ctx, pi, level, &S.lca[0]))) return rc;
        if ((rc = build_lca(ctx, ph, level, &S.lca[1]))) return rc;
        const int C = kCh[level], Cp = act_pitch(C);
        const std::string pf[2] = {pi, ph};
        for (int src = 0; src < 2; ++src) {
            // GEMM on tensor `src` (0 = I, 1 = HV): q of LCA[src] (x = own tensor), k,v of LCA[1-src] (y = this tensor)
            const std::string &own_p = pf[src], &oth_p = pf[1 - src];
            std::vector<WeightSegment> segs;
            if (S.lca[src].live)
                segs.push_back({R[own_p + ".ffn.q.weight"].data(), C, 0, R[own_p + ".norm.weight"].data(), R[own_p + ".norm.bias"].data()});
            if (S.lca[1 - src].live) {
                const float* kv = R[oth_p + ".ffn.kv.weight"].data();
                const float* lw = R[oth_p + ".norm.weight"].data();
                const float* lb = R[oth_p + ".norm.bias"].data();
                segs.push_back({kv, C, Cp, lw, lb});
                segs.push_back({kv + (size_t)C * C, C, 2 * Cp, lw, lb});
            }
            if ((rc = pack_conv_segments(&S.qkv[src], segs, C, 1, 3 * Cp, true))) return rc;
            own(ctx, S.qkv[src]);
        }
    }
    return CIDNET_OK;
}

// -------------------------------------------------------------- workspace ---
struct Bump {
    uint8_t* base; int64_t off = 0;
    explicit Bump(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
    template <typename T> T* take(int64_t count) {
        off = (off + 1023) & ~int64_t(1023);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * (int64_t)sizeof(T);
        return p;
    }
};

struct Plan {
    int B, H[4], W[4];
    float* hvi; float* out_hvi;
    act_t *i_enc0, *hv_0, *id1, *hvd1;                     // L0, pitch 40
    act_t *enc_i[4], *enc_hv[4];                           // enc_x[l] = output of block l (level l), l = 1..3
    act_t *lca_i[7], *lca_hv[7];                           // outputs of LCA n = 1..6
    act_t *dec_i[4], *dec_hv[4];                           // dec_x[l] = output of up block l+1 -> level l (l = 1, 2)
    act_t *tup_i[4], *tup_hv[4];                           // low-res pre-composed conv outputs at level l (l = 1..3)
    // per-level scratch for an LCA stage pair
    act_t *qkv[4][2], *qk[4][2], *vdw[4][2], *xp[4][2], *tin[4][2], *g[4][2], *mfold[4][2];
    float* slab;                                           // split-K partials of the stage in flight (cab.cuh)
    float2* sa_stats[2];                                   // MSSA: per-pixel (mean, max) of one up-block output per branch
    float* stat[6];                                        // per stage: reduced raw [Gram | sq | sk] per (problem, image)
    float* statsum;                                        // peer transport: the cross-rank sum of the stage in flight
    int64_t bytes;
};

// Workspace layout.  Long-lived tensors (block0 outputs, the HVI image, block / LCA outputs that serve as skip connections)
// get their own ranges; everything that only lives inside a stage is carved out of two ARENAS (one per branch, so the I and
// HV launches of a pair never share memory), by liveness:
//
//   arena[s], viewed at level l (element offsets per level-l pixel; Cp = channel pitch, hp = padded hidden width):
//     [0, Cp)            xp   = x + CAB(..)          written by attnv, read until project_out
//     [Cp, Cp + 2 hp)    tin  = project_in output    written after the Gram / fold, read by the gate
//       [Cp, 4 Cp)         qkv  (ln_qkv output)        dead once dw3x3 has run   } both inside tin's range: 3 Cp + 2 Cp <= 2 hp
//       [4 Cp, 6 Cp)       qk   (dw3x3 output)         dead once the Gram has run }
//     [Cp + 2 hp, Cp + 3 hp)  g = gate output       written by the gate, read by project_out
//       [3 hp, Cp + 3 hp)      vdw (dw3x3 output)      read by attnv, long before the gate writes g
//   The level-1 arena (the largest) hosts the level-2 and level-3 views side by side (stages 2-5 run while level 1 is idle) and,
//   after stage 6, the last up block's outputs id1 / hvd1.  The spatial-attention statistics of the MSSA variant share the
//   range of the output_hvi tap (the head writes that tap after the last gate has consumed the statistics).
// ~0.85 KB per input pixel instead of the 1.9 KB of the round-1 bump allocation (cfg 4: 13 GB instead of 29 GB).
void make_plan(Plan* P, void* ws, int B, int H, int W) {
    Bump bp(ws);
    P->B = B;
    for (int l = 0; l < 4; ++l) { P->H[l] = H >> l; P->W[l] = W >> l; }
    auto px = [&](int l) { return (int64_t)B * P->H[l] * P->W[l]; };
    P->hvi = bp.take<float>(px(0) * 3);
    P->out_hvi = bp.take<float>(px(0) * 4);                        // 12 B/px tap | 16 B/px of (mean, max) statistics
    P->sa_stats[0] = reinterpret_cast<float2*>(P->out_hvi);
    P->sa_stats[1] = P->sa_stats[0] ? P->sa_stats[0] + px(0) : nullptr;
    P->i_enc0 = bp.take<act_t>(px(0) * 40); P->hv_0 = bp.take<act_t>(px(0) * 40);
    int Cp[4], hp[4];
    for (int l = 1; l <= 3; ++l) { Cp[l] = act_pitch(kCh[l]); hp[l] = round_up((int)(kCh[l] * 2.66), 16); }
    // the two arenas: sized for the level-1 view, which also covers level 2 + level 3 side by side and id1 / hvd1
    int64_t arena_elems = px(1) * (Cp[1] + 3 * hp[1]);
    arena_elems = std::max(arena_elems, px(2) * (Cp[2] + 3 * hp[2]) + px(3) * (Cp[3] + 3 * hp[3]) + 2 * 512);
    arena_elems = std::max(arena_elems, px(0) * 40);
    for (int s = 0; s < 2; ++s) {
        act_t* arena = bp.take<act_t>(arena_elems);
        auto align512 = [](int64_t e) { return (e + 511) & ~int64_t(511); };      // 1 KB
        int64_t base[4] = {0, 0, 0, 0};
        base[3] = align512(px(2) * (Cp[2] + 3 * hp[2]));                       // level 3 behind level 2
        for (int l = 1; l <= 3; ++l) {
            act_t* a0 = arena ? arena + base[l] : nullptr;
            auto at = [&](int64_t per_px) { return a0 ? a0 + px(l) * per_px : nullptr; };
            P->xp[l][s] = at(0);
            P->tin[l][s] = at(Cp[l]);
            P->qkv[l][s] = at(Cp[l]);
            P->qk[l][s] = at(4 * Cp[l]);
            P->g[l][s] = at(Cp[l] + 2 * hp[l]);
            P->vdw[l][s] = at(3 * hp[l]);
        }
        (s == 0 ? P->id1 : P->hvd1) = arena;
    }
    for (int l = 1; l <= 3; ++l) {
        P->enc_i[l] = bp.take<act_t>(px(l) * Cp[l]); P->enc_hv[l] = bp.take<act_t>(px(l) * Cp[l]);
        const int Cup = act_pitch(kCh[l - 1]);
        P->tup_i[l] = bp.take<act_t>(px(l) * Cup); P->tup_hv[l] = bp.take<act_t>(px(l) * Cup);
        if (l <= 2) { P->dec_i[l] = bp.take<act_t>(px(l) * Cp[l]); P->dec_hv[l] = bp.take<act_t>(px(l) * Cp[l]); }
        PackedWeights f; choose_blocking(kCh[l], &f.block_n, &f.n_blocks);
        const int64_t fold_elems = (int64_t)B * f.block_n * f.n_blocks * ceil_div(kCh[l], 64) * 64;
        for (int s = 0; s < 2; ++s) P->mfold[l][s] = bp.take<act_t>(fold_elems);
    }
    for (int n = 1; n <= 6; ++n) {
        const int l = n <= 3 ? n : 7 - n;
        P->lca_i[n] = bp.take<act_t>(px(l) * Cp[l]);
        P->lca_hv[n] = bp.take<act_t>(px(l) * Cp[l]);
    }
    // attention statistics: the split-K slab (re-used by every stage) and the reduced vector of each stage
    P->slab = bp.take<float>((int64_t)gram_max_slab_entries(2 * B) * (8 * 324 + 2 * 144));
    for (int n = 1; n <= 6; ++n) {
        const int l = n <= 3 ? n : 7 - n;
        P->stat[n - 1] = bp.take<float>((int64_t)2 * B * (kHeads[l] * 324 + 2 * Cp[l]));
    }
    P->statsum = bp.take<float>((int64_t)2 * B * (8 * 324 + 2 * 144));
    P->bytes = bp.off + 1024;
}

// every NHWC activation of a plan in a fixed order: index i of one rank's list is the same tensor as index i of another's
std::vector<const void*> plan_tensors(const Plan& P) {
    std::vector<const void*> v = {P.i_enc0, P.hv_0, P.id1, P.hvd1};
    for (int l = 1; l <= 3; ++l) {
        v.push_back(P.enc_i[l]); v.push_back(P.enc_hv[l]); v.push_back(P.tup_i[l]); v.push_back(P.tup_hv[l]);
        if (l <= 2) { v.push_back(P.dec_i[l]); v.push_back(P.dec_hv[l]); }
        for (int s = 0; s < 2; ++s) {
            v.push_back(P.qkv[l][s]); v.push_back(P.qk[l][s]); v.push_back(P.vdw[l][s]); v.push_back(P.xp[l][s]);
            v.push_back(P.tin[l][s]); v.push_back(P.g[l][s]);
        }
    }
    for (int n = 1; n <= 6; ++n) { v.push_back(P.lca_i[n]); v.push_back(P.lca_hv[n]); }
    return v;
}

// ---------------------------------------------------------------- forward ---
// Row-strip sharding of one image over the ranks of a node (cidnet_forward_sharded).  The local image
// = owned rows + `halo` rows of the neighbours on every interior side; all kernels run over the whole
// local image exactly as in the unsharded forward, so every 3x3 stage leaves one more outermost halo
// row invalid.  `margin[t]` = number of halo rows of tensor t (at t's own level) that still hold the
// true values; when a stage needs more than is left, the halo is refreshed from the neighbours.
struct ShardState {
    bool on = false, dry = false;
    int rank = 0, nranks = 1, gH = 0, row0 = 0, halo_top = 0, halo_bot = 0;
    cidnet_halo_fn halo_fn = nullptr; cidnet_allreduce_fn ar_fn = nullptr; void* user = nullptr;
    std::map<const void*, int> margin;
    int halo_calls = 0, allreduce_calls = 0;
    // peer-memory transport (cidnet_forward_sharded_peer): every rank's workspace is mapped into this process
    bool peer = false;
    uint8_t* ws_all[kPeerMaxRanks] = {};       // [r]: this process's mapping of rank r's workspace (header first, plan after it)
    std::vector<Plan> plans;                   // every rank's plan (addresses inside ws_all[r])
    std::vector<cidnet_shard> shards;          // every rank's shard geometry
    int64_t halo_bytes = 0;
};
struct HaloT { const void* p; int level; int pitch; };
static const int kNoLimit = 1 << 28;

struct Fwd {
    cidnet_ctx* ctx; Plan P; cudaStream_t st; int launches = 0;
    ShardState sh;

    // two-branch mode (only while the forward is being captured into a CUDA graph): the I and the HV launch
    // of a pair go to two streams = two parallel graph branches, each on half of the SMs
    bool par = false; cudaStream_t st2 = nullptr;
    cudaStream_t S(int br) const { return (par && br == 1) ? st2 : st; }
    void fork() { if (par) { cudaEventRecord(ctx->ev_fork, st); cudaStreamWaitEvent(st2, ctx->ev_fork, 0); } }
    void join() { if (par) { cudaEventRecord(ctx->ev_join, st2); cudaStreamWaitEvent(st, ctx->ev_join, 0); } }
    bool live() const { return !sh.dry; }
    bool split() const { return sh.on && sh.nranks > 1; }
    int ht(int l) const { return sh.halo_top >> l; }
    int hb(int l) const { return sh.halo_bot >> l; }
    int own_rows(int l) const { return P.H[l] - ht(l) - hb(l); }
    int mg(const void* p) {
        if (!split()) return kNoLimit;
        auto it = sh.margin.find(p);
        return it == sh.margin.end() ? 0 : it->second;
    }
    void setm(const void* p, int m) { if (split()) sh.margin[p] = m < kNoLimit ? m : kNoLimit; }
    // make sure every listed tensor has at least `need` valid halo rows: one halo callback for all of
    // those that do not (the callback refreshes the whole halo -> margin = halo rows of the level)
    int ensure(const std::vector<HaloT>& ts, int need) {
        if (!split()) return CIDNET_OK;
        std::vector<cidnet_halo_req> reqs;
        for (const HaloT& t : ts) {
            if (!t.p || mg(t.p) >= need) continue;
            bool dup = false;
            for (const auto& r : reqs) dup |= (r.base == t.p);
            if (dup) continue;
            cidnet_halo_req r;
            r.base = const_cast<void*>(t.p);
            r.row_bytes = (int64_t)P.W[t.level] * t.pitch * (int64_t)sizeof(act_t);
            r.rows = P.H[t.level]; r.halo_top = ht(t.level); r.halo_bot = hb(t.level); r.reserved = t.level;
            reqs.push_back(r);
            const int full = (sh.halo_top > 0 ? sh.halo_top : sh.halo_bot) >> t.level;
            CIDNET_CHECK(full >= need, CIDNET_ERR_INVALID, "forward_sharded: halo too small for this stage");
            sh.margin[t.p] = full;
        }
        if (reqs.empty()) return CIDNET_OK;
        ++sh.halo_calls;
        if (sh.peer) return peer_halo(reqs);
        const int rc = sh.halo_fn(sh.user, reqs.data(), (int)reqs.size());
        CIDNET_CHECK(rc == 0, CIDNET_ERR_STATE, "forward_sharded: halo exchange callback failed (" + std::to_string(rc) + ")");
        return CIDNET_OK;
    }

    // ---- peer-memory transport -------------------------------------------------------------------------------------
    void peer_sync(PeerSync* ps, bool all_ranks) {
        ps->me = reinterpret_cast<PeerHdr*>(sh.ws_all[sh.rank]);
        ps->rank = sh.rank; ps->npartners = 0;
        for (int r = 0; r < sh.nranks; ++r) {
            if (r == sh.rank || (!all_ranks && r != sh.rank - 1 && r != sh.rank + 1)) continue;
            ps->partner[ps->npartners] = reinterpret_cast<PeerHdr*>(sh.ws_all[r]);
            ps->partner_rank[ps->npartners++] = r;
        }
    }
    // pull the neighbours' boundary rows of the requested tensors into this rank's halo rows (one kernel)
    int peer_halo(const std::vector<cidnet_halo_req>& reqs) {
        PeerHaloArgs a; memset(&a, 0, sizeof a);
        peer_sync(&a.sync, false);
        const std::vector<const void*> mine = plan_tensors(P);
        for (const cidnet_halo_req& q : reqs) {
            size_t idx = 0;
            while (idx < mine.size() && mine[idx] != q.base) ++idx;
            CIDNET_CHECK(idx < mine.size(), CIDNET_ERR_STATE, "forward_sharded_peer: halo request for an unknown tensor");
            uint8_t* base = reinterpret_cast<uint8_t*>(q.base);
            for (int dir = 0; dir < 2; ++dir) {               // 0: rows from rank - 1 (top halo), 1: from rank + 1 (bottom halo)
                const int h = dir == 0 ? q.halo_top : q.halo_bot;
                if (h == 0) continue;
                const int nb = dir == 0 ? sh.rank - 1 : sh.rank + 1;
                const cidnet_shard& ns = sh.shards[nb];
                // the neighbour's local image at this tensor's level: [its top halo | its owned rows | its bottom halo]
                const int lvl_shift = q.reserved;             // the tensor's level (set by ensure())
                const int n_top = ((nb > 0 ? ns.halo : 0) >> lvl_shift), n_own = (ns.row_end - ns.row_begin) >> lvl_shift;
                const int src_row = dir == 0 ? n_top + n_own - h : n_top;
                const int dst_row = dir == 0 ? 0 : q.rows - h;
                const uint8_t* nbase = reinterpret_cast<const uint8_t*>(plan_tensors(sh.plans[nb])[idx]);
                CIDNET_CHECK(a.njobs < kPeerMaxJobs, CIDNET_ERR_STATE, "forward_sharded_peer: too many halo jobs");
                a.job[a.njobs++] = PeerCopyJob{base + (int64_t)dst_row * q.row_bytes, nbase + (int64_t)src_row * q.row_bytes,
                                               (long long)h * q.row_bytes};
                sh.halo_bytes += (int64_t)h * q.row_bytes;
            }
        }
        ++launches;
        return live() ? launch_peer_halo(a, st) : CIDNET_OK;
    }
    // sum the ranks' partial statistics vectors (rank order) into P.statsum
    int peer_allreduce(int stage, int count) {
        PeerReduceArgs a; memset(&a, 0, sizeof a);
        peer_sync(&a.sync, true);
        for (int r = 0; r < sh.nranks; ++r) a.src[r] = sh.plans[r].stat[stage];
        a.dst = P.statsum; a.count = count; a.nranks = sh.nranks;
        ++launches;
        return live() ? launch_peer_allreduce(a, st) : CIDNET_OK;
    }

    void tap(const std::string& name, const void* p, int C, int l, int pitch, bool f32 = false) {
        ctx->taps[name] = Tap{p, C, P.H[l], P.W[l], pitch, f32};
    }
    // profiling mark: called right before every kernel launch (br = the graph branch / stream the launch goes to).
    // Launch i owns events 2i (recorded before it) and 2i+1 (recorded after it, on ITS stream, when the next mark
    // comes -- nothing but fork / join event edges is enqueued in between).  While profiling, the forward runs eagerly
    // with the same two-stream split as the captured graph, so a kernel's time is the one it has in that schedule.
    cudaStream_t pending = nullptr; bool has_pending = false;
    void close_pending() {
        if (has_pending) cudaEventRecord(ctx->events[2 * ctx->recs.size() - 1], pending);
        has_pending = false;
    }
    void mark(const std::string& name, double bytes, double flops, int br = 0) {
        ++launches;
        if (!ctx->profiling) return;
        close_pending();
        const size_t i = ctx->recs.size();
        if (ctx->events.size() < 2 * i + 2) ctx->events.resize(2 * i + 2, nullptr);
        for (size_t j = 2 * i; j < 2 * i + 2; ++j)
            if (ctx->events[j] == nullptr) cudaEventCreate(&ctx->events[j]);
        cudaEventRecord(ctx->events[2 * i], S(br));
        ctx->recs.push_back({name, bytes, flops});
        pending = S(br); has_pending = true;
    }
    void finish_marks() {
        if (ctx->profiling) close_pending();
    }
    int gemm(ConvGemmLaunch& L, const std::string& name, int br = 0, bool paired = true) {
        const PackedWeights& w = *L.wt;
        const double px_in = (double)L.B * L.H * L.W;
        const double px_out = L.mode == EPI_DOWN ? px_in / 4 : px_in;
        double bytes = px_in * w.cin * 2 + px_out * w.n_out * 2 + (double)w.n_out * w.cin * w.taps * 2 * (w.n_img > 1 ? L.B : 1);
        if (L.in2) bytes += px_out * L.wt2->cin * 2;
        if (L.up) bytes += px_out / 4 * w.n_out * 2;
        mark(name, bytes, 2.0 * px_in * w.n_out * w.cin * w.taps, br);
        if (par && paired) L.max_ctas = device_sm_count() / 2;
        return live() ? launch_conv_gemm(L, S(br)) : CIDNET_OK;
    }

    int down(int br, int n, const act_t* in, act_t* out) {   // level n-1 -> n
        const DownWeights& D = ctx->down[br][n - 1];
        ConvGemmLaunch L;
        L.mode = EPI_DOWN; L.in = in; L.B = P.B; L.H = P.H[n - 1]; L.W = P.W[n - 1]; L.in_pitch = act_pitch(kCh[n - 1]);
        L.wt = &D.w; L.out = out; L.out_pitch = act_pitch(kCh[n]); L.prelu = D.prelu;
        if (sh.on) { L.gH = sh.gH >> (n - 1); L.grow = sh.row0 >> (n - 1); }
        // conv rows 2y, 2y+1 feed output row y: an input margin m leaves (m - 1) / 2 valid output halo rows
        CIDNET_CHECK(mg(in) >= 1, CIDNET_ERR_STATE, "forward_sharded: down block without a valid input halo");
        setm(out, mg(in) >= kNoLimit ? kNoLimit : (mg(in) - 1) / 2);
        return gemm(L, "down" + std::to_string(n) + ".conv3x3_bilinear_prelu", br);
    }
    int up(int br, int n, const act_t* x, const act_t* skip, act_t* t, act_t* out) {   // level n -> n-1
        const UpWeights& U = ctx->up[br][3 - n];
        ConvGemmLaunch A;
        A.mode = EPI_STORE; A.in = x; A.B = P.B; A.H = P.H[n]; A.W = P.W[n]; A.in_pitch = act_pitch(kCh[n]);
        A.wt = &U.w3; A.out = t; A.out_pitch = act_pitch(kCh[n - 1]);
        // output row Y reads low-res rows floor(Y*r), +1 with r < 1/2: a low-res margin m_t gives
        // 2 (m_t - 1) valid output halo rows; m_t = margin(x) - 1 after the 3x3
        CIDNET_CHECK(mg(x) >= 2, CIDNET_ERR_STATE, "forward_sharded: up block without a valid input halo");
        setm(t, mg(x) - 1);
        setm(out, std::min(mg(x) >= kNoLimit ? kNoLimit : 2 * (mg(x) - 2), mg(skip)));
        int rc = gemm(A, "up" + std::to_string(n) + ".conv3x3_composed", br);
        if (rc) return rc;
        ConvGemmLaunch Bq;
        Bq.mode = EPI_UP; Bq.in = skip; Bq.B = P.B; Bq.H = P.H[n - 1]; Bq.W = P.W[n - 1]; Bq.in_pitch = act_pitch(kCh[n - 1]);
        Bq.flat = true; Bq.wt = &U.w1; Bq.out = out; Bq.out_pitch = act_pitch(kCh[n - 1]);
        Bq.up = t; Bq.up_pitch = act_pitch(kCh[n - 1]); Bq.prelu = U.prelu;
        // MSSA variant: the SpatialAttention gate that follows needs the channel mean / max of this output -- the epilogue
        // thread holds its pixel's whole channel vector, so the statistics cost no extra pass over the tensor
        if (ctx->variant == CIDNET_VARIANT_MSSA) Bq.sa_stats = P.sa_stats[br];
        if (sh.on) { Bq.gH = sh.gH >> (n - 1); Bq.grow = sh.row0 >> (n - 1); }
        return gemm(Bq, "up" + std::to_string(n) + ".skip1x1_bilinear_prelu", br);
    }

    // MSSA variant: SpatialAttention gates on the I / HV outputs of an up-block pair, in place (level l tensors).
    // The 7x7 conv over the (mean, max) map needs three valid halo rows (row-strip sharding).
    int sa_pair(int l, act_t* xi, act_t* xhv) {
        if (ctx->variant != CIDNET_VARIANT_MSSA) return CIDNET_OK;
        const int C = kCh[l], Cp = act_pitch(C);
        int rc;
        if ((rc = ensure({{xi, l, Cp}, {xhv, l, Cp}}, 3))) return rc;
        SaArgs a;
        a.x[0] = xi; a.x[1] = xhv; a.w[0] = ctx->sa_w[0][l]; a.w[1] = ctx->sa_w[1][l];
        a.stats[0] = P.sa_stats[0]; a.stats[1] = P.sa_stats[1];
        a.B = P.B; a.H = P.H[l]; a.W = P.W[l]; a.C = C; a.pitch = Cp; a.nprob = 2;
        const double px = 2.0 * P.B * P.H[l] * P.W[l];
        // (mean, max) per pixel were written by the up block's epilogue (ConvGemmLaunch::sa_stats)
        mark("sa" + std::to_string(l + 1) + ".conv7x7_sigmoid_gate", px * (4.0 * C + 8), px * (2.0 * 98 + C));
        if (live() && (rc = launch_sa_gate(a, st))) return rc;
        if (split()) { setm(xi, mg(xi) - 3); setm(xhv, mg(xhv) - 3); }
        return CIDNET_OK;
    }

    // one LCA stage: I_LCA(x_i, x_hv) and HV_LCA(x_hv, x_i)   (net/LCA.py:78-81, 90-93)
    int lca_stage(int n, const act_t* x_i, const act_t* x_hv, act_t* out_i, act_t* out_hv) {
        const int l = n <= 3 ? n : 7 - n;
        StageWeights& S = ctx->stage[n - 1];
        const int C = kCh[l], Cp = act_pitch(C), heads = kHeads[l];
        const int H = P.H[l], W = P.W[l];
        const act_t* x[2] = {x_i, x_hv};
        act_t* out[2] = {out_i, out_hv};
        int rc;
        // the depthwise 3x3 of q|k|v needs one valid halo row of both inputs
        if ((rc = ensure({{x_i, l, Cp}, {x_hv, l, Cp}}, 1))) return rc;
        const int m_dw = split() ? std::min(mg(x_i), mg(x_hv)) - 1 : kNoLimit;
        // 1. LayerNorm + q / kv 1x1 of both branches: one GEMM per input tensor
        fork();
        for (int s = 0; s < 2; ++s) {
            ConvGemmLaunch L;
            L.mode = EPI_LN; L.in = x[s]; L.B = P.B; L.H = H; L.W = W; L.in_pitch = Cp; L.flat = true;
            L.wt = &S.qkv[s]; L.out = P.qkv[l][s]; L.out_pitch = 3 * Cp;
            if ((rc = gemm(L, "L" + std::to_string(l) + ".ln_qkv_1x1", s))) return rc;
        }
        join();
        // 2. depthwise 3x3 of [q | k | v], then the Gram (+ sum q^2, sum k^2) on the tensor cores
        int probs[2], np = 0;
        for (int s = 0; s < 2; ++s) if (S.lca[s].live) probs[np++] = s;
        const int E = heads * 324 + 2 * Cp;                    // floats of one [Gram | sq | sk] vector
        int nsplit = 1;
        {
            Dw3Args a; memset(&a, 0, sizeof a);
            GramLaunch gl; memset(&gl, 0, sizeof gl);
            for (int i = 0; i < np; ++i) {
                const int s = probs[i];
                a.src[i][0] = P.qkv[l][s];                     // q of LCA s comes from its own tensor
                a.src[i][1] = P.qkv[l][1 - s] + Cp;            // k, v from the sibling tensor
                a.src[i][2] = P.qkv[l][1 - s] + 2 * Cp;
                a.dst_qk[i] = P.qk[l][s]; a.dst_v[i] = P.vdw[l][s];
                a.w[i] = S.lca[s].wqkv;
                gl.q[i] = P.qk[l][s]; gl.k[i] = P.qk[l][s] + Cp;
            }
            a.src_pitch = 3 * Cp; a.B = P.B; a.H = H; a.W = W;
            a.nv = 3 * Cp / 8; a.seg_vecs = Cp / 8; a.nprob = np;
            // sharded: sum q^2, sum k^2 and the Gram run over the OWNED rows only (partial sums, all-reduced below)
            const int own0 = sh.on ? ht(l) : 0, own_n = sh.on ? own_rows(l) : H;
            for (int i = 0; i < np; ++i) { gl.q[i] += (long long)own0 * W * 2 * Cp; gl.k[i] += (long long)own0 * W * 2 * Cp; }
            mark("L" + std::to_string(l) + ".cab_dw3x3_qkv", (double)np * P.B * H * W * 12.0 * C, (double)np * P.B * H * W * 2.0 * 27 * C);
            if (live() && (rc = launch_dw3(a, st))) return rc;
            gl.pitch = 2 * Cp; gl.B = P.B; gl.H = own_n; gl.W = W; gl.C = C; gl.heads = heads; gl.nprob = np;
            gl.img_stride_px = (long long)H * W;
            gl.slab = P.slab;
            mark("L" + std::to_string(l) + ".cab_gram_tc", (double)np * P.B * own_n * W * 4.0 * C, (double)np * P.B * own_n * W * 2.0 * C * C);
            if (live() && (rc = launch_gram(gl, st, &nsplit))) return rc;
            if (split()) {
                // this rank's partial [Gram | sum q^2 | sum k^2] of the live problems (fixed-order sum of its slab), then ONE
                // all-reduce (sum, fp32) across the ranks; normalisation, temperature and softmax then run identically everywhere
                ++launches;
                if (live() && (rc = launch_cab_reduce(P.slab, P.stat[n - 1], nsplit, E, np * P.B, st))) return rc;
                ++sh.allreduce_calls;
                if (sh.peer) {
                    if ((rc = peer_allreduce(n - 1, np * P.B * E))) return rc;
                } else {
                    const int arc = sh.ar_fn(sh.user, P.stat[n - 1], (int64_t)np * P.B * E);
                    CIDNET_CHECK(arc == 0, CIDNET_ERR_STATE, "forward_sharded: all-reduce callback failed (" + std::to_string(arc) + ")");
                }
            }
        }
        // 3. fixed-order sum of the split-K partials, normalise + temperature + softmax + fold into project_out
        {
            CabFoldArgs f; memset(&f, 0, sizeof f);
            for (int i = 0; i < np; ++i) {
                const int s = probs[i];
                f.temp[i] = S.lca[s].temp; f.wo[i] = S.lca[s].wo; f.m_out[i] = P.mfold[l][s];
            }
            if (split()) { f.slab = sh.peer ? P.statsum : P.stat[n - 1]; f.nsplit = 1; f.raw_out = nullptr; }
            else         { f.slab = P.slab; f.nsplit = nsplit; f.raw_out = P.stat[n - 1]; }
            const PackedWeights& t = S.lca[probs[0]].fold_tmpl;
            f.B = P.B; f.C = C; f.Cp = Cp; f.heads = heads; f.nprob = np; f.n_rows = t.n_rows; f.kt = t.ktot();
            mark("L" + std::to_string(l) + ".cab_softmax_fold", (double)np * P.B * (C * C * 6.0 + 4.0 * nsplit * E), (double)np * P.B * 36.0 * C * C);
            if (live() && (rc = launch_cab_fold(f, st))) return rc;
        }
        const bool both = np == 2;        // stage 5 has a single live problem: it keeps all the SMs
        fork();
        for (int i = 0; i < np; ++i) {
            const int s = probs[i];
            LcaWeights& Lw = S.lca[s];
            // 4. x' = x + (W_o * blockdiag(attn_b)) v      (per-image 1x1)
            PackedWeights fw = Lw.fold_tmpl; fw.w = P.mfold[l][s]; fw.n_img = P.B > 1 ? P.B : 1;
            ConvGemmLaunch A;
            A.mode = EPI_STORE; A.in = P.vdw[l][s]; A.B = P.B; A.H = H; A.W = W; A.in_pitch = Cp; A.flat = true;
            A.wt = &fw; A.dynamic_weights = true; A.out = P.xp[l][s]; A.out_pitch = Cp; A.in2 = x[s]; A.in2_pitch = Cp; A.wt2 = &ctx->eye[l];
            if (P.B == 1) fw.n_img = 1;
            if ((rc = gemm(A, "L" + std::to_string(l) + ".cab_attnv_proj_res", s, both))) return rc;
            setm(P.xp[l][s], m_dw);
        }
        // the IEL gate chains two depthwise 3x3: two valid halo rows of x' (project_in is recomputed on them).  The exchange
        // is issued on the main stream: both branches must have produced their x' first
        if (split()) join();
        if ((rc = ensure({{S.lca[0].live ? P.xp[l][0] : nullptr, l, Cp}, {P.xp[l][1], l, Cp}}, 2))) return rc;
        if (split()) fork();
        for (int i = 0; i < np; ++i) {
            const int s = probs[i];
            LcaWeights& Lw = S.lca[s];
            // 5. LayerNorm + project_in
            ConvGemmLaunch Bq;
            Bq.mode = EPI_LN; Bq.in = P.xp[l][s]; Bq.B = P.B; Bq.H = H; Bq.W = W; Bq.in_pitch = Cp; Bq.flat = true;
            Bq.wt = &Lw.w_in; Bq.out = P.tin[l][s]; Bq.out_pitch = 2 * Lw.hp;
            if ((rc = gemm(Bq, "L" + std::to_string(l) + ".ln_iel_project_in", s, both))) return rc;
        }
        join();
        // 6. IEL gate (dw 3x3 -> dw 3x3 + tanh + residual -> product)
        {
            IelGateArgs g; memset(&g, 0, sizeof g);
            for (int i = 0; i < np; ++i) {
                const int s = probs[i];
                g.t[i] = P.tin[l][s]; g.g[i] = P.g[l][s];
                g.w0[i] = S.lca[s].dw0; g.w1[i] = S.lca[s].dw1; g.w2[i] = S.lca[s].dw2;
            }
            g.B = P.B; g.H = H; g.W = W; g.hp = S.lca[probs[0]].hp; g.nprob = np;
            mark("L" + std::to_string(l) + ".iel_gate", (double)np * P.B * H * W * 6.0 * S.lca[probs[0]].h, (double)np * P.B * H * W * 2.0 * 36 * S.lca[probs[0]].h);
            if (live() && (rc = launch_iel_gate(g, st))) return rc;
        }
        // 7. project_out (+ residual for I_LCA only)
        fork();
        for (int i = 0; i < np; ++i) {
            const int s = probs[i];
            LcaWeights& Lw = S.lca[s];
            ConvGemmLaunch Cq;
            Cq.mode = EPI_STORE; Cq.in = P.g[l][s]; Cq.B = P.B; Cq.H = H; Cq.W = W; Cq.in_pitch = Lw.hp; Cq.flat = true;
            Cq.wt = &Lw.w_out; Cq.out = out[s]; Cq.out_pitch = Cp;
            if (s == 0) { Cq.in2 = P.xp[l][s]; Cq.in2_pitch = Cp; Cq.wt2 = &ctx->eye[l]; }
            if ((rc = gemm(Cq, "L" + std::to_string(l) + ".iel_project_out", s, both))) return rc;
            setm(out[s], mg(P.xp[l][s]) >= kNoLimit ? kNoLimit : mg(P.xp[l][s]) - 2);
            tap(std::string(s == 0 ? "I_LCA" : "HV_LCA") + std::to_string(n), out[s], C, l, Cp);
        }
        join();
        return CIDNET_OK;
    }

    // 8-bit I/O fused into the stem's load and the head's store (cidnet_forward_u8): rgb_in / rgb_out are then uint8 HWC
    struct { bool on = false; int h = 0, w = 0; float gamma = 1.f; } u8;

    int run(const void* rgb_in, void* rgb_out, const float* k_dev, int gated, float alpha_s, int gated2, float alpha) {
        int rc;
        StemArgs sa{rgb_in, P.hvi, P.i_enc0, P.hv_0, ctx->stem_whv, ctx->stem_wi, k_dev ? k_dev : ctx->k_dev,
                    ctx->k_host, P.B, P.H[0], P.W[0], 40, ctx->stem_bfrag};
        if (u8.on) { sa.in_u8 = 1; sa.h_src = u8.h; sa.w_src = u8.w; sa.gamma = u8.gamma; }
        mark("L0.stem_hvit_block0", (double)P.B * P.H[0] * P.W[0] * (u8.on ? 159.0 : 168.0), (double)P.B * P.H[0] * P.W[0] * 2.0 * 1296);
        if (live() && (rc = launch_stem(sa, st))) return rc;
        // the local input image carries the neighbours' rows: the replicate-padded 3x3 spoils the outermost one
        setm(P.i_enc0, (sh.halo_top > 0 ? sh.halo_top : sh.halo_bot) - 1);
        setm(P.hv_0, (sh.halo_top > 0 ? sh.halo_top : sh.halo_bot) - 1);
        tap("hvi", P.hvi, 3, 0, 0, true); tap("i_enc0", P.i_enc0, 36, 0, 40); tap("hv_0", P.hv_0, 36, 0, 40);
        if ((rc = ensure({{P.i_enc0, 0, 40}, {P.hv_0, 0, 40}}, 1))) return rc;
        fork();
        if ((rc = down(0, 1, P.i_enc0, P.enc_i[1]))) return rc;
        if ((rc = down(1, 1, P.hv_0, P.enc_hv[1]))) return rc;
        join();
        tap("i_enc1", P.enc_i[1], 36, 1, 40); tap("hv_1", P.enc_hv[1], 36, 1, 40);
        if ((rc = lca_stage(1, P.enc_i[1], P.enc_hv[1], P.lca_i[1], P.lca_hv[1]))) return rc;
        if ((rc = ensure({{P.lca_i[1], 1, 40}, {P.lca_hv[1], 1, 40}}, 1))) return rc;
        fork();
        if ((rc = down(0, 2, P.lca_i[1], P.enc_i[2]))) return rc;
        if ((rc = down(1, 2, P.lca_hv[1], P.enc_hv[2]))) return rc;
        join();
        tap("i_enc2", P.enc_i[2], 72, 2, 72); tap("hv_2", P.enc_hv[2], 72, 2, 72);
        if ((rc = lca_stage(2, P.enc_i[2], P.enc_hv[2], P.lca_i[2], P.lca_hv[2]))) return rc;
        // block3 consumes the PRE-LCA2 tensors (CIDNet.py:94-95)
        if ((rc = ensure({{P.enc_i[2], 2, 72}, {P.enc_hv[2], 2, 72}}, 1))) return rc;
        fork();
        if ((rc = down(0, 3, P.enc_i[2], P.enc_i[3]))) return rc;
        if ((rc = down(1, 3, P.enc_hv[2], P.enc_hv[3]))) return rc;
        join();
        tap("i_enc3", P.enc_i[3], 144, 3, 144); tap("hv_3", P.enc_hv[3], 144, 3, 144);
        if ((rc = lca_stage(3, P.enc_i[3], P.enc_hv[3], P.lca_i[3], P.lca_hv[3]))) return rc;
        // LCA4: both consume the LCA3 outputs (HV_LCA4 sees i_enc4, CIDNet.py:101)
        if ((rc = lca_stage(4, P.lca_i[3], P.lca_hv[3], P.lca_i[4], P.lca_hv[4]))) return rc;
        if ((rc = ensure({{P.lca_hv[4], 3, 144}, {P.lca_i[4], 3, 144}}, 2))) return rc;
        fork();
        if ((rc = up(1, 3, P.lca_hv[4], P.lca_hv[2], P.tup_hv[3], P.dec_hv[2]))) return rc;
        if ((rc = up(0, 3, P.lca_i[4], P.lca_i[2], P.tup_i[3], P.dec_i[2]))) return rc;
        join();
        if ((rc = sa_pair(2, P.dec_i[2], P.dec_hv[2]))) return rc;
        tap("hvd3", P.dec_hv[2], 72, 2, 72); tap("id3", P.dec_i[2], 72, 2, 72);
        // stage 5: I_LCA5 is dead in the base graph, only HV_LCA5(hv_3, i_dec3); the MSSA variant runs both and
        // feeds ID_block2 with I_LCA5's output (CIDNet_MSSA.py:139-146)
        const bool mssa = ctx->variant == CIDNET_VARIANT_MSSA;
        act_t* i_dec2_in = mssa ? P.lca_i[5] : P.dec_i[2];
        if ((rc = lca_stage(5, P.dec_i[2], P.dec_hv[2], mssa ? P.lca_i[5] : nullptr, P.lca_hv[5]))) return rc;
        if ((rc = ensure({{P.lca_hv[5], 2, 72}, {i_dec2_in, 2, 72}}, 2))) return rc;
        fork();
        if ((rc = up(1, 2, P.lca_hv[5], P.lca_hv[1], P.tup_hv[2], P.dec_hv[1]))) return rc;
        if ((rc = up(0, 2, i_dec2_in, P.lca_i[1], P.tup_i[2], P.dec_i[1]))) return rc;   // base graph: takes i_dec3 (:109)
        join();
        if ((rc = sa_pair(1, P.dec_i[1], P.dec_hv[1]))) return rc;
        tap("hvd2", P.dec_hv[1], 36, 1, 40); tap("id2", P.dec_i[1], 36, 1, 40);
        if ((rc = lca_stage(6, P.dec_i[1], P.dec_hv[1], P.lca_i[6], P.lca_hv[6]))) return rc;
        if ((rc = ensure({{P.lca_i[6], 1, 40}, {P.lca_hv[6], 1, 40}}, 2))) return rc;
        fork();
        if ((rc = up(0, 1, P.lca_i[6], P.i_enc0, P.tup_i[1], P.id1))) return rc;
        if ((rc = up(1, 1, P.lca_hv[6], P.hv_0, P.tup_hv[1], P.hvd1))) return rc;
        join();
        if ((rc = sa_pair(0, P.id1, P.hvd1))) return rc;
        tap("id1", P.id1, 36, 0, 40); tap("hvd1", P.hvd1, 36, 0, 40);
        if ((rc = ensure({{P.id1, 0, 40}, {P.hvd1, 0, 40}}, 1))) return rc;
        HeadArgs ha{P.id1, P.hvd1, P.hvi, rgb_out, P.out_hvi, ctx->head_wi, ctx->head_whv,
                    k_dev ? k_dev : ctx->k_dev, ctx->k_host, alpha_s, alpha, gated, gated2, P.B, P.H[0], P.W[0], 40,
                    ctx->head_bfrag};
        if (u8.on) { ha.out_u8 = 1; ha.h_dst = u8.h; ha.w_dst = u8.w; }
        mark("L0.head_block0_phvit", (double)P.B * P.H[0] * P.W[0] * (u8.on ? 159.0 : 168.0), (double)P.B * P.H[0] * P.W[0] * 2.0 * 972);
        if (live() && (rc = launch_head(ha, st))) return rc;
        finish_marks();
        tap("out_hvi", P.out_hvi, 3, 0, 0, true);
        return CIDNET_OK;
    }
};

}  // namespace

// ------------------------------------------------------------------ C ABI ---
extern "C" int cidnet_create(cidnet_ctx** out, int device) {
    CIDNET_CHECK(out != nullptr, CIDNET_ERR_INVALID, "create: null out pointer");
    *out = nullptr;
    int ndev = 0;
    CIDNET_CUDA_OK(cudaGetDeviceCount(&ndev));
    CIDNET_CHECK(device >= 0 && device < ndev, CIDNET_ERR_INVALID, "create: no such CUDA device");
    cudaDeviceProp prop;
    CIDNET_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    CIDNET_CHECK(prop.major == 10, CIDNET_ERR_ARCH,
                 std::string("create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                     std::to_string(prop.minor) + "; this library contains sm_100a code only (no fallback)");
    cidnet_ctx* c = new cidnet_ctx();
    c->device = device;
    if (getenv("CIDNET_NO_GRAPH")) c->use_graphs = false;   // e.g. under ncu: profile plain launches
    *out = c;
    return CIDNET_OK;
}

static void release_device(cidnet_ctx* ctx) {
    for (auto* g : ctx->graphs) {
        if (g->exec) cudaGraphExecDestroy(g->exec);
        if (g->graph) cudaGraphDestroy(g->graph);
        delete g;
    }
    ctx->graphs.clear();
    for (void* p : ctx->owned) cudaFree(p);
    ctx->owned.clear();
    ctx->finalized = false;
}

extern "C" int cidnet_destroy(cidnet_ctx* ctx) {
    if (!ctx) return CIDNET_OK;
    cudaSetDevice(ctx->device);
    release_device(ctx);
    for (cudaEvent_t e : ctx->events) if (e) cudaEventDestroy(e);
    if (ctx->cap_stream) cudaStreamDestroy(ctx->cap_stream);
    if (ctx->cap_stream2) cudaStreamDestroy(ctx->cap_stream2);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
    return CIDNET_OK;
}

extern "C" int cidnet_set_variant(cidnet_ctx* ctx, int variant) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "set_variant: null ctx");
    CIDNET_CHECK(variant == CIDNET_VARIANT_BASE || variant == CIDNET_VARIANT_MSSA, CIDNET_ERR_INVALID, "set_variant: unknown variant");
    if (variant != ctx->variant) ctx->finalized = false;
    ctx->variant = variant;
    return CIDNET_OK;
}

extern "C" int cidnet_set_weight(cidnet_ctx* ctx, const char* key, const float* host, int64_t numel) {
    CIDNET_CHECK(ctx && key && host, CIDNET_ERR_INVALID, "set_weight: null argument");
    const int64_t want = expected_numel(key);
    CIDNET_CHECK(want > 0, CIDNET_ERR_INVALID, std::string("set_weight: unexpected key '") + key + "'");
    CIDNET_CHECK(want == numel, CIDNET_ERR_INVALID,
                 std::string("set_weight: size mismatch for '") + key + "': got " + std::to_string(numel) +
                     ", expected " + std::to_string(want));
    ctx->raw[key].assign(host, host + numel);
    ctx->finalized = false;
    return CIDNET_OK;
}

extern "C" int cidnet_finalize_weights(cidnet_ctx* ctx) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "finalize_weights: null ctx");
    CIDNET_CUDA_OK(cudaSetDevice(ctx->device));
    CIDNET_CUDA_OK(cudaDeviceSynchronize());   // previous packed weights may still be in use
    release_device(ctx);
    for (int i = 0; i < 6; ++i) ctx->stage[i] = StageWeights();
    int rc = build_weights(ctx);
    if (rc) { release_device(ctx); return rc; }
    ctx->finalized = true;
    return CIDNET_OK;
}

extern "C" int64_t cidnet_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0 || H % 8 || W % 8) return 0;
    Plan P;
    make_plan(&P, nullptr, B, H, W);
    return P.bytes;
}

// One forward of an already planned executor `f` (unsharded, or a row strip with the peer-memory transport): replayed as a
// CUDA graph when possible -- one entry per (shape, workspace, flags, extra); the input / output image pointers are patched
// into the stem / head kernel nodes when they change.  Skipped while profiling, inside somebody else's capture, or when
// graphs are disabled: then the launches go out eagerly.
static int run_forward_cached(cidnet_ctx* ctx, Fwd& f, uint64_t extra, const void* rgb_in, void* rgb_out, const float* k_dev,
                              int gated, float alpha_s, int gated2, float alpha, cudaStream_t st) {
    const int B = f.P.B, H = f.P.H[0], W = f.P.W[0];
    const void* workspace = f.P.hvi;                       // first tensor of the plan: identifies the workspace
    cidnet_ctx::GraphEntry* ge = nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (ctx->use_graphs && !ctx->profiling && cudaStreamIsCapturing(st, &cap) == cudaSuccess &&
        cap == cudaStreamCaptureStatusNone) {
        for (auto* g : ctx->graphs)
            if (g->B == B && g->H == H && g->W == W && g->ws == workspace && g->k_dev == k_dev && g->gated == gated &&
                g->gated2 == gated2 && g->alpha_s == alpha_s && g->alpha == alpha && g->extra == extra) { ge = g; break; }
        if (!ge) {
            if (ctx->graphs.size() >= 8) {                    // evict the least recently used entry
                size_t lru = 0;
                for (size_t i = 1; i < ctx->graphs.size(); ++i) if (ctx->graphs[i]->last_use < ctx->graphs[lru]->last_use) lru = i;
                cidnet_ctx::GraphEntry* old = ctx->graphs[lru];
                if (old->exec) cudaGraphExecDestroy(old->exec);
                if (old->graph) cudaGraphDestroy(old->graph);
                delete old;
                ctx->graphs.erase(ctx->graphs.begin() + lru);
            }
            ge = new cidnet_ctx::GraphEntry();
            ge->B = B; ge->H = H; ge->W = W; ge->ws = workspace; ge->k_dev = k_dev; ge->gated = gated; ge->gated2 = gated2;
            ge->alpha_s = alpha_s; ge->alpha = alpha; ge->extra = extra;
            ctx->graphs.push_back(ge);
        }
        ge->last_use = ++ctx->use_clock;
    }
    if (ge && ge->exec) {
        // replay; patch the external image pointers if they moved
        if (ge->cur_in != rgb_in) {
            ge->cur_in = rgb_in;
            ge->stem_args[kStemArgIn] = &ge->cur_in;
            ge->stem_p.kernelParams = ge->stem_args;
            CIDNET_CUDA_OK(cudaGraphExecKernelNodeSetParams(ge->exec, ge->stem_node, &ge->stem_p));
        }
        if (ge->cur_out != rgb_out) {
            ge->cur_out = rgb_out;
            ge->head_args[kHeadArgOut] = &ge->cur_out;
            ge->head_p.kernelParams = ge->head_args;
            CIDNET_CUDA_OK(cudaGraphExecKernelNodeSetParams(ge->exec, ge->head_node, &ge->head_p));
        }
        CIDNET_CUDA_OK(cudaGraphLaunch(ge->exec, st));
        ctx->taps = ge->taps;
        ctx->launches = ge->launches;
        return CIDNET_OK;
    }
    const bool capture = ge && ge->seen >= 1;     // first call with a new key runs eagerly (one-time setup)
    if (ge) ge->seen++;
    ctx->taps.clear();
    ctx->recs.clear();
    static const bool one_branch = getenv("CIDNET_ONE_BRANCH") != nullptr;
    if ((capture || ctx->profiling) && !one_branch) {
        // two branches: the I and HV launches of a pair on two streams (graph capture; or eagerly while profiling, so
        // that the per-kernel times describe the schedule the replayed graph runs)
        if (!ctx->cap_stream2) CIDNET_CUDA_OK(cudaStreamCreateWithFlags(&ctx->cap_stream2, cudaStreamNonBlocking));
        if (!ctx->ev_fork) CIDNET_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        if (!ctx->ev_join) CIDNET_CUDA_OK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        f.par = true; f.st2 = ctx->cap_stream2;
    }
    if (capture) {
        if (!ctx->cap_stream) CIDNET_CUDA_OK(cudaStreamCreateWithFlags(&ctx->cap_stream, cudaStreamNonBlocking));
        f.st = ctx->cap_stream;
        CIDNET_CUDA_OK(cudaStreamBeginCapture(f.st, cudaStreamCaptureModeRelaxed));
    }
    const std::map<const void*, int> margin0 = f.sh.margin;
    int rc = f.run(rgb_in, rgb_out, k_dev, gated, alpha_s, gated2, alpha);
    ctx->launches = f.launches;
    if (capture) {
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(f.st, &graph);
        f.st = st;
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        CIDNET_CHECK(e == cudaSuccess && graph, CIDNET_ERR_CUDA, std::string("forward: graph capture failed: ") + cudaGetErrorString(e));
        // locate the stem / head kernel nodes
        size_t nn = 0;
        CIDNET_CUDA_OK(cudaGraphGetNodes(graph, nullptr, &nn));
        std::vector<cudaGraphNode_t> nodes(nn);
        CIDNET_CUDA_OK(cudaGraphGetNodes(graph, nodes.data(), &nn));
        for (cudaGraphNode_t nd : nodes) {
            cudaGraphNodeType ty;
            if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp{};
            if (cudaGraphKernelNodeGetParams(nd, &kp) != cudaSuccess) continue;
            if (kp.func == stem_kernel_func()) {
                ge->stem_node = nd; ge->stem_p = kp;
                for (int i = 0; i < kStemNumArgs; ++i) ge->stem_args[i] = kp.kernelParams[i];
            } else if (kp.func == head_kernel_func()) {
                ge->head_node = nd; ge->head_p = kp;
                for (int i = 0; i < kHeadNumArgs; ++i) ge->head_args[i] = kp.kernelParams[i];
            }
        }
        cudaGraphExec_t exec = nullptr;
        e = (ge->stem_node && ge->head_node) ? cudaGraphInstantiate(&exec, graph, 0) : cudaErrorUnknown;
        if (e != cudaSuccess) {               // fall back to eager launches for this key
            cudaGetLastError();
            cudaGraphDestroy(graph);
            ge->seen = -1000000;
            f.par = false;
            f.sh.margin = margin0; f.sh.halo_calls = f.sh.allreduce_calls = 0; f.launches = 0;
            return f.run(rgb_in, rgb_out, k_dev, gated, alpha_s, gated2, alpha);
        }
        ge->graph = graph; ge->exec = exec;
        ge->cur_in = rgb_in; ge->cur_out = rgb_out;
        ge->taps = ctx->taps; ge->launches = f.launches;
        CIDNET_CUDA_OK(cudaGraphLaunch(exec, st));
        return CIDNET_OK;
    }
    return rc;
}

extern "C" int cidnet_forward(cidnet_ctx* ctx, const float* rgb_in, float* rgb_out, int B, int H, int W,
                              void* workspace, int64_t workspace_bytes, const float* k_dev,
                              int gated, float alpha_s, int gated2, float alpha, void* stream) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "forward: null ctx");
    CIDNET_CHECK(ctx->finalized, CIDNET_ERR_STATE, "forward: weights not finalized (call cidnet_finalize_weights)");
    CIDNET_CHECK(B >= 0 && H > 0 && W > 0, CIDNET_ERR_INVALID, "forward: bad shape");
    CIDNET_CHECK(H % 8 == 0 && W % 8 == 0, CIDNET_ERR_INVALID,
                 "forward: H and W must be multiples of 8 (got " + std::to_string(H) + "x" + std::to_string(W) +
                     "); the reference fails in torch.cat for such inputs, callers pad first");
    if (B == 0) return CIDNET_OK;
    DeviceGuard guard(ctx->device);
    CIDNET_CHECK(rgb_in && rgb_out && workspace, CIDNET_ERR_INVALID, "forward: null pointer");
    CIDNET_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, CIDNET_ERR_INVALID, "forward: workspace must be 1024-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Fwd f;
    f.ctx = ctx; f.st = st;
    make_plan(&f.P, workspace, B, H, W);
    CIDNET_CHECK(workspace_bytes >= f.P.bytes, CIDNET_ERR_STATE,
                 "forward: workspace too small: need " + std::to_string(f.P.bytes) + " bytes");
    ctx->last_B = B;
    return run_forward_cached(ctx, f, 0, rgb_in, rgb_out, k_dev, gated, alpha_s, gated2, alpha, st);
}

// ---- 8-bit in, 8-bit out: the callers' pre / post-processing fused into the stem's load and the head's store ------------
extern "C" int cidnet_forward_u8(cidnet_ctx* ctx, const uint8_t* in_hwc, uint8_t* out_hwc, int B, int h, int w, float gamma,
                                 void* workspace, int64_t workspace_bytes, const float* k_dev, int gated, float alpha_s,
                                 int gated2, float alpha, void* stream) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "forward_u8: null ctx");
    CIDNET_CHECK(ctx->finalized, CIDNET_ERR_STATE, "forward_u8: weights not finalized (call cidnet_finalize_weights)");
    CIDNET_CHECK(B >= 0 && h > 0 && w > 0, CIDNET_ERR_INVALID, "forward_u8: bad shape");
    if (B == 0) return CIDNET_OK;
    // the callers pad to the next multiple of 8 only when needed (data/eval_sets.py:22-27)
    const int H = (h % 8) ? (h / 8 + 1) * 8 : h, W = (w % 8) ? (w / 8 + 1) * 8 : w;
    CIDNET_CHECK(H - h < h && W - w < w, CIDNET_ERR_INVALID, "forward_u8: reflect padding must be smaller than the image");
    DeviceGuard guard(ctx->device);
    CIDNET_CHECK(in_hwc && out_hwc && workspace, CIDNET_ERR_INVALID, "forward_u8: null pointer");
    CIDNET_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, CIDNET_ERR_INVALID, "forward_u8: workspace must be 1024-byte aligned");
    Fwd f;
    f.ctx = ctx; f.st = (cudaStream_t)stream;
    make_plan(&f.P, workspace, B, H, W);
    CIDNET_CHECK(workspace_bytes >= f.P.bytes, CIDNET_ERR_STATE,
                 "forward_u8: workspace too small: need " + std::to_string(f.P.bytes) + " bytes (cidnet_workspace_bytes of the padded shape)");
    f.u8.on = true; f.u8.h = h; f.u8.w = w; f.u8.gamma = gamma;
    uint32_t gbits; memcpy(&gbits, &gamma, 4);
    const uint64_t extra = (((uint64_t)h << 40) ^ ((uint64_t)w << 20) ^ gbits) * 0x9E3779B97F4A7C15ull | 2ull;
    ctx->last_B = B;
    return run_forward_cached(ctx, f, extra, in_hwc, out_hwc, k_dev, gated, alpha_s, gated2, alpha, (cudaStream_t)stream);
}

// ---- row-strip sharded forward (single image over the GPUs of one node) -------------------
static int check_shard(const cidnet_shard* sh, int W) {
    CIDNET_CHECK(sh != nullptr, CIDNET_ERR_INVALID, "forward_sharded: null shard descriptor");
    CIDNET_CHECK(sh->nranks >= 1 && sh->rank >= 0 && sh->rank < sh->nranks, CIDNET_ERR_INVALID, "forward_sharded: bad rank");
    CIDNET_CHECK(W > 0 && W % 8 == 0 && sh->H_global > 0 && sh->H_global % 8 == 0, CIDNET_ERR_INVALID,
                 "forward_sharded: H and W must be multiples of 8");
    CIDNET_CHECK(sh->row_begin % 8 == 0 && sh->row_end % 8 == 0 && sh->row_begin >= 0 && sh->row_begin < sh->row_end &&
                     sh->row_end <= sh->H_global, CIDNET_ERR_INVALID, "forward_sharded: bad owned row range");
    CIDNET_CHECK((sh->rank == 0) == (sh->row_begin == 0) && (sh->rank == sh->nranks - 1) == (sh->row_end == sh->H_global),
                 CIDNET_ERR_INVALID, "forward_sharded: the first / last rank must own the first / last rows");
    if (sh->nranks > 1)
        CIDNET_CHECK(sh->halo >= 16 && sh->halo % 16 == 0 && sh->halo <= sh->row_end - sh->row_begin, CIDNET_ERR_INVALID,
                     "forward_sharded: halo must be a multiple of 16, >= 16 and <= the owned rows");
    return CIDNET_OK;
}
static void shard_state(const cidnet_shard* sh, ShardState* s) {
    s->on = true; s->rank = sh->rank; s->nranks = sh->nranks; s->gH = sh->H_global;
    s->halo_top = (sh->nranks > 1 && sh->rank > 0) ? sh->halo : 0;
    s->halo_bot = (sh->nranks > 1 && sh->rank < sh->nranks - 1) ? sh->halo : 0;
    s->row0 = sh->row_begin - s->halo_top;
}

extern "C" int cidnet_shard_plan(int H, int nranks, int rank, int halo, cidnet_shard* out) {
    CIDNET_CHECK(out != nullptr && H > 0 && H % 8 == 0 && nranks >= 1 && rank >= 0 && rank < nranks, CIDNET_ERR_INVALID,
                 "shard_plan: bad arguments");
    const int rows3 = H / 8, base = rows3 / nranks, extra = rows3 % nranks;
    CIDNET_CHECK(base >= 1, CIDNET_ERR_INVALID, "shard_plan: more ranks than coarsest-level rows");
    const int b3 = rank * base + (rank < extra ? rank : extra);
    const int n3 = base + (rank < extra ? 1 : 0);
    out->rank = rank; out->nranks = nranks; out->H_global = H;
    out->row_begin = 8 * b3; out->row_end = 8 * (b3 + n3); out->halo = nranks > 1 ? halo : 0;
    return check_shard(out, 8);
}

extern "C" int cidnet_shard_local_rows(const cidnet_shard* sh) {
    if (!sh) return 0;
    ShardState s; shard_state(sh, &s);
    return sh->row_end - sh->row_begin + s.halo_top + s.halo_bot;
}

extern "C" int cidnet_forward_sharded(cidnet_ctx* ctx, const float* rgb_local, float* rgb_out_local, int W,
                                      const cidnet_shard* sh, void* workspace, int64_t workspace_bytes,
                                      const float* k_dev, int gated, float alpha_s, int gated2, float alpha,
                                      cidnet_halo_fn halo_fn, cidnet_allreduce_fn allreduce_fn, void* user,
                                      void* stream) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "forward_sharded: null ctx");
    CIDNET_CHECK(ctx->finalized, CIDNET_ERR_STATE, "forward_sharded: weights not finalized (call cidnet_finalize_weights)");
    int rc = check_shard(sh, W);
    if (rc) return rc;
    CIDNET_CHECK(sh->nranks == 1 || (halo_fn && allreduce_fn), CIDNET_ERR_INVALID, "forward_sharded: callbacks required");
    CIDNET_CHECK(rgb_local && rgb_out_local && workspace, CIDNET_ERR_INVALID, "forward_sharded: null pointer");
    CIDNET_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, CIDNET_ERR_INVALID, "forward_sharded: workspace must be 1024-byte aligned");
    DeviceGuard guard(ctx->device);
    Fwd f;
    f.ctx = ctx; f.st = (cudaStream_t)stream;
    shard_state(sh, &f.sh);
    f.sh.halo_fn = halo_fn; f.sh.ar_fn = allreduce_fn; f.sh.user = user;
    const int H = cidnet_shard_local_rows(sh);
    make_plan(&f.P, workspace, 1, H, W);
    CIDNET_CHECK(workspace_bytes >= f.P.bytes, CIDNET_ERR_STATE,
                 "forward_sharded: workspace too small: need " + std::to_string(f.P.bytes) + " bytes");
    ctx->last_B = 1;
    ctx->taps.clear();
    ctx->recs.clear();
    rc = f.run(rgb_local, rgb_out_local, k_dev, gated, alpha_s, gated2, alpha);
    ctx->launches = f.launches;
    return rc;
}

// ---- the same with the peer-memory transport: no host callbacks, the exchanges are kernels of this library --------
extern "C" int64_t cidnet_peer_workspace_bytes(int H_global, int W, int nranks, int halo) {
    // one size for every rank (the largest strip) + the synchronisation header
    int64_t mx = 0;
    for (int r = 0; r < nranks; ++r) {
        cidnet_shard s;
        if (cidnet_shard_plan(H_global, nranks, r, halo, &s)) return 0;
        Plan P;
        make_plan(&P, nullptr, 1, cidnet_shard_local_rows(&s), W);
        mx = std::max(mx, P.bytes);
    }
    return mx + kPeerHdrBytes;
}

extern "C" int cidnet_forward_sharded_peer(cidnet_ctx* ctx, const float* rgb_local, float* rgb_out_local, int W,
                                           const cidnet_shard* sh, void* const* ws_all, int64_t ws_bytes,
                                           const float* k_dev, int gated, float alpha_s, int gated2, float alpha,
                                           void* stream) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "forward_sharded_peer: null ctx");
    CIDNET_CHECK(ctx->finalized, CIDNET_ERR_STATE, "forward_sharded_peer: weights not finalized (call cidnet_finalize_weights)");
    int rc = check_shard(sh, W);
    if (rc) return rc;
    CIDNET_CHECK(sh->nranks <= kPeerMaxRanks && ws_all && rgb_local && rgb_out_local, CIDNET_ERR_INVALID,
                 "forward_sharded_peer: bad arguments (at most 8 ranks)");
    CIDNET_CHECK(ws_bytes >= cidnet_peer_workspace_bytes(sh->H_global, W, sh->nranks, sh->halo), CIDNET_ERR_STATE,
                 "forward_sharded_peer: workspace too small (cidnet_peer_workspace_bytes)");
    DeviceGuard guard(ctx->device);
    Fwd f;
    f.ctx = ctx; f.st = (cudaStream_t)stream;
    shard_state(sh, &f.sh);
    f.sh.peer = sh->nranks > 1;
    uint64_t sig = 1469598103934665603ull;                     // FNV-1a over the shard geometry and the mappings
    auto mix = [&](uint64_t v) { sig = (sig ^ v) * 1099511628211ull; };
    mix((uint64_t)sh->rank); mix((uint64_t)sh->nranks); mix((uint64_t)sh->H_global); mix((uint64_t)sh->row_begin);
    mix((uint64_t)sh->row_end); mix((uint64_t)sh->halo);
    f.sh.plans.resize(sh->nranks); f.sh.shards.resize(sh->nranks);
    for (int r = 0; r < sh->nranks; ++r) {
        CIDNET_CHECK(ws_all[r] && (reinterpret_cast<uintptr_t>(ws_all[r]) & 1023) == 0, CIDNET_ERR_INVALID,
                     "forward_sharded_peer: every rank's workspace mapping must be non-null and 1024-byte aligned");
        f.sh.ws_all[r] = reinterpret_cast<uint8_t*>(ws_all[r]);
        if ((rc = cidnet_shard_plan(sh->H_global, sh->nranks, r, sh->halo, &f.sh.shards[r]))) return rc;
        make_plan(&f.sh.plans[r], f.sh.ws_all[r] + kPeerHdrBytes, 1, cidnet_shard_local_rows(&f.sh.shards[r]), W);
        mix(reinterpret_cast<uint64_t>(ws_all[r]));
    }
    f.P = f.sh.plans[sh->rank];
    ctx->last_B = 1;
    rc = run_forward_cached(ctx, f, sig | 1ull, rgb_local, rgb_out_local, k_dev, gated, alpha_s, gated2, alpha, (cudaStream_t)stream);
    return rc;
}

extern "C" int cidnet_forward_sharded_dry(int W, const cidnet_shard* sh, void* workspace, int64_t workspace_bytes,
                                          cidnet_halo_fn halo_fn, cidnet_allreduce_fn allreduce_fn, void* user,
                                          int* n_halo_calls, int* n_allreduce_calls) {
    return cidnet_forward_sharded_dry_variant(CIDNET_VARIANT_BASE, W, sh, workspace, workspace_bytes, halo_fn, allreduce_fn, user,
                                              n_halo_calls, n_allreduce_calls);
}

extern "C" int cidnet_forward_sharded_dry_variant(int variant, int W, const cidnet_shard* sh, void* workspace,
                                                  int64_t workspace_bytes, cidnet_halo_fn halo_fn,
                                                  cidnet_allreduce_fn allreduce_fn, void* user, int* n_halo_calls,
                                                  int* n_allreduce_calls) {
    CIDNET_CHECK(variant == CIDNET_VARIANT_BASE || variant == CIDNET_VARIANT_MSSA, CIDNET_ERR_INVALID, "forward_sharded_dry: unknown variant");
    int rc = check_shard(sh, W);
    if (rc) return rc;
    CIDNET_CHECK(sh->nranks == 1 || (halo_fn && allreduce_fn), CIDNET_ERR_INVALID, "forward_sharded_dry: callbacks required");
    CIDNET_CHECK(workspace && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, CIDNET_ERR_INVALID,
                 "forward_sharded_dry: workspace must be 1024-byte aligned");
    cidnet_ctx dummy;                      // weights are only dereferenced by the (skipped) launches
    dummy.variant = variant;
    for (int n = 0; n < 6; ++n) {
        const int l = n < 3 ? n + 1 : 6 - n;
        for (int s = 0; s < 2; ++s) {
            LcaWeights& L = dummy.stage[n].lca[s];
            L.live = !(n == 4 && s == 0) || variant == CIDNET_VARIANT_MSSA;
            L.C = kCh[l]; L.Cp = act_pitch(kCh[l]); L.heads = kHeads[l]; L.h = (int)(kCh[l] * 2.66); L.hp = round_up(L.h, 16);
            choose_blocking(L.C, &L.fold_tmpl.block_n, &L.fold_tmpl.n_blocks);
            L.fold_tmpl.n_rows = L.fold_tmpl.block_n * L.fold_tmpl.n_blocks;
        }
    }
    Fwd f;
    f.ctx = &dummy; f.st = nullptr;
    shard_state(sh, &f.sh);
    f.sh.dry = true;
    f.sh.halo_fn = halo_fn; f.sh.ar_fn = allreduce_fn; f.sh.user = user;
    make_plan(&f.P, workspace, 1, cidnet_shard_local_rows(sh), W);
    CIDNET_CHECK(workspace_bytes >= f.P.bytes, CIDNET_ERR_STATE, "forward_sharded_dry: workspace too small");
    rc = f.run(nullptr, nullptr, nullptr, 0, 1.f, 0, 1.f);
    if (n_halo_calls) *n_halo_calls = f.sh.halo_calls;
    if (n_allreduce_calls) *n_allreduce_calls = f.sh.allreduce_calls;
    return rc;
}

extern "C" int cidnet_forward_launches(cidnet_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int cidnet_read_tap(cidnet_ctx* ctx, const char* name, float* dst, int64_t dst_numel, int* dims,
                               void* stream) {
    CIDNET_CHECK(ctx && name && dims, CIDNET_ERR_INVALID, "read_tap: null argument");
    auto it = ctx->taps.find(name);
    CIDNET_CHECK(it != ctx->taps.end(), CIDNET_ERR_INVALID, std::string("read_tap: unknown tap '") + name + "'");
    const Tap& t = it->second;
    dims[0] = t.C; dims[1] = t.H; dims[2] = t.W;
    if (dst == nullptr) return CIDNET_OK;      // query only
    const int64_t n = (int64_t)ctx->last_B * t.C * t.H * t.W;
    CIDNET_CHECK(dst_numel >= n, CIDNET_ERR_INVALID, "read_tap: destination too small");
    if (t.f32_nchw) {
        CIDNET_CUDA_OK(cudaMemcpyAsync(dst, t.ptr, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return CIDNET_OK;
    }
    return launch_nhwc_to_nchw(reinterpret_cast<const act_t*>(t.ptr), dst, ctx->last_B, t.C, t.H, t.W, t.pitch,
                               (cudaStream_t)stream);
}

// ---- per-launch profiling (bench.py's roofline; events on the launching stream) ------------
extern "C" int cidnet_profile_enable(cidnet_ctx* ctx, int enable) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "profile_enable: null ctx");
    ctx->profiling = enable != 0;
    return CIDNET_OK;
}
extern "C" int cidnet_profile_count(cidnet_ctx* ctx) { return ctx ? (int)ctx->recs.size() : 0; }
extern "C" int cidnet_profile_get(cidnet_ctx* ctx, int i, char* name, int name_cap, float* ms, double* alg_bytes,
                                  double* flops) {
    CIDNET_CHECK(ctx && i >= 0 && i < (int)ctx->recs.size(), CIDNET_ERR_INVALID, "profile_get: bad index");
    const auto& r = ctx->recs[i];
    if (name && name_cap > 0) { snprintf(name, name_cap, "%s", r.name.c_str()); }
    if (alg_bytes) *alg_bytes = r.bytes;
    if (flops) *flops = r.flops;
    if (ms) CIDNET_CUDA_OK(cudaEventElapsedTime(ms, ctx->events[2 * i], ctx->events[2 * i + 1]));
    return CIDNET_OK;
}

// start / end of launch i relative to the first launch of the profiled forward (ms): launches of a pair run concurrently on two
// streams, so the wall time of a pair is max(end) - min(start), not the sum of the two durations
extern "C" int cidnet_profile_get_span(cidnet_ctx* ctx, int i, float* start_ms, float* end_ms) {
    CIDNET_CHECK(ctx && i >= 0 && i < (int)ctx->recs.size() && start_ms && end_ms, CIDNET_ERR_INVALID, "profile_get_span: bad index");
    CIDNET_CUDA_OK(cudaEventElapsedTime(start_ms, ctx->events[0], ctx->events[2 * i]));
    CIDNET_CUDA_OK(cudaEventElapsedTime(end_ms, ctx->events[0], ctx->events[2 * i + 1]));
    return CIDNET_OK;
}

// CUDA-graph replay of cidnet_forward is on by default; 0 turns it off (every call launches eagerly)
extern "C" int cidnet_set_graphs(cidnet_ctx* ctx, int enable) {
    CIDNET_CHECK(ctx, CIDNET_ERR_INVALID, "set_graphs: null ctx");
    ctx->use_graphs = enable != 0;
    return CIDNET_OK;
}

// ---- unit-test hook: ONE LCA stage pair, isolated from the rest of the network ---------------------------------
// Runs I_LCA<n>(x_i, x_hv) and HV_LCA<n>(x_hv, x_i) (net/LCA.py:71-93) of a finalized context on the given fp32 NCHW
// tensors [B, C, H, W] (H, W = the resolution of the stage's level; multiples of 1) and returns, as fp32 NCHW, the
// tensors after the attention (x + CAB(norm x, norm y)) and the block outputs.  stat_y0 / stat_y1 (0, 0 = all rows)
// restrict the rows that enter the Gram / sum q^2 / sum k^2 statistics, exactly as the row-strip sharded forward
// restricts them to a rank's owned rows.  Allocates scratch memory and synchronises (tests only).
extern "C" CIDNET_API int cidnet_test_lca_stage(cidnet_ctx* ctx, int n, const float* x_i, const float* x_hv,
                                                float* after_cab_i, float* after_cab_hv, float* out_i, float* out_hv,
                                                int B, int H, int W, int stat_y0, int stat_y1, void* stream_) {
    CIDNET_CHECK(ctx && ctx->finalized, CIDNET_ERR_STATE, "test_lca_stage: weights not finalized");
    CIDNET_CHECK(n >= 1 && n <= 6 && x_i && x_hv && B > 0 && H > 0 && W > 0, CIDNET_ERR_INVALID, "test_lca_stage: bad arguments");
    DeviceGuard guard(ctx->device);
    cudaStream_t st = (cudaStream_t)stream_;
    const int l = n <= 3 ? n : 7 - n;
    const int C = kCh[l], Cp = act_pitch(C);
    if (stat_y1 <= 0) { stat_y0 = 0; stat_y1 = H; }
    CIDNET_CHECK(stat_y0 >= 0 && stat_y0 < stat_y1 && stat_y1 <= H, CIDNET_ERR_INVALID, "test_lca_stage: bad row range");
    Fwd f;
    f.ctx = ctx; f.st = st;
    make_plan(&f.P, nullptr, B, H << l, W << l);
    void* ws = nullptr;
    CIDNET_CUDA_OK(cudaMalloc(&ws, (size_t)f.P.bytes));
    make_plan(&f.P, ws, B, H << l, W << l);
    // a "shard" of one rank whose owned rows are [stat_y0, stat_y1): no exchange, but the statistics see only those rows
    f.sh.on = true; f.sh.nranks = 1; f.sh.gH = H << l;
    f.sh.halo_top = stat_y0 << l; f.sh.halo_bot = (H - stat_y1) << l;
    act_t* xi = f.P.enc_i[l]; act_t* xh = f.P.enc_hv[l];
    act_t* oi = f.P.lca_i[n]; act_t* oh = f.P.lca_hv[n];
    ctx->last_B = B;
    int rc = launch_nchw_to_nhwc(x_i, xi, B, C, H, W, Cp, st);
    if (!rc) rc = launch_nchw_to_nhwc(x_hv, xh, B, C, H, W, Cp, st);
    const bool i_live = ctx->stage[n - 1].lca[0].live;
    if (!rc) rc = f.lca_stage(n, xi, xh, i_live ? oi : nullptr, oh);
    if (!rc && after_cab_i && i_live) rc = launch_nhwc_to_nchw(f.P.xp[l][0], after_cab_i, B, C, H, W, Cp, st);
    if (!rc && after_cab_hv) rc = launch_nhwc_to_nchw(f.P.xp[l][1], after_cab_hv, B, C, H, W, Cp, st);
    if (!rc && out_i && i_live) rc = launch_nhwc_to_nchw(oi, out_i, B, C, H, W, Cp, st);
    if (!rc && out_hv) rc = launch_nhwc_to_nchw(oh, out_hv, B, C, H, W, Cp, st);
    cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(ws);
    if (rc) return rc;
    CIDNET_CHECK(e == cudaSuccess, CIDNET_ERR_CUDA, std::string("test_lca_stage: ") + cudaGetErrorString(e));
    return CIDNET_OK;
}
