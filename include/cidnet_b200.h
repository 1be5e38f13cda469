/* libcidnet_b200.so -- C ABI of the B200-native CIDNet inference forward path.
 *
 * The reference (KitaharaH/HVI-CIDNet) has no FFI: its boundary is the Python
 * class net.CIDNet.CIDNet (net/CIDNet.py:8).  The host-side mirror of that class
 * (hvi-cidnet_b200/net/CIDNet.py) binds exactly these entry points through
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer marked "dev" is a CUDA device pointer owned by the caller
 *     (a torch tensor kept alive by the Python wrapper); "host" pointers are
 *     ordinary host memory;
 *   - `stream` is a cudaStream_t passed as void*; all compute entry points are
 *     stream-ordered and never synchronise or allocate;
 *   - return value: 0 = ok, <0 = error (see CIDNET_ERR_*); the message is
 *     available from cidnet_last_error() (thread-local);
 *   - images are fp32 NCHW planar, exactly as the reference takes them;
 *   - a context is bound to one device and is not thread-safe (the reference
 *     object is not re-entrant either: app.py:30,42-43).
 */
#ifndef CIDNET_B200_H_
#define CIDNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CIDNET_API __attribute__((visibility("default")))
#else
#define CIDNET_API
#endif

#define CIDNET_OK            0
#define CIDNET_ERR_INVALID  (-1)  /* bad shape / argument (reference: RuntimeError from torch.cat when H or W % 8 != 0) */
#define CIDNET_ERR_ARCH     (-2)  /* device is not sm_100 */
#define CIDNET_ERR_CUDA     (-3)  /* CUDA runtime / driver error */
#define CIDNET_ERR_STATE    (-4)  /* weights missing / not finalized / workspace too small */

typedef struct cidnet_ctx cidnet_ctx;

CIDNET_API const char* cidnet_last_error(void);
CIDNET_API int cidnet_abi_version(void);
/* 0 = fp16 operands/activations, 1 = bf16 (build-time choice) */
CIDNET_API int cidnet_act_dtype(void);

/* ---- RGB <-> HVI transform ------------------------------------------------
 * replaces RGB_HVI.HVIT   (net/HVI_transform.py:16-47)  and
 *          RGB_HVI.PHVIT  (net/HVI_transform.py:49-122).
 * rgb/hvi: dev fp32 [B,3,H,W].  k = density_k (HVIT) / this_k (PHVIT).
 * k_dev (optional, dev fp32[1]): when non-NULL the kernel reads k from device
 * memory and the scalar `k` is ignored -- lets the caller pass the live
 * `density_k` parameter without the host sync of the reference's k.item() (:38). */
CIDNET_API int cidnet_hvit(const float* rgb, float* hvi, int B, int H, int W, float k, const float* k_dev,
                           void* stream);
CIDNET_API int cidnet_phvit(const float* hvi, float* rgb, int B, int H, int W, float k, const float* k_dev,
                            int gated, float alpha_s, int gated2, float alpha, void* stream);

/* ---- backward of the transform (vector-Jacobian products) -------------------
 * replaces what autograd derives from RGB_HVI.HVIT / PHVIT (net/HVI_transform.py:16-47, :49-122) in the reference's
 * training loop: `model.HVIT(output_rgb)` inside the loss (train.py:61-62) and `self.trans.PHVIT(output_hvi)` at the end
 * of the forward (net/CIDNet.py:121).  One launch each; nothing is saved by the forward: the per-pixel quantities are
 * recomputed from the forward's INPUT (rgb for HVIT, hvi for PHVIT).  All images dev fp32 [B,3,H,W].
 *   cidnet_hvit_backward : grad_rgb = J^T grad_hvi; grad_k (optional, dev fp32[1]) = d/d density_k summed over every pixel,
 *                          deterministic (per-CTA slots in `scratch`, dev memory of cidnet_hvi_backward_scratch_bytes(),
 *                          added in index order by a finishing launch); scratch may be NULL when grad_k is NULL.
 *   cidnet_phvit_backward: grad_hvi = J^T grad_rgb; k = this_k (a python float in the reference: no gradient). */
CIDNET_API int64_t cidnet_hvi_backward_scratch_bytes(void);
CIDNET_API int cidnet_hvit_backward(const float* rgb, const float* grad_hvi, float* grad_rgb, float* grad_k, void* scratch,
                                    int B, int H, int W, float k, const float* k_dev, void* stream);
CIDNET_API int cidnet_phvit_backward(const float* hvi, const float* grad_rgb, float* grad_hvi, int B, int H, int W, float k,
                                     const float* k_dev, int gated, float alpha_s, int gated2, float alpha, void* stream);

/* ---- model context --------------------------------------------------------
 * replaces CIDNet.__init__ / load_state_dict (net/CIDNet.py:9-69; 191 tensors,
 * SURVEY App. B).  set_weight copies one fp32 state_dict tensor (host memory);
 * finalize_weights folds / packs them into the device layouts the kernels use. */
CIDNET_API int cidnet_create(cidnet_ctx** ctx, int device);
CIDNET_API int cidnet_destroy(cidnet_ctx* ctx);
CIDNET_API int cidnet_set_weight(cidnet_ctx* ctx, const char* key, const float* host, int64_t numel);
/* Model variant (before finalize_weights).  CIDNET_VARIANT_MSSA replaces the fork's
 * net.CIDNet_MSSA.CIDNet (net/CIDNet_MSSA.py:28-161): six SpatialAttention gates (:10-25; keys
 * sa_{hv,i}{1,2,3}.conv1.weight [1,2,7,7]) after the up blocks, I_LCA5 live, ID_block2 fed by I_LCA5. */
#define CIDNET_VARIANT_BASE 0
#define CIDNET_VARIANT_MSSA 1
CIDNET_API int cidnet_set_variant(cidnet_ctx* ctx, int variant);
CIDNET_API int cidnet_finalize_weights(cidnet_ctx* ctx);

/* ---- forward --------------------------------------------------------------
 * replaces CIDNet.forward (net/CIDNet.py:71-122).  rgb_in/rgb_out: dev fp32
 * [B,3,H,W]; H and W must be multiples of 8.  `workspace` is dev memory of at
 * least cidnet_workspace_bytes(B,H,W) bytes, 1024-byte aligned.  k_dev (optional,
 * dev fp32[1]) overrides the density_k given to set_weight with the live parameter. */
CIDNET_API int64_t cidnet_workspace_bytes(int B, int H, int W);
CIDNET_API int cidnet_forward(cidnet_ctx* ctx, const float* rgb_in, float* rgb_out, int B, int H, int W,
                   void* workspace, int64_t workspace_bytes, const float* k_dev,
                   int gated, float alpha_s, int gated2, float alpha, void* stream);
/* cidnet_forward replays a CUDA graph of its ~90 kernel launches from the second call with the same
 * (shape, workspace, flags) on (the image pointers may change freely: they are patched into the
 * graph).  cidnet_set_graphs(ctx, 0) turns this off. */
CIDNET_API int cidnet_set_graphs(cidnet_ctx* ctx, int enable);
/* number of kernels one cidnet_forward launches (for bench.py's gpu_launches) */
CIDNET_API int cidnet_forward_launches(cidnet_ctx* ctx);

/* ---- single image, rows sharded over the GPUs of one node ---------------------
 * BASELINE.json configs[4] (1x3x2160x3840 over 8 B200): the reference has no multi-GPU path
 * at all; this is the spatial partition of CIDNet.forward (net/CIDNet.py:71-122).
 * Rank r owns the image rows [row_begin, row_end) (multiples of 8) and works on a LOCAL image
 * = those rows plus `halo` rows of each existing neighbour (rank-1 above, rank+1 below), laid
 * out contiguously: [halo_top | owned | halo_bot] with halo_top = (rank > 0 ? halo : 0),
 * halo_bot = (rank < nranks-1 ? halo : 0).
 *
 * Two things cross the ranks, both through callbacks the HOST supplies (torch.distributed /
 * NCCL over NVLink in hvi-cidnet_b200/dist.py) so that the library itself stays free of any
 * communication dependency:
 *   halo      every 3x3 stage invalidates one more outermost halo row; before a stage that needs
 *             more valid halo rows than are left, the library asks for the halos of a list of
 *             local tensors to be refreshed.  For request q the rank must (stream-ordered)
 *               send  its first owned  q.halo_top rows to rank-1   (if halo_top > 0)
 *               send  its last  owned  q.halo_bot rows to rank+1   (if halo_bot > 0)
 *               receive rank-1's rows into local rows [0, halo_top)
 *               receive rank+1's rows into local rows [rows-halo_bot, rows)
 *             rows are contiguous (NHWC): row i starts at base + i*row_bytes.
 *   allreduce sum over all ranks, in place, of `count` fp32 values: the raw partial per-head Gram
 *             and the partial sums of q^2, k^2 of one LCA stage (CAB, net/LCA.py:30-33);
 *             normalisation, temperature and softmax then run identically on every rank.
 * Every rank issues the same sequence of callbacks (the schedule depends only on `halo`).
 * The callbacks must enqueue their work on `stream` (or make `stream` wait for it) and return 0.
 */
typedef struct cidnet_shard {
    int32_t rank, nranks;
    int32_t H_global;            /* rows of the whole image (multiple of 8) */
    int32_t row_begin, row_end;  /* owned rows [row_begin, row_end), multiples of 8 */
    int32_t halo;                /* halo rows per interior side: multiple of 16, >= 16, <= owned rows of every rank */
} cidnet_shard;
typedef struct cidnet_halo_req {
    void* base;                  /* local tensor (dev), row 0 of the local image at the tensor's level */
    int64_t row_bytes;           /* bytes of one image row of this tensor */
    int32_t rows;                /* local rows = halo_top + owned + halo_bot */
    int32_t halo_top, halo_bot;  /* halo rows of this tensor above / below the owned rows (0 at an image border) */
    int32_t reserved;
} cidnet_halo_req;
typedef int (*cidnet_halo_fn)(void* user, const cidnet_halo_req* reqs, int n);
typedef int (*cidnet_allreduce_fn)(void* user, float* buf, int64_t count);

/* balanced partition of the H/8 coarsest-level rows over nranks (first H/8 % nranks ranks get one more) */
CIDNET_API int cidnet_shard_plan(int H, int nranks, int rank, int halo, cidnet_shard* out);
/* rows of the local image of `sh` (what rgb_local / rgb_out_local and the workspace are sized for) */
CIDNET_API int cidnet_shard_local_rows(const cidnet_shard* sh);
/* rgb_local / rgb_out_local: dev fp32 [1,3,local_rows,W]; only the owned rows of rgb_out_local are
 * meaningful.  workspace >= cidnet_workspace_bytes(1, local_rows, W).  With nranks == 1 this is
 * cidnet_forward (the callbacks are never called and may be NULL). */
CIDNET_API int cidnet_forward_sharded(cidnet_ctx* ctx, const float* rgb_local, float* rgb_out_local, int W,
                                      const cidnet_shard* sh, void* workspace, int64_t workspace_bytes,
                                      const float* k_dev, int gated, float alpha_s, int gated2, float alpha,
                                      cidnet_halo_fn halo_fn, cidnet_allreduce_fn allreduce_fn, void* user,
                                      void* stream);
/* ---- the same forward with the PEER-MEMORY transport (no host callbacks, no NCCL on the data path) -------------
 * Every rank allocates its workspace with cidnet_peer_alloc (one cudaMalloc'ed block of cidnet_peer_workspace_bytes,
 * zero-initialised; the first 4 KB are a synchronisation header) and hands the 64-byte CUDA IPC handle to the other
 * ranks of the node, which map it with cidnet_peer_open.  ws_all[r] is THIS process's pointer to rank r's workspace
 * (ws_all[rank] = the own allocation).  The halo rows and the partial attention statistics are then read straight from
 * the neighbours' memory over NVLink by two kernels of this library (csrc/peer.cu: flag handshake, rank-ordered sums),
 * and the whole strip forward replays as ONE CUDA graph.  Every rank must call this the same number of times with the
 * same geometry (the exchange counters advance in lockstep).  cidnet_peer_error reports a timed-out handshake. */
CIDNET_API int64_t cidnet_peer_workspace_bytes(int H_global, int W, int nranks, int halo);
CIDNET_API int cidnet_peer_alloc(int device, int64_t bytes, void** dev_ptr, void* handle64);
CIDNET_API int cidnet_peer_open(int device, const void* handle64, void** dev_ptr);
CIDNET_API int cidnet_peer_close(void* dev_ptr);
CIDNET_API int cidnet_peer_free(void* dev_ptr);
CIDNET_API int cidnet_peer_error(const void* own_ws, int* error_out);
CIDNET_API int cidnet_forward_sharded_peer(cidnet_ctx* ctx, const float* rgb_local, float* rgb_out_local, int W,
                                           const cidnet_shard* sh, void* const* ws_all, int64_t ws_bytes,
                                           const float* k_dev, int gated, float alpha_s, int gated2, float alpha,
                                           void* stream);
/* host-only dry run of the same schedule (no device, no kernels): calls the callbacks exactly as
 * cidnet_forward_sharded would, with `workspace` any host buffer of the same size -- used by the
 * CPU (gloo) tests of the exchange logic.  Needs no weights. */
CIDNET_API int cidnet_forward_sharded_dry(int W, const cidnet_shard* sh, void* workspace, int64_t workspace_bytes,
                                          cidnet_halo_fn halo_fn, cidnet_allreduce_fn allreduce_fn, void* user,
                                          int* n_halo_calls, int* n_allreduce_calls);
/* the same for a model variant (CIDNET_VARIANT_MSSA: one more halo refresh per up-block pair for the 7x7 gates,
 * both problems of LCA stage 5) */
CIDNET_API int cidnet_forward_sharded_dry_variant(int variant, int W, const cidnet_shard* sh, void* workspace,
                                                  int64_t workspace_bytes, cidnet_halo_fn halo_fn,
                                                  cidnet_allreduce_fn allreduce_fn, void* user, int* n_halo_calls,
                                                  int* n_allreduce_calls);

/* ---- 8-bit image I/O (the callers' pre / post-processing, one kernel each) ------------------
 * cidnet_pre_u8 : src dev u8 [B,h,w,3] (HWC, what PIL / a decoder yields) -> dst dev fp32 [B,3,H,W]:
 *                 transforms.ToTensor (x/255, HWC->CHW), reflect padding on the bottom / right up to
 *                 (H,W) (data/eval_sets.py:22-27, demo.py:47-52: H,W = h,w rounded up to multiples of 8)
 *                 and input**gamma (eval.py:64, demo.py:57).
 * cidnet_post_u8: src dev fp32 [B,3,H,W] -> dst dev u8 [B,h,w,3]: clamp(0,1) (eval.py:69), crop to
 *                 [:h,:w] (eval.py:71), transforms.ToPILImage (mul(255).byte(), i.e. truncation). */
/* cidnet_forward_u8: the whole caller loop around the model in one call -- src dev u8 [B,h,w,3] -> dst dev u8 [B,h,w,3]
 * with ToTensor + reflect pad + **gamma done inside the stem kernel's tile load and clamp + crop + quantise inside the head
 * kernel's store (no padded fp32 image in memory; 6 instead of 24 bytes per pixel at the boundary).  The workspace is
 * cidnet_workspace_bytes(B, H, W) of the PADDED shape (h, w rounded up to multiples of 8 when they are not already). */
CIDNET_API int cidnet_forward_u8(cidnet_ctx* ctx, const uint8_t* in_hwc, uint8_t* out_hwc, int B, int h, int w, float gamma,
                                 void* workspace, int64_t workspace_bytes, const float* k_dev, int gated, float alpha_s,
                                 int gated2, float alpha, void* stream);
CIDNET_API int cidnet_pre_u8(const uint8_t* src_hwc, float* dst_nchw, int B, int h, int w, int H, int W, float gamma,
                             void* stream);
CIDNET_API int cidnet_post_u8(const float* src_nchw, uint8_t* dst_hwc, int B, int h, int w, int H, int W, void* stream);

/* ---- parity taps ----------------------------------------------------------
 * After a forward, copy a named internal activation (NHWC, 16-bit) out as fp32
 * NCHW so tests can compare every stage with the oracle.  Names are those of
 * oracle/cidnet_oracle.py's taps ("i_enc0", "hv_1", "I_LCA1", ...).
 * `dims` receives {C, H, W}. */
CIDNET_API int cidnet_read_tap(cidnet_ctx* ctx, const char* name, float* dst, int64_t dst_numel,
                    int* dims, void* stream);

/* ---- per-launch profiling ----------------------------------------------------
 * With profiling enabled, cidnet_forward records a CUDA event on `stream` before
 * every kernel launch (and one after the last).  After synchronising, record i
 * gives the kernel's name, its device time, and its ALGORITHMIC bytes / flops
 * (DESIGN.md "kernels and rooflines") -- what bench.py's `roofline` is built from. */
CIDNET_API int cidnet_profile_enable(cidnet_ctx* ctx, int enable);
CIDNET_API int cidnet_profile_count(cidnet_ctx* ctx);
CIDNET_API int cidnet_profile_get(cidnet_ctx* ctx, int i, char* name, int name_cap, float* ms,
                                  double* alg_bytes, double* flops);
/* start / end of launch i in ms since the first launch of the profiled forward.  While profiling the forward runs with the
 * same two-stream split as the replayed graph: the two launches of an (I, HV) pair overlap, their wall time is
 * max(end) - min(start). */
CIDNET_API int cidnet_profile_get_span(cidnet_ctx* ctx, int i, float* start_ms, float* end_ms);

/* ---- unit-test hooks (tests/ only: they allocate scratch memory and synchronise) ------------
 * cidnet_test_conv: one implicit-GEMM convolution with one of the four fused epilogues
 *   x dev fp32 [B,Cin,H,W]; w_host host fp32 [Cout,Cin,k,k]; aux dev fp32 (residual / low-res tensor) or NULL;
 *   ln_host host fp32 [ln_w(Cin) | ln_b(Cin)] for mode 1; mode 0 STORE, 1 LN, 2 DOWN, 3 UP; out dev fp32.
 * cidnet_test_lca_stage: ONE LCA stage pair of a finalized context, isolated from the rest of the network
 *   (I_LCA<n>(x_i, x_hv), HV_LCA<n>(x_hv, x_i), net/LCA.py:71-93) on fp32 NCHW tensors [B,C,H,W] at the stage's
 *   own resolution; returns x + CAB(..) ("after_cab") and the block outputs.  stat_y0/stat_y1 (0,0 = all rows)
 *   restrict the rows entering the Gram / sum q^2 / sum k^2 exactly as row-strip sharding does. */
CIDNET_API int cidnet_test_conv(const float* x, const float* w_host, const float* aux, const float* ln_host, float* out,
                                int B, int Cin, int H, int W, int Cout, int ksize, int mode, int flat, float prelu,
                                void* stream);
CIDNET_API int cidnet_test_lca_stage(cidnet_ctx* ctx, int n, const float* x_i, const float* x_hv, float* after_cab_i,
                                     float* after_cab_hv, float* out_i, float* out_hv, int B, int H, int W,
                                     int stat_y0, int stat_y1, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CIDNET_B200_H_ */
