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

/* ---- model context --------------------------------------------------------
 * replaces CIDNet.__init__ / load_state_dict (net/CIDNet.py:9-69; 191 tensors,
 * SURVEY App. B).  set_weight copies one fp32 state_dict tensor (host memory);
 * finalize_weights folds / packs them into the device layouts the kernels use. */
CIDNET_API int cidnet_create(cidnet_ctx** ctx, int device);
CIDNET_API int cidnet_destroy(cidnet_ctx* ctx);
CIDNET_API int cidnet_set_weight(cidnet_ctx* ctx, const char* key, const float* host, int64_t numel);
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

#ifdef __cplusplus
}
#endif
#endif /* CIDNET_B200_H_ */
