/*
 * vvcdsp_cuda.h - C ABI of the B200 (sm_100a) back end for the ffvvc pixel-reconstruction
 * hot path (the VVCDSPContext tables of libavcodec/vvc).
 *
 * Two levels, both plain C (no CUDA or torch types in any signature):
 *
 *  1. Drop-in table override  ff_vvc_dsp_init_cuda()
 *     Called from ff_vvc_dsp_init() right after the arch hook
 *     (reference: libavcodec/vvc/vvcdsp.c:254-256, precedent
 *     libavcodec/x86/vvc/vvcdsp_init.c:294-361).  It overwrites entries of the
 *     reference's own struct VVCDSPContext (libavcodec/vvc/vvcdsp.h:48-168) with
 *     functions of identical signature that take HOST pointers, stage through the
 *     context's pinned/device buffers and run the CUDA kernels.
 *
 *  2. Batched stage entries  vvc_cuda_<stage>_frame()
 *     New, additive API: one call = one pipeline stage over a whole picture (or a
 *     ring of pictures of equal geometry), fed by plain-old-data descriptors that
 *     mirror the argument lists of the table entries.  The legal stage order is the
 *     reference's task graph (libavcodec/vvc/vvc_thread.c:41-51,159-167):
 *     INTER -> RECON -> LMCS -> DEBLOCK_V -> DEBLOCK_H -> SAO -> ALF.
 *     Entries without a suffix take DEVICE pointers and are asynchronous on the
 *     context's stream; entries ending in _host take HOST pointers, copy in, run,
 *     copy out and return after the result is in host memory.
 *
 * Errors: table entries return void like the reference; CUDA failures are latched in
 * the context and reported by vvc_cuda_last_error()/vvc_cuda_sync() as negative ints.
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef VVCDSP_CUDA_H
#define VVCDSP_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * Context
 * ---------------------------------------------------------------------------------------- */
typedef struct VVCCudaCtx VVCCudaCtx;

#define VVC_CUDA_OK              0
#define VVC_CUDA_ERR_CUDA      (-1)   /* a CUDA runtime call or kernel failed (sticky)  */
#define VVC_CUDA_ERR_ARG       (-2)   /* unsupported geometry / bit depth / descriptor  */
#define VVC_CUDA_ERR_NOMEM     (-3)

/* device: CUDA ordinal.  stream: a cudaStream_t passed as void* (e.g. torch's current stream)
 * or NULL to let the context create its own non-blocking stream. */
int         vvc_cuda_ctx_create(VVCCudaCtx **ctx, int device, void *stream);
void        vvc_cuda_ctx_destroy(VVCCudaCtx *ctx);
int         vvc_cuda_sync(VVCCudaCtx *ctx);            /* waits for the stream; returns last error */
int         vvc_cuda_last_error(const VVCCudaCtx *ctx);
const char *vvc_cuda_error_string(const VVCCudaCtx *ctx);
void       *vvc_cuda_stream(const VVCCudaCtx *ctx);    /* the cudaStream_t in use */
/* number of kernel launches issued through this context since creation (bench bookkeeping) */
uint64_t    vvc_cuda_launch_count(const VVCCudaCtx *ctx);
const char *vvc_cuda_version(void);
/* sizeof() of descriptor `which` as compiled into the library (0 VVCCudaFrame, 1 VVCCudaALFCtb,
 * 2 VVCCudaALFSets, ...): lets foreign-language bindings verify their struct mirrors. */
size_t      vvc_cuda_abi_sizeof(int which);

/* ------------------------------------------------------------------------------------------
 * Pictures
 * ---------------------------------------------------------------------------------------- */
/* A picture, or a ring of `batch` pictures of identical geometry: picture k, plane c starts at
 * (uint8_t*)data[c] + k * batch_stride[c].  Samples are uint16_t (bit_depth 10 or 12), as in
 * the reference's high-bit-depth template (pixel = uint16_t, libavcodec/bit_depth_template.c).
 * stride[] is the row pitch in BYTES like every stride of the reference API. */
typedef struct VVCCudaFrame {
    void     *data[3];
    ptrdiff_t stride[3];
    ptrdiff_t batch_stride[3];
    int32_t   width, height;      /* luma samples */
    int32_t   hshift, vshift;     /* chroma subsampling shifts (4:2:0 -> 1,1) */
    int32_t   bit_depth;
    int32_t   ctb_log2;           /* log2 of the CTU size (5..7) */
    int32_t   batch;              /* >= 1 */
    int32_t   chroma_format_idc;  /* 0 = luma only */
} VVCCudaFrame;

/* CTB-edge flags shared by SAO/ALF descriptors: set when filtering must not look across that
 * side of the CTB (picture border, or slice/tile border with cross filtering disabled -
 * reference: edges[] in ff_vvc_alf_filter, libavcodec/vvc/vvc_filter.c:1266-1280). */
#define VVC_CUDA_EDGE_LEFT    1
#define VVC_CUDA_EDGE_TOP     2
#define VVC_CUDA_EDGE_RIGHT   4
#define VVC_CUDA_EDGE_BOTTOM  8

/* ------------------------------------------------------------------------------------------
 * ALF stage (replaces ff_vvc_alf_filter, libavcodec/vvc/vvc_filter.c:1254-1319, and the table
 * entries alf.classify / recon_coeff_and_clip / filter[2] / filter_cc,
 * libavcodec/vvc/vvc_filter_template.c:43-408)
 * ---------------------------------------------------------------------------------------- */
/* per-CTB parameters; mirrors ALFParams (libavcodec/vvc/vvc_ctu.h:453-460) */
typedef struct VVCCudaALFCtb {
    uint8_t ctb_flag[3];           /* alf_ctb_flag[c]                                  */
    uint8_t filt_set_idx_y;        /* AlfCtbFiltSetIdxY: 0..15 fixed sets, 16+k = luma APS slot k */
    uint8_t chroma_alt_idx[2];     /* alf_ctb_filter_alt_idx[]                         */
    uint8_t cc_idc[2];             /* alf_ctb_cc_cb_idc / alf_ctb_cc_cr_idc (0 = off)  */
    uint8_t edges;                 /* VVC_CUDA_EDGE_* ; picture borders are added by the callee */
    uint8_t reserved[3];
} VVCCudaALFCtb;

#define VVC_CUDA_ALF_MAX_LUMA_APS 8
/* per-picture filter sets, already resolved from the APS ids of the slice header; mirrors the
 * fields of VVCALF (libavcodec/vvc/vvc_ps.h:163-173) */
typedef struct VVCCudaALFSets {
    int16_t luma_coeff[VVC_CUDA_ALF_MAX_LUMA_APS][25][12];
    uint8_t luma_clip_idx[VVC_CUDA_ALF_MAX_LUMA_APS][25][12];
    int16_t chroma_coeff[8][6];
    uint8_t chroma_clip_idx[8][6];
    int16_t cc_coeff[2][5][7];     /* [cb,cr][alf_ctb_cc_idc - 1][tap] */
} VVCCudaALFSets;

/* dst must not alias src (the stage ping-pongs two pictures; the reference's in-place CTU walk
 * with saved border lines is bit-identical).  ctbs: ctb_count entries per picture
 * of the batch, raster order.  sets: one per picture when sets_per_frame != 0, else shared. */
int vvc_cuda_alf_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                       const VVCCudaALFCtb *ctbs, const VVCCudaALFSets *sets, int sets_per_frame);
int vvc_cuda_alf_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                            const VVCCudaALFCtb *ctbs, const VVCCudaALFSets *sets, int sets_per_frame);

#ifdef __cplusplus
}
#endif
#endif /* VVCDSP_CUDA_H */
