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
/* Stream-ordered completion report.  fn(opaque, status) runs on a driver thread once everything submitted through the
 * context before this call has finished on the GPU (status = the context's error state then); work submitted afterwards
 * does not wait for it.  This is where a decoder wired to the batched entries calls ff_vvc_report_progress /
 * ff_vvc_report_frame_finished (report_frame_progress, libavcodec/vvc/vvc_thread.c:390-410; vvc_refs.c:532-565) for the
 * stage it has just submitted, instead of blocking a worker thread in vvc_cuda_sync.  fn must not call into CUDA or
 * into this library. */
typedef void (*vvc_cuda_notify_fn)(void *opaque, int status);
int         vvc_cuda_notify(VVCCudaCtx *ctx, vvc_cuda_notify_fn fn, void *opaque);
/* number of kernel launches issued through this context since creation (bench bookkeeping) */
uint64_t    vvc_cuda_launch_count(const VVCCudaCtx *ctx);
const char *vvc_cuda_version(void);
/* Options.  VVC_CUDA_OPT_GENERIC_KERNELS = 1 routes every stage through the generic kernels (any bit
 * depth / alignment) even where a specialised 10-bit kernel exists; both are CUDA, results identical. */
#define VVC_CUDA_OPT_GENERIC_KERNELS 1
/* VVC_CUDA_OPT_ALF_WIDE_MULTIPLY = 1 sends every ALF block through the 32-bit-multiply path that otherwise serves only
 * coefficient sets holding +128 (the packed 16x2 / IDP.2A path carries coefficients -128..127); results identical, tests
 * compare the two. */
#define VVC_CUDA_OPT_ALF_WIDE_MULTIPLY 2
/* VVC_CUDA_OPT_INTER_TMA: 1 = the DMVR / BDOF records' reference windows are fetched by the copy engine (TMA box
 * per window, one record ahead of the computation), 0 = staged by the warp's own loads.  Results identical. */
#define VVC_CUDA_OPT_INTER_TMA 3
/* VVC_CUDA_OPT_REF_PAD = n (a multiple of 16, 0 = off): the reference rings handed to the inter / reconstruction entries
 * are PRE-PADDED - every plane carries n replicated luma samples (n / 2 chroma samples) on all four sides, as
 * vvc_cuda_pad_frame() leaves them (data[] still points at sample (0, 0); stride[] spans the margins).  Those samples are
 * what ff_emulated_edge_mc / emulated_edge (libavcodec/vvc/vvc_inter.c:33-58) would fabricate per block, so windows that
 * stay inside the margin take the plain load path; vectors that point further out keep the clamped path.  Results are
 * identical with and without the option. */
#define VVC_CUDA_OPT_REF_PAD 4
int         vvc_cuda_ctx_set_option(VVCCudaCtx *ctx, int option, int value);
/* sizeof() of descriptor `which` as compiled into the library (0 VVCCudaFrame, 1 VVCCudaALFCtb,
 * 2 VVCCudaALFSets, 3 VVCCudaDbkEdge, 4 VVCCudaDeblockMaps, 5 VVCCudaSAOCtb, 6 VVCCudaInloopDesc, 7 VVCCudaTB, 8 VVCCudaPB, 9 VVCCudaWP, 10 VVCCudaProf, 11 VVCCudaDmvrOut, 12 VVCCudaRect, 13 VVCCudaReconDesc, 14 VVCCudaIntraPB, 15 VVCCudaCiip, 16 VVCCudaTBQuant, 17 VVCCudaScalingList, 18 VVCCudaCoeffs, 19 VVCCudaLmcsVpdu, 20 VVCCudaLmcsParams, 21 VVCCudaIntraBlk, 22 VVCCudaDbkTU, 23 VVCCudaDbkMvf, 24 VVCCudaDbkCtb, 25 VVCCudaDbkParams, 26 VVCCudaDbkSide): lets foreign-language bindings verify their struct mirrors. */
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

/* Replicates the border samples of every plane of every picture of the ring into a margin of `pad` luma samples
 * (pad >> hshift / vshift for chroma): sample (x, y) outside the plane becomes sample (clip(x), clip(y)).  The planes
 * must have been allocated with that margin around them.  A decoder calls it once per output picture that will be used
 * as a reference (the DPB format of VVC_CUDA_OPT_REF_PAD). */
int vvc_cuda_pad_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int pad);

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

/* ------------------------------------------------------------------------------------------
 * Deblocking stage (replaces the pixel work of ff_vvc_deblock_vertical / _horizontal,
 * libavcodec/vvc/vvc_filter.c:861-1003, i.e. the table entries lf.filter_luma[2] /
 * lf.filter_chroma[2], libavcodec/vvc/vvc_filter_template.c:466-754 and
 * libavcodec/h26x/h2656_deblock_template.c:25-99)
 * ---------------------------------------------------------------------------------------- */
/* One 4-sample (luma) / (4 >> subsampling)-sample (chroma) segment of one edge: exactly the
 * per-segment inputs the reference hands to lf.filter_luma / lf.filter_chroma
 * (_tc[i], _beta[i], _max_len_p[i], _max_len_q[i]; no_p/no_q are always 0 in the reference,
 * vvc_filter.c:870-871).  tc == 0 marks "not filtered" (bs == 0).  tc and beta are the values
 * of the reference's tctable/betatable lookups, BEFORE the bit-depth shift the DSP applies. */
typedef struct VVCCudaDbkEdge {
    uint16_t tc;
    uint8_t  beta;
    uint8_t  max_len;          /* max_len_p | max_len_q << 4 */
} VVCCudaDbkEdge;

/* Edge maps of one picture.  Index [dir][c]: dir 0 = horizontal edges (lf.filter_*[0]),
 * dir 1 = vertical edges (lf.filter_*[1]), c = plane.  In plane c's own sample units, with
 * G = 4 (luma) or 8 (chroma) the edge grid and L = 4 (luma) or 4 >> shift (chroma) the segment
 * length along the edge:
 *   vertical   entry (ex, sy): edge at x = ex * G, lines   [sy * L, sy * L + L)
 *   horizontal entry (sx, ey): edge at y = ey * G, columns [sx * L, sx * L + L)
 * stored row-major with `pitch[dir][c]` entries per row; picture k of a ring starts
 * `size[dir][c]` entries after picture k-1.  Entries with ex == 0 / ey == 0 (picture border)
 * are ignored, as in the reference (vvc_filter.c:897,965). */
typedef struct VVCCudaDeblockMaps {
    const VVCCudaDbkEdge *edge[2][3];
    int32_t pitch[2][3];
    int32_t rows[2][3];
    int64_t size[2][3];
} VVCCudaDeblockMaps;

/* dir: 1 = all vertical edges of the picture (DEBLOCK_V), 0 = all horizontal edges (DEBLOCK_H);
 * the reference's task graph requires V before H (vvc_thread.c:159-167).  Out of place
 * (dst != src): every edge reads unfiltered samples of `src`; for edge sets that obey the
 * reference's filter-length rules (vvc_filter.c:374-397) this equals the reference's in-place
 * sequential walk.  The maps struct is passed by value semantics (host memory), the arrays it
 * points to are device memory (or host memory for the _host entry). */
int vvc_cuda_deblock_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                           const VVCCudaDeblockMaps *maps, int dir);
/* both passes: src -> tmp (vertical) -> dst (horizontal) */
int vvc_cuda_deblock_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                const VVCCudaDeblockMaps *maps);

/* ------------------------------------------------------------------------------------------
 * Deblocking parameters on the device: boundary strengths, maximum filter lengths and the QP -> beta / tc mapping with
 * the luma-adaptive offset (replaces vvc_deblock_bs with its luma / chroma / sub-block passes and boundary_strength,
 * derive_max_filter_length_luma, max_filter_length_chroma, get_qp_y / get_qp_c, TC_CALC and the per-edge part of
 * ff_vvc_deblock_vertical / _horizontal, libavcodec/vvc/vvc_filter.c:308-1003, and lf.ladf_level,
 * vvc_filter_template.c:788-804).  Input is what the parser knows per transform unit and per motion rectangle - the
 * content of the reference's per-4x4 side tables (fc->tab.tb_pos_x0 / tb_width / tu_coded_flag / pcmf / qp / cb_pos_x /
 * cb_width / msf / iaf / mvf) in list form; the device scatters it into those tables and fills the VVCCudaDeblockMaps of
 * one direction, which vvc_cuda_deblock_frame then consumes.  4:2:0 and 4:0:0.
 * ---------------------------------------------------------------------------------------- */
#define VVC_CUDA_DBK_TU_LUMA      1   /* planes: the unit has a luma transform block (single tree, dual-tree luma)   */
#define VVC_CUDA_DBK_TU_CHROMA    2   /* planes: it has chroma transform blocks (single tree, dual-tree chroma)      */
#define VVC_CUDA_DBK_CBF_Y        1   /* flags: tu_y_coded_flag / tu_cb_coded_flag / tu_cr_coded_flag                */
#define VVC_CUDA_DBK_CBF_CB       2
#define VVC_CUDA_DBK_CBF_CR       4
#define VVC_CUDA_DBK_JOINT        8   /* tu_joint_cbcr_residual_flag                                                 */
#define VVC_CUDA_DBK_BDPCM_Y      16  /* fc->tab.pcmf[LUMA] (set for BDPCM blocks, vvc_ctu.c:1246-1247)              */
#define VVC_CUDA_DBK_BDPCM_C      32  /* fc->tab.pcmf[CHROMA]                                                        */
#define VVC_CUDA_DBK_CU_SUBBLOCK  1   /* cu_flags: merge_subblock_flag | inter_affine_flag (fc->tab.msf / iaf)       */

typedef struct VVCCudaDbkTU {
    uint16_t x0, y0;              /* luma position of the transform unit                                         */
    uint8_t  log2_w, log2_h;      /* luma size (4..64); the chroma blocks are that shifted by the subsampling    */
    uint8_t  planes;              /* VVC_CUDA_DBK_TU_*                                                            */
    uint8_t  flags;               /* VVC_CUDA_DBK_CBF_* | JOINT | BDPCM_*                                          */
    int8_t   qp[3];               /* fc->tab.qp[LUMA / CB / CR] at this unit (ff_vvc_get_qPy, get_qPc)            */
    uint8_t  cu_flags;            /* VVC_CUDA_DBK_CU_*                                                            */
    uint8_t  cu_dx, cu_dy;        /* (x0 - CbPosX) >> 2, (y0 - CbPosY) >> 2: where the unit sits in its coding block */
    uint8_t  cb_log2_w, cb_log2_h;/* CbWidth, CbHeight of the luma coding block                                   */
    uint8_t  pic, reserved[3];
} VVCCudaDbkTU;                   /* 20 bytes */

/* motion of a rectangle of 4x4 units, as ff_vvc_set_mvf / ff_vvc_set_intra_mvf leave it in fc->tab.mvf */
typedef struct VVCCudaDbkMvf {
    uint16_t x0, y0;              /* luma position                                                              */
    uint8_t  w4, h4;              /* size in 4-sample units                                                     */
    uint8_t  pred_flag;           /* PF_INTRA 0, PF_L0 1, PF_L1 2, PF_BI 3                                       */
    uint8_t  ciip_flag;
    int16_t  ref_pic[2];          /* identity of the picture each list refers to (rpl[l].list[ref_idx[l]]: a POC  */
                                  /* or DPB slot number - only compared for equality)                            */
    int32_t  mv[2][2];            /* [list][x, y], 1/16 sample                                                  */
    uint8_t  pic, reserved[3];
} VVCCudaDbkMvf;                  /* 32 bytes */

typedef struct VVCCudaDbkCtb {    /* per CTB, raster order, per picture                                         */
    int8_t  beta_offset[3];       /* DBParams (vvc_ctu.h), per plane                                             */
    int8_t  tc_offset[3];
    uint8_t no_left, no_top;      /* the CTB's left / upper border is a slice or tile border that must not be filtered
                                     across (vvc_filter.c:498-506)                                               */
} VVCCudaDbkCtb;                  /* 8 bytes */

typedef struct VVCCudaDbkParams { /* host memory                                                                 */
    int32_t  qp_bd_offset;
    int32_t  ladf_enabled;        /* sps_ladf_enabled_flag                                                       */
    int32_t  num_ladf_intervals;  /* sps->num_ladf_intervals (2..5)                                              */
    int32_t  ladf_lowest_interval_qp_offset;
    int32_t  ladf_qp_offset[4];   /* sps_ladf_qp_offset[]                                                        */
    int32_t  ladf_interval_lower_bound[5];
} VVCCudaDbkParams;

/* frame: the picture the pass of direction `dir` will read (luma levels for LADF: the V pass reads the reconstructed
 * picture, the H pass the output of the V pass).  tus, mvfs, ctbs: device memory, covering the whole ring.  maps: host
 * struct whose edge[dir][*] device arrays (sized by pitch / rows / size as for vvc_cuda_deblock_frame) are filled. */
int vvc_cuda_deblock_params_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaDbkTU *tus, int n_tus,
                                  const VVCCudaDbkMvf *mvfs, int n_mvfs, const VVCCudaDbkCtb *ctbs,
                                  const VVCCudaDbkParams *params, const VVCCudaDeblockMaps *maps, int dir);

/* ------------------------------------------------------------------------------------------
 * SAO stage (replaces ff_vvc_sao_filter, libavcodec/vvc/vvc_filter.c:154-298, and the table
 * entries sao.band_filter[9] / edge_filter[9] / edge_restore[2],
 * libavcodec/h26x/h2656_sao_template.c:24-215)
 * ---------------------------------------------------------------------------------------- */
/* per-CTB parameters; mirrors SAOParams (libavcodec/vvc/vvc_ctu.h:440-451) plus the
 * "unfilterable edge" flags ff_vvc_sao_filter derives per CTB (vvc_filter.c:163-213). */
typedef struct VVCCudaSAOCtb {
    uint8_t type_idx[3];       /* 0 off, 1 band, 2 edge                                   */
    uint8_t band_position[3];
    uint8_t eo_class[3];       /* 0 horizontal, 1 vertical, 2 135 degrees, 3 45 degrees   */
    uint8_t restore;           /* 1: use the edge_restore[1] rules with the flags below   */
    uint8_t no_filter;         /* bit0-1 vert_edge[0..1], bit2-3 horiz_edge[0..1], bit4-7 diag_edge[0..3] */
    uint8_t reserved;
    int16_t offset_val[3][5];  /* SaoOffsetVal, [0] is 0 in a conforming stream            */
} VVCCudaSAOCtb;

/* dst != src; ctbs: ctb_count entries per picture of the ring, raster order. */
int vvc_cuda_sao_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                       const VVCCudaSAOCtb *ctbs);
int vvc_cuda_sao_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                            const VVCCudaSAOCtb *ctbs);

/* ------------------------------------------------------------------------------------------
 * In-loop chain: DEBLOCK_V -> DEBLOCK_H -> SAO -> ALF in one call (the tail of the reference's
 * per-CTU stage list, libavcodec/vvc/vvc_thread.c:41-51, run stage by stage over the picture,
 * which its dependency scores allow - SURVEY.md 3.3).  Intermediate pictures live in the
 * context's scratch area.
 * ---------------------------------------------------------------------------------------- */
/* the list inputs of vvc_cuda_deblock_params_frame for a picture ring (host struct, device arrays) */
typedef struct VVCCudaDbkSide {
    const VVCCudaDbkTU     *tus;
    const VVCCudaDbkMvf    *mvfs;
    const VVCCudaDbkCtb    *ctbs;
    const VVCCudaDbkParams *params;       /* host memory */
    int32_t                 n_tus, n_mvfs;
} VVCCudaDbkSide;

typedef struct VVCCudaInloopDesc {
    const VVCCudaDeblockMaps *deblock;        /* host struct; its arrays follow the entry's memory space */
    const VVCCudaSAOCtb      *sao;
    const VVCCudaALFCtb      *alf;
    const VVCCudaALFSets     *alf_sets;
    int32_t                   alf_sets_per_frame;
    int32_t                   reserved;
    const VVCCudaDbkSide     *dbk_side;       /* device entries only, optional: the deblocking parameters are derived on
                                                 the device (vvc_cuda_deblock_params_frame before each pass, maps in the
                                                 context's scratch area) and `deblock` is not read                     */
} VVCCudaInloopDesc;

int vvc_cuda_inloop_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                          const VVCCudaInloopDesc *desc);
/* host pictures and descriptors: copies in (pictures + all maps), runs the four stages, copies the
 * finished pictures out.  With batch > 1 the copies of picture k+1 overlap the kernels of picture k. */
int vvc_cuda_inloop_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                               const VVCCudaInloopDesc *desc);

/* ------------------------------------------------------------------------------------------
 * Residual stage: inverse LFNST + inverse transform + add_residual for a list of transform
 * blocks (replaces the per-TB tail of itransform(), libavcodec/vvc/vvc_intra.c:432-478, i.e.
 * ilfnst_transform :65-127 / ff_vvc_inv_lfnst_1d vvc_itx_1d.c:708-721, the table entries
 * itx.itx[trh][trv][log2w][log2h] (vvcdsp.c:94-195, vvc_itx_1d.c:70-706), itx.transform_bdpcm,
 * itx.add_residual, itx.add_residual_joint (vvcdsp_template.c:32-95)).
 * vvc_cuda_itx_frame takes coefficients as the reference stores them after dequant(): one dense
 * row-major int32[h][w] per TB.  vvc_cuda_itx_frame_q (below) takes the compact window layout and
 * can run dequant() on the device in front of the transform.
 * ---------------------------------------------------------------------------------------- */
#define VVC_CUDA_TB_TS              1   /* transform skip: no LFNST / transform (tb->ts)                 */
#define VVC_CUDA_TB_BDPCM           2   /* run itx.transform_bdpcm first (horizontal accumulate)         */
#define VVC_CUDA_TB_BDPCM_VERT      4   /* ... vertical accumulate                                       */
#define VVC_CUDA_TB_JOINT           8   /* joint CbCr: also add (res * joint_sign) >> joint_shift to plane joint_c_idx */
#define VVC_CUDA_TB_STORE_RESIDUAL 16   /* write the residual back over the coefficients (what the reference's
                                           in-place itx leaves in tb->coeffs) instead of adding it to the picture */

typedef struct VVCCudaTB {
    uint32_t coeff_offset;   /* first coefficient of this TB, in int32 units from the buffer start     */
    uint16_t x0, y0;         /* top-left sample in plane c_idx (that plane's own sample units)          */
    uint8_t  log2_w, log2_h; /* 0..6 (1 x N and N x 1 blocks as in the reference's 1-D cells)           */
    uint8_t  c_idx;
    uint8_t  trh, trv;       /* enum TxType (vvcdsp.h:30-35): 0 DCT2, 1 DST7, 2 DCT8                    */
    uint8_t  nzw, nzh;       /* max_scan_x + 1, max_scan_y + 1, as passed to itx.itx[]                  */
    uint8_t  flags;          /* VVC_CUDA_TB_*                                                           */
    uint8_t  lfnst;          /* 0 off; else bits0-1 lfnst_idx (1|2), bits2-3 transform set
                                (ff_vvc_lfnst_tr_set_index[predModeIntra], 1 for wide-angle < 0),
                                bit4 transpose (predModeIntra > 34), bit5 nonZeroSize == 8             */
    int8_t   joint_sign;
    uint8_t  joint_shift;
    uint8_t  joint_c_idx;
    uint8_t  pic;            /* picture of the ring this TB belongs to                                  */
    uint8_t  reserved;
    uint16_t chroma_scale;   /* LMCS chroma residual scaling (lmcs_scale_chroma, vvc_intra_template.c:431-448):
                                0 = off; with VVCCudaCoeffs.lmcs_scales == NULL the scale itself
                                (lmcs->chroma_scale_coeff[i]), else 1 + the index of this block's VPDU in that
                                array (filled on the device by vvc_cuda_lmcs_chroma_scale)                 */
} VVCCudaTB;                 /* 24 bytes */

/* frame: prediction in, reconstruction out (in place - TBs are disjoint).  coeffs: dequantised
 * coefficients; only written when a TB has VVC_CUDA_TB_STORE_RESIDUAL. */
int vvc_cuda_itx_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int32_t *coeffs,
                       const VVCCudaTB *tbs, int n_tbs, int log2_transform_range);
int vvc_cuda_itx_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int32_t *coeffs, size_t n_coeffs,
                            const VVCCudaTB *tbs, int n_tbs, int log2_transform_range);

/* ------------------------------------------------------------------------------------------
 * Compact coefficient input and fused dequantisation (replaces dequant() with derive_qp's shift,
 * derive_scale, derive_scale_m and scale_coeff, libavcodec/vvc/vvc_intra.c:277-417, in front of the
 * residual stage; the reference's order transform_bdpcm -> dequant -> LFNST -> itx -> add_residual,
 * vvc_intra.c:453-470, is kept).
 *
 * Coefficient layouts (VVCCudaTB.coeff_offset counts elements of the layout's own type):
 *   VVC_CUDA_COEFF_DENSE32   int32[h][w] per TB, as tb->coeffs
 *   VVC_CUDA_COEFF_WINDOW16  int16[nzh][nzw] per TB: only the window the residual coder wrote,
 *                            rows 0..max_scan_y, columns 0..max_scan_x, row pitch nzw.  Quantised levels
 *                            and dequantised coefficients are both 16-bit at log2_transform_range 15
 *                            (CoeffMinY..CoeffMaxY), so the layout is lossless there; it is rejected for
 *                            larger ranges.  4-8x fewer bytes over PCIe than DENSE32.
 * ---------------------------------------------------------------------------------------- */
#define VVC_CUDA_COEFF_DENSE32   0
#define VVC_CUDA_COEFF_WINDOW16  1

/* per-TB quantisation parameters, entry i belongs to tbs[i] */
typedef struct VVCCudaTBQuant {
    uint8_t qp;          /* tb->qp as derive_qp leaves it (offsets applied, clipped), vvc_intra.c:277-308   */
    uint8_t dep_quant;   /* sh_dep_quant_used_flag of the slice                                            */
    uint8_t sl_id;       /* 0: flat scaling (ff_vvc_default_scale_m, m = 16); else 1 + the scaling matrix id
                            of Table 38 (ids[][][] in derive_scale_m, vvc_intra.c:341-355)                */
    uint8_t reserved;
} VVCCudaTBQuant;        /* 4 bytes */

/* mirrors VVCScalingList (libavcodec/vvc/vvc_ps.h:187-190) */
typedef struct VVCCudaScalingList {
    uint8_t matrix_rec[28][64];   /* ScalingMatrixRec[id][8 * y + x]      */
    uint8_t dc_rec[14];           /* ScalingMatrixDcRec[id - 14]          */
    uint8_t reserved[2];
} VVCCudaScalingList;

typedef struct VVCCudaCoeffs {
    void                     *data;      /* int32_t* or int16_t*, by `format`                              */
    size_t                    n;         /* elements in data (read by the _host entries only)              */
    int32_t                   format;    /* VVC_CUDA_COEFF_*                                               */
    int32_t                   reserved;
    const VVCCudaTBQuant     *quant;     /* NULL: data holds dequantised coefficients; else quantised levels
                                            (TransCoeffLevel) and the device runs dequant()                */
    const VVCCudaScalingList *scaling;   /* the picture's scaling list (fc->ps.sl); may be NULL when every
                                            sl_id is 0                                                     */
    const uint16_t           *lmcs_scales; /* optional: per-VPDU chroma residual scales, see VVCCudaTB.chroma_scale */
} VVCCudaCoeffs;

/* LMCS chroma residual scaling, the derivation (replaces lmcs_derive_chroma_scale, vvc_intra_template.c:389-428): per
 * 64x64 VPDU (min(CtbSizeY, 64)) the average of the reconstructed luma samples left of and above it - in the mapped
 * domain, i.e. after the luma residual and BEFORE the inverse LMCS stage - picks the bin whose chroma_scale_coeff scales
 * the chroma residuals of every CU that starts in the VPDU.  The luma blocks of a picture therefore go through the
 * residual stage first, then this call, then the chroma blocks with VVCCudaCoeffs.lmcs_scales = scales. */
typedef struct VVCCudaLmcsVpdu {
    uint16_t x, y;            /* luma position of the VPDU (cu->x0 & ~(size - 1), cu->y0 & ~(size - 1))              */
    uint8_t  avail_l, avail_t;/* ff_vvc_get_left_available / _top_available(lc, x, y, 1, 0) != 0                      */
    uint8_t  pic, reserved;
} VVCCudaLmcsVpdu;            /* 8 bytes */
typedef struct VVCCudaLmcsParams {  /* of VVCLMCS, libavcodec/vvc/vvc_ps.h:192-202 */
    uint16_t pivot[17];
    uint16_t chroma_scale_coeff[16];
    uint8_t  min_bin_idx, max_bin_idx;
} VVCCudaLmcsParams;
/* frame: the picture (ring) being reconstructed; vpdus, params, scales (n entries out): device memory */
int vvc_cuda_lmcs_chroma_scale(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaLmcsVpdu *vpdus, int n,
                               const VVCCudaLmcsParams *params, uint16_t *scales);

/* Same stage as vvc_cuda_itx_frame.  VVC_CUDA_TB_STORE_RESIDUAL needs DENSE32. */
int vvc_cuda_itx_frame_q(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *coeffs,
                         const VVCCudaTB *tbs, int n_tbs, int log2_transform_range);
int vvc_cuda_itx_frame_q_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *coeffs,
                              const VVCCudaTB *tbs, int n_tbs, int log2_transform_range);

/* ------------------------------------------------------------------------------------------
 * LMCS stage (replaces lmcs.filter, libavcodec/vvc/vvc_filter_template.c:25-36, as driven by
 * ff_vvc_lmcs_filter, vvc_filter.c:1322-1332 (inverse LUT per CTU) and by predict_inter,
 * vvc_inter.c:888-891 (forward LUT on the luma of inter CUs))
 * ---------------------------------------------------------------------------------------- */
typedef struct VVCCudaRect {
    uint16_t x, y, w, h;     /* luma samples */
    uint16_t pic;            /* picture of the ring */
    uint16_t reserved;
} VVCCudaRect;

/* luma[x] = lut[luma[x]] in place on every CTB whose ctb_enable byte is non-zero (NULL = all CTBs;
 * ctb_count bytes per picture of the ring).  lut: (1 << bit_depth) uint16 entries
 * (VVCLMCS.inv_lut / fwd_lut, libavcodec/vvc/vvc_ps.h:192-199). */
int vvc_cuda_lmcs_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const uint16_t *lut, const uint8_t *ctb_enable);
/* same mapping on a list of luma rectangles (the inter CUs of a picture) */
int vvc_cuda_lmcs_rects(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const uint16_t *lut,
                        const VVCCudaRect *rects, int n_rects);
int vvc_cuda_lmcs_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const uint16_t *lut, const uint8_t *ctb_enable);

/* ------------------------------------------------------------------------------------------
 * Inter prediction stage (replaces the pixel work of ff_vvc_predict_inter, libavcodec/vvc/
 * vvc_inter.c:899-913: pred_regular_blk :782-811 incl. dmvr_mv_refine :685-748, pred_affine_blk
 * :828-873, pred_gpm_blk :466-521, and the table entries inter.put / put_uni / put_uni_w / avg /
 * w_avg / put_gpm / bdof_fetch_samples / fetch_samples / prof_grad_filter / apply_prof* /
 * apply_bdof / sad / dmvr, libavcodec/h26x/h2656_inter_template.c:29-577,
 * libavcodec/vvc/vvc_inter_template.c:25-436, libavcodec/vvc/vvcdsp.c:29-65).
 *
 * The host (CABAC + MV derivation) emits one prediction-block record per motion-compensated block
 * of at most 16x16 luma samples: DMVR/BDOF sub-blocks are 16x16 by construction
 * (vvc_inter.c:796-797), larger plain CUs are simply cut into 16x16 pieces (MC is per-sample),
 * affine CUs give one luma record per 4x4 sub-block plus one chroma record per 2x2 group
 * (:862-868).  Reference pictures live in one ring (the DPB); ref[] is the ring slot.
 * ---------------------------------------------------------------------------------------- */
#define VVC_CUDA_PB_LUMA        1    /* planes: predict luma                                      */
#define VVC_CUDA_PB_CHROMA      2    /* planes: predict Cb and Cr (4:2:0: (w/2) x (h/2) at (x0/2, y0/2)) */

#define VVC_CUDA_PB_DMVR        1    /* flags: pu->dmvr_flag (bi-pred only, w*h >= 128)             */
#define VVC_CUDA_PB_BDOF        2    /* pu->bdof_flag (bi-pred only)                                */
#define VVC_CUDA_PB_PROF0       4    /* pu->cb_prof_flag[0] (4x4 affine luma blocks)                */
#define VVC_CUDA_PB_PROF1       8    /* pu->cb_prof_flag[1]                                         */
#define VVC_CUDA_PB_GPM        16    /* geometric partition blend of two uni predictions (mv[0], mv[1]) */
#define VVC_CUDA_PB_WEIGHTED   32    /* the reference's weight_flag is set (derive_weight(_uni), vvc_inter.c:129-177): wp valid */

typedef struct VVCCudaPB {
    uint16_t x0, y0;          /* luma position of the block                                         */
    uint8_t  w, h;            /* luma size: 4, 8 or 16 each                                         */
    uint8_t  planes;          /* VVC_CUDA_PB_LUMA | VVC_CUDA_PB_CHROMA                              */
    uint8_t  pred_flag;       /* PF_L0 = 1, PF_L1 = 2, PF_BI = 3 (vvc_ctu.h:215-220).  Entry i of mv[] / ref[]
                                 is list i as in MvField; a GPM block uses both entries (gpm_mv[0], gpm_mv[1]) */
    uint8_t  ref[2];          /* DPB ring slot of the reference picture of each used entry          */
    uint8_t  pic;             /* destination picture slot                                           */
    uint8_t  flags;           /* VVC_CUDA_PB_*                                                      */
    int32_t  mv[2][2];        /* [entry][x,y] in 1/16 luma sample                                   */
    uint8_t  filt;            /* luma filter set: 0 regular, 1 alternative half-pel (hpel_if_idx), 2 affine */
    uint8_t  bcw_idx;         /* 0 = off, else w1 = {4,5,3,10,-2}[bcw_idx], denom 2 (vvc_inter.c:31,161-165) */
    uint16_t wp;              /* index into the weighted-prediction table (VVC_CUDA_PB_WEIGHTED)     */
    uint16_t prof;            /* index into the PROF table: diff_mv of the CU                        */
    int16_t  gpm_step_x;      /* GPM: weight pointer steps for luma (+-1, +-112), vvc_inter.c:491-502; */
    int16_t  gpm_step_y;      /*      chroma uses twice these steps from the same first weight        */
    uint16_t reserved;
    int32_t  gpm_weights;     /* GPM: index of this block's first weight in the 6 x 112 x 112 table   */
} VVCCudaPB;                  /* 44 bytes */

/* explicit weights resolved per (ref_idx_l0, ref_idx_l1) pair, mirrors PredWeightTable
 * (libavcodec/vvc/vvc_ps.h:135-143) as read by derive_weight(_uni), vvc_inter.c:129-177 */
typedef struct VVCCudaWP {
    int16_t weight[2][3];     /* [list][plane] */
    int16_t offset[2][3];
    uint8_t log2_denom[2];    /* luma, chroma */
    uint8_t reserved[2];
} VVCCudaWP;                  /* 28 bytes */

/* PROF displacement of one affine CU, mirrors pu->diff_mv_x/y (vvc_ctu.h:268-270) */
typedef struct VVCCudaProf {
    int16_t diff_mv_x[2][16];
    int16_t diff_mv_y[2][16];
} VVCCudaProf;                /* 128 bytes */

/* per block outputs of DMVR, what set_dmvr_info() stores for later pictures (vvc_inter.c:750-762) */
typedef struct VVCCudaDmvrOut {
    int32_t mv[2][2];         /* refined motion vectors */
    int32_t min_sad;
    int32_t bdof_applied;     /* sb_bdof_flag after the min_sad < 2wh test (:745) */
} VVCCudaDmvrOut;             /* 24 bytes */

/* dst: picture ring being predicted (record i writes slot pbs[i].pic); refs: the DPB ring (must not
 * alias the written slots).  dmvr_out: optional, n_pbs entries (only DMVR records are written).
 * CIIP blocks are not part of this entry: their inter half needs the intra predictor and is blended
 * by inter.put_ciip in the RECON stage (vvc_intra.c:515-516).  The forward LMCS mapping of inter CUs
 * (vvc_inter.c:888-891) is vvc_cuda_lmcs_rects().  All arrays device memory. */
int vvc_cuda_inter_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *refs,
                         const VVCCudaPB *pbs, int n_pbs, const VVCCudaWP *wp, const VVCCudaProf *prof,
                         VVCCudaDmvrOut *dmvr_out);
/* host arrays and pictures: copies refs + descriptors in, runs, copies the predicted pictures and
 * dmvr_out back */
int vvc_cuda_inter_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *refs,
                              const VVCCudaPB *pbs, int n_pbs, const VVCCudaWP *wp, int n_wp,
                              const VVCCudaProf *prof, int n_prof, VVCCudaDmvrOut *dmvr_out);

/* ------------------------------------------------------------------------------------------
 * Intra leaf predictors and the CIIP blend.  Batched form of the table entries intra.pred_planar /
 * pred_dc / pred_v / pred_h / pred_angular_v / pred_angular_h / pred_mip
 * (libavcodec/vvc/vvc_intra_template.c:686-1000) and inter.put_ciip (vvc_inter_template.c:60-76).
 * The reference lines of a block are prepared by the host exactly as prepare_intra_edge_params()
 * (vvc_intra_template.c:467-592) leaves them in IntraEdgeParams: that step (availability, substitution,
 * [1 2 1] smoothing, the projected negative-index part), the PDPC of planar/DC/H/V, CCLM and the LMCS
 * chroma scaling need the decoder's CTU state and stay with it (entries that take VVCLocalContext*).
 * Blocks of one call must not depend on each other (a wavefront of the RECON stage, or CIIP blocks).
 * ---------------------------------------------------------------------------------------- */
#define VVC_CUDA_INTRA_PLANAR     0
#define VVC_CUDA_INTRA_DC         1
#define VVC_CUDA_INTRA_VERT       2   /* pred_v */
#define VVC_CUDA_INTRA_HORZ       3   /* pred_h */
#define VVC_CUDA_INTRA_ANGULAR_V  4   /* pred_angular_v: mode >= 34 (INTRA_DIAG) */
#define VVC_CUDA_INTRA_ANGULAR_H  5   /* pred_angular_h */
#define VVC_CUDA_INTRA_MIP        6

#define VVC_CUDA_INTRA_PDPC       1   /* flags: need_pdpc argument of pred_angular_* */
#define VVC_CUDA_INTRA_MIP_TRANSPOSED 2

typedef struct VVCCudaIntraPB {
    uint16_t x0, y0;          /* top-left sample, in plane c_idx's own units                          */
    uint8_t  w, h;            /* w 4..64, h 2..64, powers of two (MIP: both 4..64)                    */
    uint8_t  c_idx, pic;
    uint8_t  kind;            /* VVC_CUDA_INTRA_*                                                      */
    int8_t   mode;            /* angular: predModeIntra after wide-angle mapping (-14..80); MIP: mode id */
    uint8_t  ref_idx;         /* intra_luma_ref_idx (angular)                                          */
    uint8_t  filter_flag;     /* IntraEdgeParams.filter_flag (angular, luma: 0 cubic, 1 gauss)         */
    uint8_t  flags;           /* VVC_CUDA_INTRA_PDPC | VVC_CUDA_INTRA_MIP_TRANSPOSED                    */
    uint8_t  reserved[3];
    uint32_t top, left;       /* index of top[0] / left[0] in the edge sample buffer; the samples the
                                 reference reads at negative indices (top[-1 - ref_idx] ...) lie before */
} VVCCudaIntraPB;             /* 24 bytes */

/* frame: picture (ring) predicted into.  edges: 16-bit reference samples.  All arrays device memory. */
int vvc_cuda_intra_leaf_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraPB *pbs, int n_pbs,
                              const uint16_t *edges);
int vvc_cuda_intra_leaf_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraPB *pbs, int n_pbs,
                                   const uint16_t *edges, size_t n_edges);

/* ------------------------------------------------------------------------------------------
 * Full intra prediction: the table entries intra.intra_pred (vvc_intra_template.c:595-683 with
 * prepare_intra_edge_params :467-592 and ref_filter :450-465) and intra.intra_cclm_pred (:29-375) on the device -
 * reference-line gather from the reconstructed picture, substitution of unavailable samples, the [1 2 1] smoothing,
 * the projected / extended part of the angular reference, wide-angle mode mapping (ff_vvc_wide_angle_mode_mapping,
 * vvc_intra.c:693-715), the predictors, and the PDPC of planar / DC / H / V; CCLM: luma down-sampling, the four
 * selected neighbour positions, min / max pairs, the a / b / k derivation and the linear predictor for Cb and Cr.
 * What a record states about the decoder's CTU state is what ff_vvc_get_left_available / _top_available
 * (vvc_intra.c:591-648) return with an unbounded target size - the callee takes min(target, that) - and
 * lc->na.cand_up_left.  Blocks of one call must not depend on each other (one wavefront).
 * ---------------------------------------------------------------------------------------- */
#define VVC_CUDA_INTRA_KIND_PRED  0   /* intra_pred, regular modes 0..66                               */
#define VVC_CUDA_INTRA_KIND_MIP   1   /* intra_pred, matrix-based (fc->tab.imf set)                    */
#define VVC_CUDA_INTRA_KIND_CCLM  2   /* intra_cclm_pred: predicts Cb and Cr; pred_mode 81 / 82 / 83   */

#define VVC_CUDA_INTRA_F_ISP          1   /* cu->isp_split_type != ISP_NO_SPLIT (luma)                 */
#define VVC_CUDA_INTRA_F_BDPCM        2   /* cu->bdpcm_flag[c_idx]                                     */
#define VVC_CUDA_INTRA_F_MIP_TRANSP   4   /* fc->tab.imtf                                              */
#define VVC_CUDA_INTRA_F_UP_LEFT      8   /* lc->na.cand_up_left                                       */
#define VVC_CUDA_INTRA_F_LUMA_AVAIL_T 16  /* CCLM: ff_vvc_get_top_available(lc, x0, y0, 1, 0) != 0     */
#define VVC_CUDA_INTRA_F_LUMA_AVAIL_L 32  /* CCLM: ff_vvc_get_left_available(lc, x0, y0, 1, 0) != 0    */
#define VVC_CUDA_INTRA_F_COLLOCATED   64  /* CCLM: sps_chroma_vertical_collocated_flag                 */

typedef struct VVCCudaIntraBlk {
    uint16_t x0, y0;          /* top-left sample in plane c_idx's own units (CCLM: chroma units)       */
    uint8_t  w, h;            /* in the same units: 2..64 (4..64 for MIP)                              */
    uint8_t  c_idx, pic;      /* CCLM: c_idx = 1 (both chroma planes are predicted)                    */
    uint8_t  kind;            /* VVC_CUDA_INTRA_KIND_*                                                  */
    uint8_t  pred_mode;       /* cu->intra_pred_mode_y / _c before wide-angle mapping; MIP: the mode id */
    uint8_t  ref_idx;         /* cu->intra_luma_ref_idx (0 for chroma)                                 */
    uint8_t  flags;           /* VVC_CUDA_INTRA_F_*                                                     */
    uint8_t  cb_w, cb_h;      /* cu->cb_width / cb_height (luma samples; read for ISP blocks only)     */
    uint8_t  avail_left;      /* ff_vvc_get_left_available(lc, x, y, 255, c_idx)                       */
    uint8_t  avail_top;       /* ff_vvc_get_top_available(lc, x, y, 255, c_idx)                        */
} VVCCudaIntraBlk;            /* 16 bytes */

/* frame: the picture (ring) being reconstructed: reference samples are read from it, the prediction is written
 * into it.  blks: device memory. */
int vvc_cuda_intra_pred_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, int n_blks);
int vvc_cuda_intra_pred_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, int n_blks);

/* All-intra reconstruction of a picture (ring): prediction and residual alternate wavefront by wavefront.  The
 * blocks and the transform blocks are sorted by wavefront; wave g owns blks[blk_end[g-1] .. blk_end[g]) and
 * tbs[tb_end[g-1] .. tb_end[g]) (blk_end / tb_end: HOST arrays of n_waves running totals).  A block belongs to the
 * first wavefront in which every block whose samples it reads (reference lines, CCLM's luma) has been reconstructed;
 * its transform blocks belong to the same wavefront.  Everything else as vvc_cuda_intra_pred_frame / _itx_frame_q. */
int vvc_cuda_intra_recon_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, const int32_t *blk_end,
                               const VVCCudaCoeffs *coeffs, const VVCCudaTB *tbs, const int32_t *tb_end, int n_waves,
                               int log2_transform_range);

/* The same reconstruction driven by dependencies on the device, ONE launch per picture ring: the steps of a decoder in
 * decoding order (per coding unit a luma step and a chroma step; step s predicts blks[blk_end[s-1] .. blk_end[s]) and then
 * adds the residual of tbs[tb_end[s-1] .. tb_end[s])), blk_end / tb_end in DEVICE memory.  Persistent CTAs draw the steps
 * in order and wait, per block, for exactly the neighbouring samples its availability counts name (a progress map of 4x4
 * units), so no wave numbers are needed.  Returns after the launch has finished (the call checks the kernel's watchdog:
 * VVC_CUDA_ERR_ARG when a step names samples no earlier step reconstructs). */
int vvc_cuda_intra_recon_frame_ordered(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaIntraBlk *blks, const int32_t *blk_end,
                                       const VVCCudaCoeffs *coeffs, const VVCCudaTB *tbs, const int32_t *tb_end, int n_steps,
                                       int n_blks, int n_tbs, int log2_transform_range);

typedef struct VVCCudaCiip {
    uint16_t x0, y0;          /* in plane c_idx's own units */
    uint8_t  w, h, c_idx, pic;
    uint8_t  intra_weight;    /* 1..3 (ciip_derive_intra_weight, vvc_inter.c:523-543)                  */
    uint8_t  reserved[3];
} VVCCudaCiip;                /* 12 bytes */

/* dst holds the intra prediction of each block and receives (dst * wi + inter * (4 - wi) + 2) >> 2;
 * inter: a picture (ring) of the same geometry holding the inter prediction at the same positions. */
int vvc_cuda_ciip_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *inter,
                        const VVCCudaCiip *blocks, int n_blocks);
int vvc_cuda_ciip_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *inter,
                             const VVCCudaCiip *blocks, int n_blocks);

/* ------------------------------------------------------------------------------------------
 * Whole-picture reconstruction: INTER -> RECON (residual) -> LMCS -> DEBLOCK_V -> DEBLOCK_H -> SAO ->
 * ALF, the reference's per-CTU stage list (libavcodec/vvc/vvc_thread.c:41-51) run stage by stage
 * over the picture (7 stage calls; 17 kernel launches for a 10-bit 4:2:0 picture ring: inter 8, residual 4, the other stages 1 each; + 1 when forward-LMCS rectangles are given).
 * ---------------------------------------------------------------------------------------- */
typedef struct VVCCudaReconDesc {
    /* INTER */
    const VVCCudaPB   *pbs;
    const VVCCudaWP   *wp;
    const VVCCudaProf *prof;
    VVCCudaDmvrOut    *dmvr_out;          /* optional */
    int32_t            n_pbs, n_wp, n_prof;
    int32_t            log2_transform_range;
    /* forward LMCS on the luma of inter CUs (vvc_inter.c:888-891); optional */
    const uint16_t    *lmcs_fwd_lut;
    const VVCCudaRect *lmcs_rects;
    int32_t            n_lmcs_rects;
    /* RECON: residual of every TB */
    int32_t            n_tbs;
    int32_t           *coeffs;            /* int16_t* when coeff_format == VVC_CUDA_COEFF_WINDOW16 */
    size_t             n_coeffs;          /* elements; only read by the _host entry */
    const VVCCudaTB   *tbs;
    int32_t            coeff_format;      /* VVC_CUDA_COEFF_* */
    uint32_t           ref_slots;         /* _host entry: bit s set = this picture reads DPB ring slot s (the slice's
                                             reference picture lists); host reference pictures are then copied in right
                                             before the first picture that needs them, overlapping earlier kernels.
                                             0 = unknown: the whole ring goes up before the first picture */
    const VVCCudaTBQuant     *quant;      /* optional: n_tbs entries, dequant() on the device */
    const VVCCudaScalingList *scaling;    /* optional */
    /* LMCS inverse mapping per CTU; NULL lut = stage skipped (sh_lmcs_used_flag == 0) */
    const uint16_t    *lmcs_inv_lut;
    const uint8_t     *lmcs_ctb_enable;
    /* DEBLOCK_V, DEBLOCK_H, SAO, ALF */
    VVCCudaInloopDesc  inloop;
    /* _host entry: set by vvc_cuda_recon_arena_bind() - all arrays above live in this one (pinned) block, in the device
     * slot's own layout, and go up as one copy.  NULL = arrays anywhere, one copy each. */
    const void        *arena;
    size_t             arena_bytes;
    /* LMCS chroma residual scaling (optional, ph_chroma_residual_scale_flag): with n_lmcs_vpdus > 0 the TB list holds its
     * n_luma_tbs luma blocks first; the stage then runs luma residuals -> vvc_cuda_lmcs_chroma_scale over lmcs_vpdus ->
     * chroma residuals, whose VVCCudaTB.chroma_scale is 1 + the index of the block's VPDU in lmcs_vpdus. */
    const VVCCudaLmcsVpdu   *lmcs_vpdus;
    const VVCCudaLmcsParams *lmcs_params;
    int32_t            n_lmcs_vpdus, n_luma_tbs;
} VVCCudaReconDesc;

/* Device memory everywhere.  cur: the picture (ring) being reconstructed: prediction, residual and
 * LMCS happen in place in it; out: the filtered output picture (ring); refs: the DPB ring.
 * The record lists of `desc` cover the whole ring (their `pic` fields select the picture). */
int vvc_cuda_recon_frame(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *cur,
                         const VVCCudaFrame *refs, const VVCCudaReconDesc *desc);
/* Host memory everywhere.  descs: one entry per picture of the `out` ring, its records use pic == 0.
 * Picture k's descriptors and coefficients are copied in on a copy stream while picture k-1 is being
 * reconstructed and picture k-2 is copied out.  refs: the DPB ring, either host memory (copied in once per
 * call) or device memory (a GPU-resident DPB, e.g. earlier outputs kept in HBM: used in place).  Returns when
 * all output pictures (and dmvr_out arrays) are in host memory. */
/* One pinned arena per picture.  Fill in the descriptor's counts (n_pbs, n_wp, n_prof, n_lmcs_rects, n_tbs, n_coeffs,
 * coeff_format) and maps->size[][] / pitch / rows, ask for the size, allocate (256-byte aligned, ideally pinned), bind:
 * every array pointer of `desc` and `maps` then points into the arena and the caller writes its records through them.
 * Optional arrays may be reset to NULL afterwards.  vvc_cuda_recon_frame_host() uploads a bound descriptor with a single
 * copy.  `out` / `refs` of that entry may each be host or device memory (a device `out` ring stays in HBM: no copy back). */
size_t vvc_cuda_recon_arena_size(const VVCCudaFrame *frame, const VVCCudaReconDesc *desc);
int    vvc_cuda_recon_arena_bind(const VVCCudaFrame *frame, VVCCudaReconDesc *desc, VVCCudaDeblockMaps *maps, void *arena);
int vvc_cuda_recon_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *refs,
                              const VVCCudaReconDesc *descs);
/* The same, returning as soon as everything is queued: the pictures, arenas and dmvr_out arrays must stay valid and
 * untouched until vvc_cuda_sync() (or vvc_cuda_notify()'s report).  Consecutive calls keep the three-stage pipeline
 * full across the call boundary - the copy-out of one call's last pictures runs under the next call's uploads and
 * kernels - which is how a decoder that delivers pictures continuously should drive it (ff_vvc_frame_submit per frame,
 * libavcodec/vvc/vvc_thread.c:432-566, no wait between frames).  Calls must share one geometry and descriptor layout to
 * overlap; a call with a different layout first waits for the earlier ones. */
int vvc_cuda_recon_frame_host_async(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *refs,
                              const VVCCudaReconDesc *descs);

#ifdef __cplusplus
}
#endif
#endif /* VVCDSP_CUDA_H */
