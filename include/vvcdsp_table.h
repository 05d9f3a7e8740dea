/*
 * vvcdsp_table.h - the reference's DSP function-pointer tables as libvvcdsp_cuda.so sees them, and the
 * drop-in hook that overrides entries with CUDA-backed functions of identical signature.
 *
 * Layout contract: struct VVCDSPContext and its seven sub-tables exactly as libavcodec/vvc/vvcdsp.h:48-168
 * declares them (same member order, same array extents, plain C function pointers).  When the reference
 * header has already been included (AVCODEC_VVC_VVCDSP_H) its own declarations are used and this file only
 * adds the hook; otherwise the mirror below is declared.  ffvvc_b200/dsp_tables.py is the same mirror for
 * ctypes; tests check sizeof() of both against the compiled reference.
 *
 *   void ff_vvc_dsp_init_cuda(VVCDSPContext *c, int bit_depth);
 *
 * is called from ff_vvc_dsp_init() after the C (and x86) entries are installed
 * (libavcodec/vvc/vvcdsp.c:254-256, precedent ff_vvc_dsp_init_x86, libavcodec/x86/vvc/vvcdsp_init.c:294-361)
 * and, for bit_depth == 10, overwrites
 *   itx.itx[trh][trv][log2w][log2h]   every valid cell (vvcdsp.c:140-195)
 *   itx.transform_bdpcm               (vvcdsp_template.c:76-95)
 *   itx.add_residual                  (vvcdsp_template.c:32-46)
 *   lmcs.filter                       (vvc_filter_template.c:25-36)
 * Every other entry keeps the value the caller installed (as ff_vvc_dsp_init_x86 leaves what it does not
 * accelerate): their CUDA implementations exist only in batched form (vvcdsp_cuda.h section 2), because a
 * per-block PCIe round trip cannot be made useful and the entries that take VVCLocalContext* cannot be served
 * without the decoder's structures.  Overridden entries take HOST pointers, run the same CUDA kernels as
 * the batched path on a lazily created process-wide context (device VVC_CUDA_DEVICE, default 0), and return
 * after the result is back in the caller's buffer.  They return void like the reference; failures are
 * latched and reported by ff_vvc_dsp_cuda_last_error().  There is no CPU fallback behind them.
 */
#ifndef VVCDSP_TABLE_H
#define VVCDSP_TABLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef AVCODEC_VVC_VVCDSP_H

struct VVCLocalContext;
struct SAOParams;

enum { VVCT_DCT2, VVCT_DST7, VVCT_DCT8, VVCT_N_TX_TYPE };          /* enum VVCTxType,  vvcdsp.h:30-35 */
enum { VVCT_N_TX_SIZE = 7 };                                       /* enum VVCTxSize,  vvcdsp.h:37-46 */

/* inter (vvcdsp.h:48-93) */
typedef void (*vvct_put_fn)(int16_t *dst, const uint8_t *src, ptrdiff_t src_stride, int height,
                            const int8_t *hf, const int8_t *vf, int width);
typedef void (*vvct_put_uni_fn)(uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *src, ptrdiff_t src_stride, int height,
                                const int8_t *hf, const int8_t *vf, int width);
typedef void (*vvct_put_uni_w_fn)(uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *src, ptrdiff_t src_stride, int height,
                                  int denom, int wx, int ox, const int8_t *hf, const int8_t *vf, int width);
typedef void (*vvct_dmvr_fn)(int16_t *dst, const uint8_t *src, ptrdiff_t src_stride, int height,
                             intptr_t mx, intptr_t my, int width);

typedef struct VVCInterDSPContext {
    vvct_put_fn       put[2][7][2][2];          /* [luma, chroma][log2(width) - 1][int, frac][int, frac] */
    vvct_put_uni_fn   put_uni[2][7][2][2];
    vvct_put_uni_w_fn put_uni_w[2][7][2][2];
    void (*avg)(uint8_t *dst, ptrdiff_t dst_stride, const int16_t *src0, const int16_t *src1, int width, int height);
    void (*w_avg)(uint8_t *dst, ptrdiff_t dst_stride, const int16_t *src0, const int16_t *src1, int width, int height,
                  int denom, int w0, int w1, int o0, int o1);
    void (*put_ciip)(uint8_t *dst, ptrdiff_t dst_stride, int width, int height,
                     const uint8_t *inter, ptrdiff_t inter_stride, int inter_weight);
    void (*put_gpm)(uint8_t *dst, ptrdiff_t dst_stride, int width, int height, const int16_t *src0, const int16_t *src1,
                    const uint8_t *weights, int step_x, int step_y);
    void (*fetch_samples)(int16_t *dst, const uint8_t *src, ptrdiff_t src_stride, int x_frac, int y_frac);
    void (*bdof_fetch_samples)(int16_t *dst, const uint8_t *src, ptrdiff_t src_stride, int x_frac, int y_frac,
                               int width, int height);
    void (*prof_grad_filter)(int16_t *gradient_h, int16_t *gradient_v, ptrdiff_t gradient_stride,
                             const int16_t *src, ptrdiff_t src_stride, int width, int height, int pad);
    void (*apply_prof)(int16_t *dst, const int16_t *src, const int16_t *diff_mv_x, const int16_t *diff_mv_y);
    void (*apply_prof_uni)(uint8_t *dst, ptrdiff_t dst_stride, const int16_t *src,
                           const int16_t *diff_mv_x, const int16_t *diff_mv_y);
    void (*apply_prof_uni_w)(uint8_t *dst, ptrdiff_t dst_stride, const int16_t *src,
                             const int16_t *diff_mv_x, const int16_t *diff_mv_y, int denom, int wx, int ox);
    void (*apply_bdof)(uint8_t *dst, ptrdiff_t dst_stride, int16_t *src0, int16_t *src1, int block_w, int block_h);
    int  (*sad)(const int16_t *src0, const int16_t *src1, int dx, int dy, int block_w, int block_h);
    vvct_dmvr_fn      dmvr[2][2];
} VVCInterDSPContext;

/* intra (vvcdsp.h:97-111) */
typedef void (*vvct_pred_angular_fn)(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride,
                                     int c_idx, int mode, int ref_idx, int filter_flag, int need_pdpc);
typedef struct VVCIntraDSPContext {
    void (*intra_cclm_pred)(const struct VVCLocalContext *lc, int x0, int y0, int w, int h);
    void (*lmcs_scale_chroma)(struct VVCLocalContext *lc, int *dst, const int *coeff, int w, int h, int x0_cu, int y0_cu);
    void (*intra_pred)(const struct VVCLocalContext *lc, int x0, int y0, int w, int h, int c_idx);
    void (*pred_planar)(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride);
    void (*pred_mip)(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride,
                     int mode_id, int is_transpose);
    void (*pred_dc)(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride);
    void (*pred_v)(uint8_t *src, const uint8_t *top, int w, int h, ptrdiff_t stride);
    void (*pred_h)(uint8_t *src, const uint8_t *left, int w, int h, ptrdiff_t stride);
    vvct_pred_angular_fn pred_angular_v;
    vvct_pred_angular_fn pred_angular_h;
} VVCIntraDSPContext;

/* itx (vvcdsp.h:113-121) */
typedef void (*vvct_itx_fn)(int *coeffs, size_t nzw, size_t nzh, intptr_t log2_transform_range, intptr_t bit_depth);
typedef struct VVCItxDSPContext {
    void (*add_residual)(uint8_t *dst, const int *res, int width, int height, ptrdiff_t stride);
    void (*add_residual_joint)(uint8_t *dst, const int *res, int width, int height, ptrdiff_t stride, int c_sign, int shift);
    void (*pred_residual_joint)(int *buf, int width, int height, int c_sign, int shift);
    vvct_itx_fn itx[VVCT_N_TX_TYPE][VVCT_N_TX_TYPE][VVCT_N_TX_SIZE][VVCT_N_TX_SIZE];   /* [trh][trv][log2 w][log2 h] */
    void (*transform_bdpcm)(int *coeffs, int width, int height, int vertical, int log2_transform_range);
} VVCItxDSPContext;

/* lmcs, lf, sao, alf (vvcdsp.h:123-158) */
typedef struct VVCLMCSDSPContext {
    void (*filter)(uint8_t *dst, ptrdiff_t dst_stride, int width, int height, const uint8_t *lut);
} VVCLMCSDSPContext;

typedef void (*vvct_lf_fn)(uint8_t *pix, ptrdiff_t stride, const int32_t *beta, const int32_t *tc,
                           const uint8_t *no_p, const uint8_t *no_q, const uint8_t *max_len_p, const uint8_t *max_len_q, int arg);
typedef struct VVCLFDSPContext {
    int (*ladf_level[2])(const uint8_t *pix, ptrdiff_t stride);       /* [h, v] */
    vvct_lf_fn filter_luma[2];                                        /* arg = hor_ctu_edge */
    vvct_lf_fn filter_chroma[2];                                      /* arg = shift */
} VVCLFDSPContext;

typedef struct VVCSAODSPContext {
    void (*band_filter[9])(uint8_t *dst, const uint8_t *src, ptrdiff_t dst_stride, ptrdiff_t src_stride,
                           const int16_t *sao_offset_val, int sao_left_class, int width, int height);
    void (*edge_filter[9])(uint8_t *dst, const uint8_t *src, ptrdiff_t dst_stride,
                           const int16_t *sao_offset_val, int sao_eo_class, int width, int height);
    void (*edge_restore[2])(uint8_t *dst, const uint8_t *src, ptrdiff_t dst_stride, ptrdiff_t src_stride,
                            const struct SAOParams *sao, const int *borders, int width, int height, int c_idx,
                            const uint8_t *vert_edge, const uint8_t *horiz_edge, const uint8_t *diag_edge);
} VVCSAODSPContext;

typedef struct VVCALFDSPContext {
    void (*filter[2])(uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *src, ptrdiff_t src_stride,
                      int width, int height, const int16_t *filter, const int16_t *clip, int vb_pos);   /* [luma, chroma] */
    void (*filter_cc)(uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *luma, ptrdiff_t luma_stride,
                      int width, int height, int hs, int vs, const int16_t *filter, int vb_pos);
    void (*classify)(int *class_idx, int *transpose_idx, const uint8_t *src, ptrdiff_t src_stride, int width, int height,
                     int vb_pos, int *gradient_tmp);
    void (*recon_coeff_and_clip)(int16_t *coeff, int16_t *clip, const int *class_idx, const int *transpose_idx, int size,
                                 const int16_t *coeff_set, const uint8_t *clip_idx_set, const uint8_t *class_to_filt);
} VVCALFDSPContext;

typedef struct VVCDSPContext {
    VVCInterDSPContext inter;
    VVCIntraDSPContext intra;
    VVCItxDSPContext   itx;
    VVCLMCSDSPContext  lmcs;
    VVCLFDSPContext    lf;
    VVCSAODSPContext   sao;
    VVCALFDSPContext   alf;
} VVCDSPContext;

#endif /* AVCODEC_VVC_VVCDSP_H */

/* Overrides the entries listed above when bit_depth == 10; any other depth leaves the table untouched. */
void ff_vvc_dsp_init_cuda(VVCDSPContext *c, int bit_depth);

/* 0, or the first error (VVC_CUDA_ERR_*) any overridden entry met since the process started / the last reset. */
int  ff_vvc_dsp_cuda_last_error(void);
const char *ff_vvc_dsp_cuda_error_string(void);
void ff_vvc_dsp_cuda_reset_error(void);
size_t ff_vvc_dsp_cuda_sizeof_table(void);      /* sizeof(VVCDSPContext) as compiled into the library */

#ifdef __cplusplus
}
#endif
#endif /* VVCDSP_TABLE_H */
