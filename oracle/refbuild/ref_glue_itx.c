/*
 * TEST INFRASTRUCTURE - residual stage through the UNMODIFIED reference entries:
 *   itx.transform_bdpcm, ff_vvc_inv_lfnst_1d, itx.itx[trh][trv][log2w][log2h],
 *   itx.add_residual, itx.add_residual_joint
 * in the order of itransform() (libavcodec/vvc/vvc_intra.c:432-478).  The LFNST gather/scatter
 * around ff_vvc_inv_lfnst_1d follows ilfnst_transform (:65-127), which is static and needs decoder
 * contexts, so its 20 lines of index shuffling are re-expressed here; the arithmetic is the
 * reference's.
 */
#include <stdint.h>
#include <string.h>
#include "libavcodec/vvc/vvcdsp.h"
#include "libavcodec/vvc/vvc_data.h"
#include "libavcodec/vvc/vvc_itx_1d.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

static int pred_mode_for_set(int set)
{
    if (set == 1)
        return -1;                                 /* wide angle -> set 1 (vvc_itx_1d.c:711) */
    for (int m = 0; m < 95; m++)
        if (ff_vvc_lfnst_tr_set_index[m] == set)
            return m;
    return 0;
}

static void ref_lfnst(int *c, int w, int h, int lfnst, int range)
{
    const int idx = lfnst & 3, set = (lfnst >> 2) & 3, transpose = (lfnst >> 4) & 1;
    const int n_in = (lfnst >> 5) & 1 ? 8 : 16;
    const int big = w >= 8 && h >= 8, n_out = big ? 48 : 16, side = big ? 8 : 4;
    int u[16], v[48];
    for (int x = 0; x < n_in; x++)
        u[x] = c[w * ff_vvc_diag_scan_y[2][2][x] + ff_vvc_diag_scan_x[2][2][x]];
    ff_vvc_inv_lfnst_1d(v, u, n_in, n_out, pred_mode_for_set(set), idx, range);
    {
        const int *src = v;
        for (int y = 0; y < side; y++) {
            const int len = y < 4 ? side : 4;
            for (int x = 0; x < len; x++) {
                if (transpose) c[x * w + y] = src[x];
                else           c[y * w + x] = src[x];
            }
            src += len;
        }
    }
}

void vvcref_lmcs_scale_block(int *dst, const int *coeff, int w, int h, int scale, int bit_depth);      /* ref_glue_lmcs_chroma.c */

void vvcref_itx_frame(const VVCCudaFrame *f, int32_t *coeffs, const VVCCudaTB *tbs, int n_tbs, int range)
{
    const VVCDSPContext *dsp = vvcref_dsp(f->bit_depth);
    static _Thread_local int tmp[64 * 64];
    for (int i = 0; i < n_tbs; i++) {
        const VVCCudaTB *tb = &tbs[i];
        const int w = 1 << tb->log2_w, h = 1 << tb->log2_h;
        int nzw = tb->nzw, nzh = tb->nzh;
        int *src = coeffs + tb->coeff_offset;
        int *c = (tb->flags & VVC_CUDA_TB_STORE_RESIDUAL) ? src : tmp;
        if (c != src)
            memcpy(c, src, sizeof(int) * w * h);
        if (tb->flags & (VVC_CUDA_TB_BDPCM | VVC_CUDA_TB_BDPCM_VERT))
            dsp->itx.transform_bdpcm(c, w, h, !!(tb->flags & VVC_CUDA_TB_BDPCM_VERT), range);
        if (!(tb->flags & VVC_CUDA_TB_TS)) {
            if (tb->lfnst) {
                ref_lfnst(c, w, h, tb->lfnst, range);
                nzw = nzh = (w >= 8 && h >= 8) ? 8 : 4;
            }
            dsp->itx.itx[tb->trh][tb->trv][tb->log2_w][tb->log2_h](c, nzw, nzh, range, f->bit_depth);
        }
        if (tb->flags & VVC_CUDA_TB_STORE_RESIDUAL)
            continue;
        if (tb->chroma_scale) {
            /* itransform() / add_residual_for_joint_coding_chroma with chroma_scale (vvc_intra.c:166-186, 468-475); the scale
             * is the literal one of the record (the derivation is pinned separately, ref_glue_lmcs_chroma.c) */
            static _Thread_local int temp[64 * 64];
            uint8_t *plane = (uint8_t *)f->data[tb->c_idx] + tb->pic * f->batch_stride[tb->c_idx];
            vvcref_lmcs_scale_block(temp, c, w, h, tb->chroma_scale, f->bit_depth);
            dsp->itx.add_residual(plane + tb->y0 * f->stride[tb->c_idx] + tb->x0 * 2, temp, w, h, f->stride[tb->c_idx]);
            if (tb->flags & VVC_CUDA_TB_JOINT) {
                const int jc = tb->joint_c_idx;
                uint8_t *pj = (uint8_t *)f->data[jc] + tb->pic * f->batch_stride[jc];
                dsp->itx.pred_residual_joint(c, w, h, tb->joint_sign, tb->joint_shift);
                vvcref_lmcs_scale_block(c, c, w, h, tb->chroma_scale, f->bit_depth);
                dsp->itx.add_residual(pj + tb->y0 * f->stride[jc] + tb->x0 * 2, c, w, h, f->stride[jc]);
            }
            continue;
        }
        {
            uint8_t *plane = (uint8_t *)f->data[tb->c_idx] + tb->pic * f->batch_stride[tb->c_idx];
            dsp->itx.add_residual(plane + tb->y0 * f->stride[tb->c_idx] + tb->x0 * 2, c, w, h, f->stride[tb->c_idx]);
        }
        if (tb->flags & VVC_CUDA_TB_JOINT) {
            const int jc = tb->joint_c_idx;
            uint8_t *plane = (uint8_t *)f->data[jc] + tb->pic * f->batch_stride[jc];
            dsp->itx.add_residual_joint(plane + tb->y0 * f->stride[jc] + tb->x0 * 2, c, w, h, f->stride[jc],
                                        tb->joint_sign, tb->joint_shift);
        }
    }
}
