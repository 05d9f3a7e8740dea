/*
 * TEST INFRASTRUCTURE - the reference's own deblocking drivers ff_vvc_deblock_vertical / ff_vvc_deblock_horizontal
 * (libavcodec/vvc/vvc_filter.c:861-1003, with vvc_deblock_bs :756-781 and everything it calls), run CTB by CTB over a
 * picture whose per-4x4 side tables (fc->tab.*) were filled from the list inputs of vvc_cuda_deblock_params_frame.
 * The reference filters the picture in place; the parity tests compare that picture with the one the derived parameter
 * maps produce, so strengths, lengths, beta and tc are pinned through their effect on every sample.
 */
#include <stdlib.h>
#include <string.h>
#include "libavcodec/vvc/vvcdec.h"
#include "libavcodec/vvc/vvc_ctu.h"
#include "libavcodec/vvc/vvc_filter.h"
#include "libavutil/frame.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

static _Thread_local const VVCFrameContext *qp_fc;

/* libavcodec/vvc/vvc_ctu.c:2532-2538 (the parser's translation unit is not part of this build) */
int ff_vvc_get_qPy(const VVCFrameContext *fc, const int xc, const int yc)
{
    const int l2 = fc->ps.sps->min_cb_log2_size_y;
    return fc->tab.qp[LUMA][(xc >> l2) + (yc >> l2) * fc->ps.pps->min_cb_width];
}

#define NEW(n, type) ((type *)calloc((n), sizeof(type)))

void vvcref_deblock_params_filter(const VVCCudaFrame *f, const VVCCudaDbkTU *tus, int n_tus, const VVCCudaDbkMvf *mvfs, int n_mvfs,
                                  const VVCCudaDbkCtb *ctbs, const VVCCudaDbkParams *prm)
{
    VVCFrameContext *fc = NEW(1, VVCFrameContext);
    VVCLocalContext *lc = NEW(1, VVCLocalContext);
    SliceContext *sc = NEW(1, SliceContext);
    VVCSPS *sps = NEW(1, VVCSPS);
    VVCPPS *pps = NEW(1, VVCPPS);
    H266RawSPS *rsps = NEW(1, H266RawSPS);
    H266RawPPS *rpps = NEW(1, H266RawPPS);
    AVFrame *frame = NEW(1, AVFrame);
    RefPicList *rpl = NEW(2, RefPicList);
    const int uw = (f->width + 3) >> 2, uh = (f->height + 3) >> 2, ctb = 1 << f->ctb_log2;
    const int cw = (f->width + ctb - 1) >> f->ctb_log2, ch = (f->height + ctb - 1) >> f->ctb_log2;
    const size_t nu = (size_t)uw * uh;
    (void)qp_fc;

    sps->r = rsps; pps->r = rpps;
    sps->bit_depth = f->bit_depth; sps->pixel_shift = 1;
    sps->ctb_log2_size_y = f->ctb_log2; sps->ctb_size_y = ctb;
    sps->hshift[1] = sps->hshift[2] = f->hshift; sps->vshift[1] = sps->vshift[2] = f->vshift;
    sps->min_cb_log2_size_y = 2;
    sps->width = f->width; sps->height = f->height;
    sps->qp_bd_offset = prm->qp_bd_offset;
    rsps->sps_chroma_format_idc = f->chroma_format_idc;
    rsps->sps_ladf_enabled_flag = prm->ladf_enabled;
    rsps->sps_ladf_lowest_interval_qp_offset = prm->ladf_lowest_interval_qp_offset;
    sps->num_ladf_intervals = prm->num_ladf_intervals;
    for (int i = 0; i < 4; i++) rsps->sps_ladf_qp_offset[i] = prm->ladf_qp_offset[i];
    for (int i = 0; i < 5; i++) sps->ladf_interval_lower_bound[i] = prm->ladf_interval_lower_bound[i];
    pps->width = f->width; pps->height = f->height;
    pps->min_cb_width = pps->min_tu_width = pps->min_pu_width = uw;
    pps->ctb_width = cw;
    rpps->pps_loop_filter_across_slices_enabled_flag = 1;
    rpps->pps_loop_filter_across_tiles_enabled_flag = 0;        /* a record's no_left / no_top = a tile border */
    for (int l = 0; l < 2; l++)
        for (int i = 0; i < VVC_MAX_REF_ENTRIES; i++)
            rpl[l].list[i] = i;                                 /* the records name pictures by number */
    sc->rpl = rpl;
    fc->ps.sps = sps; fc->ps.pps = pps; fc->frame = frame;
    fc->vvcdsp = *vvcref_dsp(f->bit_depth);
    lc->fc = fc; lc->sc = sc;

    fc->tab.mvf = NEW(nu, MvField);
    for (int c = 0; c < 3; c++) {
        fc->tab.tu_coded_flag[c] = NEW(nu, uint8_t); fc->tab.qp[c] = NEW(nu, int8_t);
        fc->tab.vertical_bs[c] = NEW(nu, uint8_t);   fc->tab.horizontal_bs[c] = NEW(nu, uint8_t);
    }
    for (int t = 0; t < 2; t++) {
        fc->tab.tb_pos_x0[t] = NEW(nu, int); fc->tab.tb_pos_y0[t] = NEW(nu, int);
        fc->tab.tb_width[t] = NEW(nu, uint8_t); fc->tab.tb_height[t] = NEW(nu, uint8_t);
        fc->tab.pcmf[t] = NEW(nu, uint8_t);
    }
    fc->tab.tu_joint_cbcr_residual_flag = NEW(nu, uint8_t);
    fc->tab.cb_pos_x[0] = NEW(nu, int); fc->tab.cb_pos_y[0] = NEW(nu, int);
    fc->tab.cb_width[0] = NEW(nu, uint8_t); fc->tab.cb_height[0] = NEW(nu, uint8_t);
    fc->tab.msf = NEW(nu, uint8_t); fc->tab.iaf = NEW(nu, uint8_t);
    fc->tab.vertical_p = NEW(nu, uint8_t); fc->tab.vertical_q = NEW(nu, uint8_t);
    fc->tab.horizontal_p = NEW(nu, uint8_t); fc->tab.horizontal_q = NEW(nu, uint8_t);
    fc->tab.deblock = NEW((size_t)cw * ch, DBParams);

    for (int k = 0; k < f->batch; k++) {
        for (int p = 0; p < 3; p++) {
            frame->data[p] = (uint8_t *)f->data[p] + k * f->batch_stride[p];
            frame->linesize[p] = (int)f->stride[p];
        }
        for (size_t i = 0; i < nu; i++) {
            fc->tab.tb_pos_x0[0][i] = fc->tab.tb_pos_x0[1][i] = fc->tab.tb_pos_y0[0][i] = fc->tab.tb_pos_y0[1][i] = -1;
            for (int c = 0; c < 3; c++)
                fc->tab.vertical_bs[c][i] = fc->tab.horizontal_bs[c][i] = 0;
        }
        for (int i = 0; i < n_tus; i++) {
            const VVCCudaDbkTU *tu = &tus[i];
            if (tu->pic != k) continue;
            const int w = 1 << tu->log2_w, h = 1 << tu->log2_h;
            for (int y = tu->y0; y < tu->y0 + h && y < f->height; y += 4)
                for (int x = tu->x0; x < tu->x0 + w && x < f->width; x += 4) {
                    const size_t u = (size_t)(y >> 2) * uw + (x >> 2);
                    if (tu->planes & VVC_CUDA_DBK_TU_LUMA) {
                        fc->tab.tb_pos_x0[0][u] = tu->x0; fc->tab.tb_pos_y0[0][u] = tu->y0;
                        fc->tab.tb_width[0][u] = w; fc->tab.tb_height[0][u] = h;
                        fc->tab.tu_coded_flag[0][u] = !!(tu->flags & VVC_CUDA_DBK_CBF_Y);
                        fc->tab.pcmf[0][u] = !!(tu->flags & VVC_CUDA_DBK_BDPCM_Y);
                        fc->tab.qp[0][u] = tu->qp[0];
                        fc->tab.cb_pos_x[0][u] = tu->x0 - 4 * tu->cu_dx; fc->tab.cb_pos_y[0][u] = tu->y0 - 4 * tu->cu_dy;
                        fc->tab.cb_width[0][u] = 1 << tu->cb_log2_w; fc->tab.cb_height[0][u] = 1 << tu->cb_log2_h;
                        fc->tab.msf[u] = !!(tu->cu_flags & VVC_CUDA_DBK_CU_SUBBLOCK); fc->tab.iaf[u] = 0;
                    }
                    if (tu->planes & VVC_CUDA_DBK_TU_CHROMA) {
                        fc->tab.tb_pos_x0[1][u] = tu->x0; fc->tab.tb_pos_y0[1][u] = tu->y0;
                        fc->tab.tb_width[1][u] = w >> f->hshift; fc->tab.tb_height[1][u] = h >> f->vshift;
                        fc->tab.tu_coded_flag[1][u] = !!(tu->flags & VVC_CUDA_DBK_CBF_CB);
                        fc->tab.tu_coded_flag[2][u] = !!(tu->flags & VVC_CUDA_DBK_CBF_CR);
                        fc->tab.tu_joint_cbcr_residual_flag[u] = !!(tu->flags & VVC_CUDA_DBK_JOINT);
                        fc->tab.pcmf[1][u] = !!(tu->flags & VVC_CUDA_DBK_BDPCM_C);
                        fc->tab.qp[1][u] = tu->qp[1]; fc->tab.qp[2][u] = tu->qp[2];
                    }
                }
        }
        for (int i = 0; i < n_mvfs; i++) {
            const VVCCudaDbkMvf *m = &mvfs[i];
            if (m->pic != k) continue;
            for (int y = m->y0; y < m->y0 + 4 * m->h4 && y < f->height; y += 4)
                for (int x = m->x0; x < m->x0 + 4 * m->w4 && x < f->width; x += 4) {
                    MvField *mv = &fc->tab.mvf[(size_t)(y >> 2) * uw + (x >> 2)];
                    memset(mv, 0, sizeof(*mv));
                    mv->pred_flag = m->pred_flag; mv->ciip_flag = m->ciip_flag;
                    for (int l = 0; l < 2; l++) {
                        mv->ref_idx[l] = (int8_t)m->ref_pic[l];
                        mv->mv[l].x = m->mv[l][0]; mv->mv[l].y = m->mv[l][1];
                    }
                }
        }
        const VVCCudaDbkCtb *cb = ctbs + (size_t)k * cw * ch;
        for (int i = 0; i < cw * ch; i++)
            for (int c = 0; c < 3; c++) {
                fc->tab.deblock[i].beta_offset[c] = cb[i].beta_offset[c];
                fc->tab.deblock[i].tc_offset[c] = cb[i].tc_offset[c];
            }
        for (int dir = 1; dir >= 0; dir--)
            for (int ry = 0; ry < ch; ry++)
                for (int rx = 0; rx < cw; rx++) {
                    const VVCCudaDbkCtb *c = &cb[ry * cw + rx];
                    lc->boundary_flags = (c->no_left ? BOUNDARY_LEFT_TILE : 0) | (c->no_top ? BOUNDARY_UPPER_TILE : 0);
                    if (dir) ff_vvc_deblock_vertical(lc, rx << f->ctb_log2, ry << f->ctb_log2);
                    else     ff_vvc_deblock_horizontal(lc, rx << f->ctb_log2, ry << f->ctb_log2);
                }
    }
    /* test infrastructure: the tables are released with the process */
}
