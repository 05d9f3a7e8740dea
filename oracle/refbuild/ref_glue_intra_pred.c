/*
 * TEST INFRASTRUCTURE - the reference's own intra.intra_pred (with prepare_intra_edge_params and ref_filter,
 * libavcodec/vvc/vvc_intra_template.c:450-683) and intra.intra_cclm_pred (:29-375) called through the reference's table,
 * one call per VVCCudaIntraBlk record, with the smallest decoder contexts that carry what they read: the SPS / PPS
 * sizes, fc->frame, the MIP side tables fc->tab.imf / imtf / imm, the coding unit's prediction fields, lc->na and the
 * availability state that ff_vvc_get_left_available / _top_available (vvc_intra.c:591-648, the reference's own functions
 * in libvvcref.so) look at.  A record states the availability as sample counts; here it is produced in the reference's
 * own terms: at CTB borders through ctb_left_flag / ctb_up_flag / end_of_tiles_x, inside a CTB through the list of
 * reconstructed areas.
 */
#include <stdlib.h>
#include <string.h>
#include "libavcodec/vvc/vvcdec.h"
#include "libavcodec/vvc/vvc_ctu.h"
#include "libavutil/frame.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

static _Thread_local VVCFrameContext *fc;
static _Thread_local VVCLocalContext *lc;
static _Thread_local CodingUnit *cu;
static _Thread_local VVCSPS *sps;
static _Thread_local VVCPPS *pps;
static _Thread_local H266RawSPS *rsps;
static _Thread_local AVFrame *frame;
static _Thread_local uint8_t *tab[3];
static _Thread_local size_t tab_size;

static void contexts(const VVCCudaFrame *f)
{
    if (!fc) {
        fc = calloc(1, sizeof(*fc));   lc = calloc(1, sizeof(*lc));     cu = calloc(1, sizeof(*cu));
        sps = calloc(1, sizeof(*sps)); pps = calloc(1, sizeof(*pps));   rsps = calloc(1, sizeof(*rsps));
        frame = calloc(1, sizeof(*frame));
    }
    sps->r = rsps;
    sps->bit_depth = f->bit_depth; sps->pixel_shift = 1;
    sps->ctb_log2_size_y = f->ctb_log2; sps->ctb_size_y = 1 << f->ctb_log2;
    sps->hshift[0] = sps->vshift[0] = 0;
    sps->hshift[1] = sps->hshift[2] = f->hshift; sps->vshift[1] = sps->vshift[2] = f->vshift;
    sps->min_cb_log2_size_y = 2;
    sps->width = f->width; sps->height = f->height;
    pps->width = f->width; pps->height = f->height;
    pps->min_cb_width = (f->width + 3) >> 2; pps->min_cb_height = (f->height + 3) >> 2;
    const size_t need = (size_t)pps->min_cb_width * pps->min_cb_height;
    if (need > tab_size) {
        for (int i = 0; i < 3; i++) { free(tab[i]); tab[i] = calloc(need, 1); }
        tab_size = need;
    }
    fc->tab.imf = tab[0]; fc->tab.imtf = tab[1]; fc->tab.imm = tab[2];
    fc->ps.sps = sps; fc->ps.pps = pps; fc->frame = frame;
    fc->vvcdsp = *vvcref_dsp(f->bit_depth);            /* intra_pred reaches its leaf predictors through the frame context's table */
    lc->fc = fc; lc->cu = cu;
}

static void area(int ch, int x, int y, int w, int h)
{
    lc->ras[ch][lc->num_ras[ch]++] = (ReconstructedArea){ .x = x, .y = y, .w = w, .h = h };
}

/* the state that makes the reference's availability functions answer min(target, n_top) / min(target, n_left) for the
 * block at (x, y) of plane type ch (0 luma, 1 chroma; plane units) */
static void availability(const VVCCudaFrame *f, int ch, int x, int y, int n_left, int n_top)
{
    const int hs = ch ? f->hshift : 0, vs = ch ? f->vshift : 0;
    const int at_ctb_left = !(x & ((1 << (f->ctb_log2 - hs)) - 1)), at_ctb_top = !(y & ((1 << (f->ctb_log2 - vs)) - 1));
    if (at_ctb_top) {
        lc->ctb_up_flag = n_top > 0;
        if (n_top > 0 && ((x + n_top) << hs) > lc->end_of_tiles_x)
            lc->end_of_tiles_x = (x + n_top) << hs;
    } else if (n_top > 0) {
        area(ch, x, y - 1, n_top, 1);
    }
    if (at_ctb_left)
        lc->ctb_left_flag = n_left > 0;
    else if (n_left > 0)
        area(ch, x - 1, y, 1, n_left);
}

void vvcref_intra_pred_frame(const VVCCudaFrame *f, const VVCCudaIntraBlk *blks, int n)
{
    const VVCDSPContext *dsp = vvcref_dsp(f->bit_depth);
    contexts(f);
    for (int i = 0; i < n; i++) {
        const VVCCudaIntraBlk *b = &blks[i];
        const int c = b->c_idx, hs = c ? f->hshift : 0, vs = c ? f->vshift : 0;
        const int x0 = b->x0 << hs, y0 = b->y0 << vs;              /* luma position */
        for (int p = 0; p < 3; p++) {
            frame->data[p] = (uint8_t *)f->data[p] + b->pic * f->batch_stride[p];
            frame->linesize[p] = (int)f->stride[p];
        }
        memset(cu, 0, sizeof(*cu));
        cu->x0 = x0; cu->y0 = y0;
        cu->cb_width = b->cb_w; cu->cb_height = b->cb_h;
        cu->isp_split_type = (b->flags & VVC_CUDA_INTRA_F_ISP) ? ISP_HOR_SPLIT : ISP_NO_SPLIT;
        cu->intra_luma_ref_idx = b->ref_idx;
        cu->bdpcm_flag[c] = !!(b->flags & VVC_CUDA_INTRA_F_BDPCM);
        if (c) cu->intra_pred_mode_c = b->pred_mode; else cu->intra_pred_mode_y = b->pred_mode;
        lc->na.cand_up_left = !!(b->flags & VVC_CUDA_INTRA_F_UP_LEFT);
        lc->num_ras[0] = lc->num_ras[1] = 0;
        lc->ctb_up_flag = lc->ctb_left_flag = 0;
        lc->end_of_tiles_x = 0;
        rsps->sps_chroma_vertical_collocated_flag = !!(b->flags & VVC_CUDA_INTRA_F_COLLOCATED);
        if (b->kind == VVC_CUDA_INTRA_KIND_CCLM) {
            /* chroma counts for the T / L modes, one-sample luma availability for the rest */
            availability(f, 1, b->x0, b->y0, b->avail_left, b->avail_top);
            availability(f, 0, x0, y0, !!(b->flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_L), !!(b->flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_T));
            dsp->intra.intra_cclm_pred(lc, x0, y0, b->w << hs, b->h << vs);
            continue;
        }
        availability(f, c > 0, b->x0, b->y0, b->avail_left, b->avail_top);
        const size_t cb = (size_t)(y0 >> 2) * pps->min_cb_width + (x0 >> 2);
        if (b->kind == VVC_CUDA_INTRA_KIND_MIP) {
            fc->tab.imf[cb] = 1;
            fc->tab.imtf[cb] = !!(b->flags & VVC_CUDA_INTRA_F_MIP_TRANSP);
            fc->tab.imm[cb] = b->pred_mode;
            cu->mip_chroma_direct_flag = c > 0;
        }
        dsp->intra.intra_pred(lc, x0, y0, b->w << hs, b->h << vs, c);
        fc->tab.imf[cb] = fc->tab.imtf[cb] = fc->tab.imm[cb] = 0;
    }
}
