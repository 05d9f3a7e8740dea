/*
 * TEST INFRASTRUCTURE - the reference's own intra.lmcs_scale_chroma (lmcs_scale_chroma + lmcs_derive_chroma_scale +
 * lmcs_sum_samples, libavcodec/vvc/vvc_intra_template.c:377-448) called through the reference's table with the smallest
 * decoder contexts that carry what it reads: fc->ps.lmcs (pivots, chroma_scale_coeff, bin range), the SPS / PPS sizes,
 * the luma plane of fc->frame, and the availability state ff_vvc_get_top_available / _left_available
 * (vvc_intra.c:591-648, the reference's own functions in libvvcref.so) look at.
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

static void contexts(int bit_depth, int ctb_log2, int width, int height)
{
    if (!fc) {
        fc = calloc(1, sizeof(*fc));   lc = calloc(1, sizeof(*lc));     cu = calloc(1, sizeof(*cu));
        sps = calloc(1, sizeof(*sps)); pps = calloc(1, sizeof(*pps));   rsps = calloc(1, sizeof(*rsps));
        frame = calloc(1, sizeof(*frame));
    }
    sps->r = rsps;
    sps->bit_depth = bit_depth; sps->pixel_shift = 1;
    sps->ctb_log2_size_y = ctb_log2; sps->ctb_size_y = 1 << ctb_log2;
    sps->hshift[0] = sps->vshift[0] = 0; sps->hshift[1] = sps->vshift[1] = sps->hshift[2] = sps->vshift[2] = 1;
    sps->width = width; sps->height = height;
    pps->width = width; pps->height = height;
    fc->ps.sps = sps; fc->ps.pps = pps; fc->frame = frame;
    lc->fc = fc; lc->cu = cu;
}

/* The scaling arithmetic alone: the VPDU cache of lc is primed with `scale`, so lmcs_derive_chroma_scale returns it. */
void vvcref_lmcs_scale_block(int *dst, const int *coeff, int w, int h, int scale, int bit_depth)
{
    contexts(bit_depth, 7, 4096, 4096);
    lc->lmcs.x_vpdu = 0; lc->lmcs.y_vpdu = 0; lc->lmcs.chroma_scale = scale;
    vvcref_dsp(bit_depth)->intra.lmcs_scale_chroma(lc, dst, coeff, w, h, 5, 9);     /* a CU inside the VPDU at (0, 0) */
}

/* The derivation: one scale per VPDU record, from the luma plane of `f`.  The availability the records state is
 * produced in the reference's own terms: CTB-border positions through ctb_left_flag / ctb_up_flag, positions inside a
 * CTB through the list of reconstructed areas. */
void vvcref_lmcs_chroma_scale(const VVCCudaFrame *f, const VVCCudaLmcsVpdu *vpdus, int n, const VVCCudaLmcsParams *lp, uint16_t *scales)
{
    int one[1] = { 1 }, out[1];
    contexts(f->bit_depth, f->ctb_log2, f->width, f->height);
    fc->ps.lmcs.min_bin_idx = lp->min_bin_idx; fc->ps.lmcs.max_bin_idx = lp->max_bin_idx;
    memcpy(fc->ps.lmcs.pivot, lp->pivot, sizeof(lp->pivot));
    memcpy(fc->ps.lmcs.chroma_scale_coeff, lp->chroma_scale_coeff, sizeof(lp->chroma_scale_coeff));
    for (int i = 0; i < n; i++) {
        const VVCCudaLmcsVpdu *v = &vpdus[i];
        const int ctb = 1 << f->ctb_log2;
        frame->data[0] = (uint8_t *)f->data[0] + v->pic * f->batch_stride[0];
        frame->linesize[0] = (int)f->stride[0];
        cu->x0 = v->x + 4; cu->y0 = v->y + 4;
        lc->ctb_left_flag = v->avail_l; lc->ctb_up_flag = v->avail_t;
        lc->end_of_tiles_x = f->width;
        lc->num_ras[0] = 0;
        if (v->avail_t && (v->y & (ctb - 1)))
            lc->ras[0][lc->num_ras[0]++] = (ReconstructedArea){ .x = 0, .y = 0, .w = f->width, .h = v->y };
        if (v->avail_l && (v->x & (ctb - 1)))
            lc->ras[0][lc->num_ras[0]++] = (ReconstructedArea){ .x = 0, .y = v->y, .w = v->x, .h = f->height - v->y };
        lc->lmcs.x_vpdu = -1; lc->lmcs.y_vpdu = -1;
        vvcref_dsp(f->bit_depth)->intra.lmcs_scale_chroma(lc, out, one, 1, 1, cu->x0, cu->y0);
        scales[i] = (uint16_t)lc->lmcs.chroma_scale;
    }
}
