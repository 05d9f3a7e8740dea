/*
 * Oracle: LMCS luma mapping (TEST INFRASTRUCTURE, see vvc_oracle.h).
 * Restates lmcs_filter_luma (libavcodec/vvc/vvc_filter_template.c:25-36) as driven per CTU by
 * ff_vvc_lmcs_filter (libavcodec/vvc/vvc_filter.c:1322-1332) and per inter CU by predict_inter
 * (libavcodec/vvc/vvc_inter.c:888-891).
 */
#include "vvc_oracle.h"

static void map_rect(const OPlane *pl, int x0, int y0, int w, int h, const uint16_t *lut)
{
    for (int y = y0; y < y0 + h; y++)
        for (int x = x0; x < x0 + w; x++)
            pl->p[y * pl->pitch + x] = lut[pl->p[y * pl->pitch + x]];
}

void vvco_lmcs_frame(const VVCCudaFrame *f, const uint16_t *lut, const uint8_t *ctb_enable)
{
    const int ctb = 1 << f->ctb_log2, cols = o_ctb_cols(f), rows = o_ctb_rows(f);
    for (int k = 0; k < f->batch; k++) {
        const OPlane pl = o_plane(f, 0, k);
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++)
                if (!ctb_enable || ctb_enable[(size_t)k * cols * rows + cy * cols + cx])
                    map_rect(&pl, cx * ctb, cy * ctb, o_min(ctb, f->width - cx * ctb), o_min(ctb, f->height - cy * ctb), lut);
    }
}

void vvco_lmcs_rects(const VVCCudaFrame *f, const uint16_t *lut, const VVCCudaRect *rects, int n)
{
    for (int i = 0; i < n; i++) {
        const OPlane pl = o_plane(f, 0, rects[i].pic);
        map_rect(&pl, rects[i].x, rects[i].y, rects[i].w, rects[i].h, lut);
    }
}


/* lmcs_derive_chroma_scale (libavcodec/vvc/vvc_intra_template.c:377-428) per VPDU: average of the reconstructed luma left
 * of and above the VPDU (lmcs_sum_samples :377-387 replicates the last sample past the picture), bin search over the pivots */
void vvco_lmcs_chroma_scale(const VVCCudaFrame *f, const VVCCudaLmcsVpdu *vpdus, int n, const VVCCudaLmcsParams *lp, uint16_t *scales)
{
    const int ctb = 1 << f->ctb_log2, size = ctb < 64 ? ctb : 64, bd = f->bit_depth;
    for (int i = 0; i < n; i++) {
        const VVCCudaLmcsVpdu *v = &vpdus[i];
        const OPlane pl = o_plane(f, 0, v->pic);
        int luma = 0, cnt = 0, k;
        if (v->avail_l) {
            const int avail = f->height - v->y, m = avail < size ? avail : size;
            for (int j = 0; j < m; j++)
                luma += pl.p[(v->y + j) * pl.pitch + v->x - 1];
            luma += pl.p[(v->y + m - 1) * pl.pitch + v->x - 1] * (size - m);
            cnt = size;
        }
        if (v->avail_t) {
            const int avail = f->width - v->x, m = avail < size ? avail : size;
            for (int j = 0; j < m; j++)
                luma += pl.p[(v->y - 1) * pl.pitch + v->x + j];
            luma += pl.p[(v->y - 1) * pl.pitch + v->x + m - 1] * (size - m);
            cnt += size;
        }
        luma = cnt ? (luma + (cnt >> 1)) >> o_ilog2(cnt) : 1 << (bd - 1);
        for (k = lp->min_bin_idx; k <= lp->max_bin_idx; k++)
            if (luma < lp->pivot[k + 1])
                break;
        scales[i] = lp->chroma_scale_coeff[k < 15 ? k : 15];
    }
}
