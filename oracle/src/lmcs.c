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
