/*
 * Oracle: sample adaptive offset for a whole picture (TEST INFRASTRUCTURE, see vvc_oracle.h).
 *
 * Restates per sample what the reference does per CTB:
 *   driver        ff_vvc_sao_filter     libavcodec/vvc/vvc_filter.c:154-298
 *   band offset   sao_band_filter       libavcodec/h26x/h2656_sao_template.c:24-46
 *   edge offset   sao_edge_filter       libavcodec/h26x/h2656_sao_template.c:50-79
 *   border rules  sao_edge_restore_0/1  libavcodec/h26x/h2656_sao_template.c:81-215
 *
 * The reference filters in place from a per-CTB copy whose 1-sample halo was saved before SAO
 * touched the neighbours (sao_pixel_buffer_h/v), so every neighbour is a pre-SAO (deblocked)
 * sample; here the source picture is read-only and the result goes to a second picture.
 */
#include "vvc_oracle.h"

static inline int cmp3(int a, int b) { return (a > b) - (a < b); }

static void sao_ctb_plane(const OPlane *d, const OPlane *s, int x0, int y0, int w, int h, int bd,
                          const VVCCudaSAOCtb *p, int c, const int borders[4])
{
    static const int8_t step[4][2][2] = {      /* [eo][a|b][dx,dy], h2656_sao_template.c:54-59 */
        { { -1, 0 }, { 1, 0 } }, { { 0, -1 }, { 0, 1 } }, { { -1, -1 }, { 1, 1 } }, { { 1, -1 }, { -1, 1 } } };
    static const uint8_t cat[5] = { 1, 2, 0, 3, 4 };
    const int16_t *off = p->offset_val[c];
    const int type = p->type_idx[c], eo = p->eo_class[c];

    if (type == 1) {
        int table[32] = { 0 };
        for (int k = 0; k < 4; k++)
            table[(k + p->band_position[c]) & 31] = off[k + 1];
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int v = s->p[(y0 + y) * s->pitch + x0 + x];
                d->p[(y0 + y) * d->pitch + x0 + x] = (pel)o_clip_pel(v + table[(v >> (bd - 5)) & 31], bd);
            }
        return;
    }
    if (type != 2) {
        for (int y = 0; y < h; y++)
            memcpy(&d->p[(y0 + y) * d->pitch + x0], &s->p[(y0 + y) * s->pitch + x0], w * sizeof(pel));
        return;
    }

    {
        const int not_v = eo != 1, not_h = eo != 0;
        const int bl = borders[0], bt = borders[1], br = borders[2], bb = borders[3];
        /* geometry of the restore pass (:148-177) */
        const int init_x = not_v && bl, w1 = w - (not_v && br);
        const int init_y = not_h && bt, h1 = h - (not_h && bb);
        const int ve0 = p->no_filter & 1, ve1 = (p->no_filter >> 1) & 1;
        const int he0 = (p->no_filter >> 2) & 1, he1 = (p->no_filter >> 3) & 1;
        const int dg0 = (p->no_filter >> 4) & 1, dg1 = (p->no_filter >> 5) & 1;
        const int dg2 = (p->no_filter >> 6) & 1, dg3 = (p->no_filter >> 7) & 1;
        const int keep_ul = !dg0 && eo == 2 && !bl && !bt, keep_ur = !dg1 && eo == 3 && !bt && !br;
        const int keep_lr = !dg2 && eo == 2 && !br && !bb, keep_ll = !dg3 && eo == 3 && !bl && !bb;

        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int px = x0 + x, py = y0 + y;
                const int v = s->p[py * s->pitch + px];
                int out;
                const int on_pic_border = (not_v && ((bl && x == 0) || (br && x == w - 1))) ||
                                          (not_h && ((bt && y == 0) || (bb && y == h - 1)));
                if (on_pic_border) {
                    out = o_clip_pel(v + off[0], bd);
                } else {
                    const int a = s->p[(py + step[eo][0][1]) * s->pitch + px + step[eo][0][0]];
                    const int b = s->p[(py + step[eo][1][1]) * s->pitch + px + step[eo][1][0]];
                    out = o_clip_pel(v + off[cat[2 + cmp3(v, a) + cmp3(v, b)]], bd);
                }
                if (p->restore) {
                    int keep = 0;
                    if (ve0 && not_v && x == 0      && y >= init_y + keep_ul && y < h1 - keep_ll) keep = 1;
                    if (ve1 && not_v && x == w1 - 1 && y >= init_y + keep_ur && y < h1 - keep_lr) keep = 1;
                    if (he0 && not_h && y == 0      && x >= init_x + keep_ul && x < w1 - keep_ur) keep = 1;
                    if (he1 && not_h && y == h1 - 1 && x >= init_x + keep_ll && x < w1 - keep_lr) keep = 1;
                    if (dg0 && eo == 2 && x == 0      && y == 0)      keep = 1;
                    if (dg1 && eo == 3 && x == w1 - 1 && y == 0)      keep = 1;
                    if (dg2 && eo == 2 && x == w1 - 1 && y == h1 - 1) keep = 1;
                    if (dg3 && eo == 3 && x == 0      && y == h1 - 1) keep = 1;
                    if (keep)
                        out = v;
                }
                d->p[py * d->pitch + px] = (pel)out;
            }
    }
}

void vvco_sao_frame(const VVCCudaFrame *dstf, const VVCCudaFrame *srcf, const VVCCudaSAOCtb *ctbs)
{
    const int ctb = 1 << srcf->ctb_log2, cols = o_ctb_cols(srcf), rows = o_ctb_rows(srcf);
    const int planes = srcf->chroma_format_idc ? 3 : 1;
    for (int k = 0; k < srcf->batch; k++)
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++) {
                const VVCCudaSAOCtb *p = &ctbs[(size_t)k * cols * rows + cy * cols + cx];
                const int borders[4] = { cx == 0, cy == 0, cx == cols - 1, cy == rows - 1 };
                for (int c = 0; c < planes; c++) {
                    const OPlane s = o_plane(srcf, c, k), d = o_plane(dstf, c, k);
                    const int hs = c ? srcf->hshift : 0, vs = c ? srcf->vshift : 0;
                    const int x0 = (cx * ctb) >> hs, y0 = (cy * ctb) >> vs;
                    const int w = o_min(ctb >> hs, s.w - x0), h = o_min(ctb >> vs, s.h - y0);
                    sao_ctb_plane(&d, &s, x0, y0, w, h, srcf->bit_depth, p, c, borders);
                }
            }
}
