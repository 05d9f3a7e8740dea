/*
 * Oracle: adaptive loop filter for a whole picture (TEST INFRASTRUCTURE, see vvc_oracle.h).
 *
 * Restates, per sample, what the reference computes CTB by CTB:
 *   driver          ff_vvc_alf_filter            libavcodec/vvc/vvc_filter.c:1254-1319
 *   halo rule       alf_prepare_buffer           libavcodec/vvc/vvc_filter.c:1105-1137
 *   classification  alf_classify / alf_get_idx   libavcodec/vvc/vvc_filter_template.c:270-381
 *   coefficients    alf_recon_coeff_and_clip     libavcodec/vvc/vvc_filter_template.c:383-408
 *                   alf_get_coeff_and_clip       libavcodec/vvc/vvc_filter.c:1142-1170
 *   luma 7x7        alf_filter_luma              libavcodec/vvc/vvc_filter_template.c:43-135
 *   chroma 5x5      alf_filter_chroma            libavcodec/vvc/vvc_filter_template.c:137-221
 *   CC-ALF          alf_filter_cc                libavcodec/vvc/vvc_filter_template.c:223-263
 *
 * The reference filters in place out of a padded per-CTB copy whose halo comes from saved
 * pre-ALF lines; every tap therefore sees pre-ALF samples, and a halo that would cross a
 * flagged CTB edge repeats the CTB's own border sample.  Here that is a coordinate clamp on
 * a read-only source picture.
 */
#include "vvc_oracle.h"
#include "vvc_tables_c.h"

typedef struct Win {            /* source plane + the clamp window of the current CTB */
    const pel *p;
    ptrdiff_t pitch;
    int x_lo, x_hi, y_lo, y_hi; /* inclusive clamp limits in plane coordinates */
} Win;

static inline int tap(const Win *w, int x, int y)
{
    x = o_clip3(x, w->x_lo, w->x_hi);
    y = o_clip3(y, w->y_lo, w->y_hi);
    return w->p[y * w->pitch + x];
}

/* Clamp window of a CTB in plane c: a side is clamped at the CTB's own border when flagged,
 * otherwise left open (picture limits are always flagged by the caller). */
static Win make_win(const OPlane *pl, int x0, int y0, int w, int h, unsigned edges)
{
    Win win;
    win.p = pl->p; win.pitch = pl->pitch;
    win.x_lo = (edges & VVC_CUDA_EDGE_LEFT)   ? x0         : 0;
    win.x_hi = (edges & VVC_CUDA_EDGE_RIGHT)  ? x0 + w - 1 : pl->w - 1;
    win.y_lo = (edges & VVC_CUDA_EDGE_TOP)    ? y0         : 0;
    win.y_hi = (edges & VVC_CUDA_EDGE_BOTTOM) ? y0 + h - 1 : pl->h - 1;
    return win;
}

/* Row folding next to the virtual boundary: a vertical reach of k rows shrinks to what
 * fits on this side of the boundary (vvc_filter_template.c:80-96).  t = row - vb_pos. */
static inline int vb_reach(int k, int t, int span)
{
    if (t < 0 && t >= -span)
        return o_min(k, -t - 1);
    if (t >= 0 && t < span)
        return o_min(k, t);
    return k;
}

static inline int pair_clip(int cur, int a, int b, int c)
{
    return (int16_t)(o_clip3(a - cur, -c, c) + o_clip3(b - cur, -c, c));
}

/* ---- classification of one 4x4 block, (bx,by) relative to the CTB origin (x0,y0) ---- */
static void classify_block(const Win *w, int x0, int y0, int bx, int by, int vb, int bd,
                           int *cls, int *tr)
{
    static const uint8_t act_lut[16] = { 0, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4 };
    int first = 0, last = 4, scale = 2;
    int sv = 0, sh = 0, sd0 = 0, sd1 = 0;

    if (by + 4 == vb)      { last = 3;  scale = 3; }
    else if (by == vb)     { first = 1; scale = 3; }

    for (int i = first; i < last; i++) {
        const int yy = by + 2 * i;               /* loop variable y of the reference (:315) */
        int r0 = yy - 3, r1 = yy - 2, r2 = yy - 1, r3 = yy;
        if (yy == vb)          r3 = r2;          /* :321-322 */
        else if (yy == vb + 2) r0 = r1;          /* :323-324 */
        for (int j = 0; j < 4; j++) {
            const int xa = x0 + bx + 2 * j - 2;  /* column of the first sample point */
            const int xb = xa + 1;               /* column of the second (diagonal) point */
            const int c0 = 2 * tap(w, xa, y0 + r1);
            const int c1 = 2 * tap(w, xb, y0 + r2);
            sv  += o_abs(c0 - tap(w, xa, y0 + r0)     - tap(w, xa, y0 + r2))
                 + o_abs(c1 - tap(w, xb, y0 + r1)     - tap(w, xb, y0 + r3));
            sh  += o_abs(c0 - tap(w, xa - 1, y0 + r1) - tap(w, xa + 1, y0 + r1))
                 + o_abs(c1 - tap(w, xb - 1, y0 + r2) - tap(w, xb + 1, y0 + r2));
            sd0 += o_abs(c0 - tap(w, xa - 1, y0 + r0) - tap(w, xa + 1, y0 + r2))
                 + o_abs(c1 - tap(w, xb - 1, y0 + r1) - tap(w, xb + 1, y0 + r3));
            sd1 += o_abs(c0 - tap(w, xa + 1, y0 + r0) - tap(w, xa - 1, y0 + r2))
                 + o_abs(c1 - tap(w, xb + 1, y0 + r1) - tap(w, xb - 1, y0 + r3));
        }
    }

    {   /* alf_get_idx :270-297 */
        const int v_le_h = sv <= sh, d0_le_d1 = sd0 <= sd1;
        const int hv_hi = o_max(sv, sh),  hv_lo = o_min(sv, sh);
        const int d_hi  = o_max(sd0, sd1), d_lo = o_min(sd0, sd1);
        const int hv_wins = (uint64_t)d_hi * (uint64_t)hv_lo <= (uint64_t)hv_hi * (uint64_t)d_lo;
        const int hi = hv_wins ? hv_hi : d_hi, lo = hv_wins ? hv_lo : d_lo;
        int c = act_lut[o_clip_ubits(((sh + sv) * scale) >> (bd - 1), 4)];
        if (hi * 2 > 9 * lo)      c += (2 * hv_wins + 2) * 5;
        else if (hi > 2 * lo)     c += (2 * hv_wins + 1) * 5;
        *cls = c;
        *tr  = d0_le_d1 * 2 + v_le_h;
    }
}

static const uint8_t k_transpose_perm[4][12] = {   /* :387-392 */
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11 },
    { 9, 4, 10, 8, 1, 5, 11, 7, 3, 0, 2, 6 },
    { 0, 3, 2, 1, 8, 7, 6, 5, 4, 9, 10, 11 },
    { 9, 8, 10, 4, 3, 7, 11, 5, 1, 0, 2, 6 },
};
static const uint8_t k_clip_shift[4] = { 0, 3, 5, 7 };

static void block_filter(const VVCCudaALFSets *sets, int set_idx, int cls, int tr, int bd,
                         int coef[12], int clp[12])
{
    for (int j = 0; j < 12; j++) {
        const int src = k_transpose_perm[tr][j];
        if (set_idx < 16) {
            const int f = vvct_alf_class_to_filt_map[set_idx][cls];
            coef[j] = vvct_alf_fix_filt_coeff[f][src];
            clp[j]  = 1 << bd;
        } else {
            const int a = set_idx - 16;
            const int f = vvct_alf_aps_class_to_filt_map[cls];
            coef[j] = sets->luma_coeff[a][f][src];
            /* the reference indexes the clip table by class, not by mapped filter (:400) */
            clp[j]  = 1 << (bd - k_clip_shift[sets->luma_clip_idx[a][cls][src]]);
        }
    }
}

static void luma_ctb(const OPlane *dst, const Win *w, int x0, int y0, int cw, int ch,
                     int vb, int bd, const VVCCudaALFSets *sets, int set_idx)
{
    for (int by = 0; by < ch; by += 4)
        for (int bx = 0; bx < cw; bx += 4) {
            int cls, tr, f[12], c[12];
            classify_block(w, x0, y0, bx, by, vb, bd, &cls, &tr);
            block_filter(sets, set_idx, cls, tr, bd, f, c);
            for (int i = 0; i < 4; i++) {
                const int ry = by + i, t = ry - vb, y = y0 + ry;
                const int d1 = vb_reach(1, t, 4), d2 = vb_reach(2, t, 4), d3 = vb_reach(3, t, 4);
                const int near = (t == -1 || t == 0);
                for (int j = 0; j < 4; j++) {
                    const int x = x0 + bx + j, cur = tap(w, x, y);
                    int s = 0;
                    s += f[0]  * pair_clip(cur, tap(w, x,     y + d3), tap(w, x,     y - d3), c[0]);
                    s += f[1]  * pair_clip(cur, tap(w, x + 1, y + d2), tap(w, x - 1, y - d2), c[1]);
                    s += f[2]  * pair_clip(cur, tap(w, x,     y + d2), tap(w, x,     y - d2), c[2]);
                    s += f[3]  * pair_clip(cur, tap(w, x - 1, y + d2), tap(w, x + 1, y - d2), c[3]);
                    s += f[4]  * pair_clip(cur, tap(w, x + 2, y + d1), tap(w, x - 2, y - d1), c[4]);
                    s += f[5]  * pair_clip(cur, tap(w, x + 1, y + d1), tap(w, x - 1, y - d1), c[5]);
                    s += f[6]  * pair_clip(cur, tap(w, x,     y + d1), tap(w, x,     y - d1), c[6]);
                    s += f[7]  * pair_clip(cur, tap(w, x - 1, y + d1), tap(w, x + 1, y - d1), c[7]);
                    s += f[8]  * pair_clip(cur, tap(w, x - 2, y + d1), tap(w, x + 2, y - d1), c[8]);
                    s += f[9]  * pair_clip(cur, tap(w, x + 3, y),      tap(w, x - 3, y),      c[9]);
                    s += f[10] * pair_clip(cur, tap(w, x + 2, y),      tap(w, x - 2, y),      c[10]);
                    s += f[11] * pair_clip(cur, tap(w, x + 1, y),      tap(w, x - 1, y),      c[11]);
                    s = near ? (s + 512) >> 10 : (s + 64) >> 7;
                    dst->p[y * dst->pitch + x] = (pel)o_clip_pel(cur + s, bd);
                }
            }
        }
}

static void chroma_ctb(const OPlane *dst, const Win *w, int x0, int y0, int cw, int ch,
                       int vb, int bd, const int16_t f[6], const uint8_t clip_idx[6])
{
    int c[6];
    for (int i = 0; i < 6; i++)
        c[i] = 1 << (bd - k_clip_shift[clip_idx[i]]);
    for (int ry = 0; ry < ch; ry++) {
        const int t = ry - vb, y = y0 + ry;
        const int d1 = vb_reach(1, t, 2), d2 = vb_reach(2, t, 2);
        const int near = (t == -1 || t == 0);
        for (int rx = 0; rx < cw; rx++) {
            const int x = x0 + rx, cur = tap(w, x, y);
            int s = 0;
            s += f[0] * pair_clip(cur, tap(w, x,     y + d2), tap(w, x,     y - d2), c[0]);
            s += f[1] * pair_clip(cur, tap(w, x + 1, y + d1), tap(w, x - 1, y - d1), c[1]);
            s += f[2] * pair_clip(cur, tap(w, x,     y + d1), tap(w, x,     y - d1), c[2]);
            s += f[3] * pair_clip(cur, tap(w, x - 1, y + d1), tap(w, x + 1, y - d1), c[3]);
            s += f[4] * pair_clip(cur, tap(w, x + 2, y),      tap(w, x - 2, y),      c[4]);
            s += f[5] * pair_clip(cur, tap(w, x + 1, y),      tap(w, x - 1, y),      c[5]);
            s = near ? (s + 512) >> 10 : (s + 64) >> 7;
            dst->p[y * dst->pitch + x] = (pel)o_clip_pel(cur + s, bd);
        }
    }
}

/* CC-ALF: correction from pre-ALF luma added onto the (already ALF-filtered) chroma in dst. */
static void cc_ctb(const OPlane *dst, const Win *lw, int x0c, int y0c, int cw, int ch,
                   int hs, int vs, int vb_luma, int y0_luma, int bd, const int16_t f[7])
{
    for (int ry = 0; ry < ch; ry++) {
        const int ly_rel = ry << vs, t = ly_rel - vb_luma;
        int up = -1, dn = 1, dn2 = 2;
        if (!vs && (t == 0 || t == 1))
            continue;
        if (t == -2 || t == 1)        dn2 = 1;
        else if (t == -1 || t == 0)   up = dn = dn2 = 0;
        for (int rx = 0; rx < cw; rx++) {
            const int lx = (x0c + rx) << hs, ly = y0_luma + ly_rel;
            const int cur = tap(lw, lx, ly);
            pel *d = &dst->p[(y0c + ry) * dst->pitch + x0c + rx];
            int s = 0;
            s += f[0] * (tap(lw, lx,     ly + up)  - cur);
            s += f[1] * (tap(lw, lx - 1, ly)       - cur);
            s += f[2] * (tap(lw, lx + 1, ly)       - cur);
            s += f[3] * (tap(lw, lx - 1, ly + dn)  - cur);
            s += f[4] * (tap(lw, lx,     ly + dn)  - cur);
            s += f[5] * (tap(lw, lx + 1, ly + dn)  - cur);
            s += f[6] * (tap(lw, lx,     ly + dn2) - cur);
            s = o_clip3((s + 64) >> 7, -(1 << (bd - 1)), (1 << (bd - 1)) - 1);
            *d = (pel)o_clip_pel(*d + s, bd);
        }
    }
}

static void copy_rect(const OPlane *d, const OPlane *s, int x0, int y0, int w, int h)
{
    for (int y = 0; y < h; y++)
        memcpy(&d->p[(y0 + y) * d->pitch + x0], &s->p[(y0 + y) * s->pitch + x0], w * sizeof(pel));
}

void vvco_alf_frame(const VVCCudaFrame *dstf, const VVCCudaFrame *srcf,
                    const VVCCudaALFCtb *ctbs, const VVCCudaALFSets *sets, int sets_per_frame)
{
    const int ctb = 1 << srcf->ctb_log2, cols = o_ctb_cols(srcf), rows = o_ctb_rows(srcf);
    const int bd = srcf->bit_depth;
    const int planes = srcf->chroma_format_idc ? 3 : 1;

    for (int k = 0; k < srcf->batch; k++) {
        const VVCCudaALFSets *fs = sets + (sets_per_frame ? k : 0);
        const OPlane sl = o_plane(srcf, 0, k);
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++) {
                const VVCCudaALFCtb *a = &ctbs[(size_t)k * cols * rows + cy * cols + cx];
                unsigned edges = a->edges;
                const int x0 = cx * ctb, y0 = cy * ctb;
                const int lw = o_min(ctb, srcf->width - x0), lh = o_min(ctb, srcf->height - y0);
                if (cx == 0)        edges |= VVC_CUDA_EDGE_LEFT;
                if (cy == 0)        edges |= VVC_CUDA_EDGE_TOP;
                if (cx == cols - 1) edges |= VVC_CUDA_EDGE_RIGHT;
                if (cy == rows - 1) edges |= VVC_CUDA_EDGE_BOTTOM;
                const Win lwin = make_win(&sl, x0, y0, lw, lh, edges);

                for (int c = 0; c < planes; c++) {
                    const OPlane s = o_plane(srcf, c, k), d = o_plane(dstf, c, k);
                    const int hs = c ? srcf->hshift : 0, vs = c ? srcf->vshift : 0;
                    const int px = x0 >> hs, py = y0 >> vs, pw = lw >> hs, ph = lh >> vs;
                    const Win win = make_win(&s, px, py, pw, ph, edges);
                    if (!a->ctb_flag[c])
                        copy_rect(&d, &s, px, py, pw, ph);
                    else if (c == 0)
                        luma_ctb(&d, &win, px, py, pw, ph, ctb - 4, bd, fs, a->filt_set_idx_y);
                    else
                        chroma_ctb(&d, &win, px, py, pw, ph, (ctb >> vs) - 2, bd,
                                   fs->chroma_coeff[a->chroma_alt_idx[c - 1]],
                                   fs->chroma_clip_idx[a->chroma_alt_idx[c - 1]]);
                    if (c && a->cc_idc[c - 1])
                        cc_ctb(&d, &lwin, px, py, pw, ph, hs, vs, ctb - 4, y0, bd,
                               fs->cc_coeff[c - 1][a->cc_idc[c - 1] - 1]);
                }
            }
    }
}
