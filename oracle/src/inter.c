/*
 * inter.c - CPU oracle of the inter prediction stage (TEST INFRASTRUCTURE, see vvc_oracle.h).
 *
 * Restates, against the VVCCudaPB descriptors of include/vvcdsp_cuda.h, what the reference does
 * for one motion-compensated block:
 *   pred_regular_blk / derive_sb_mv / pred_regular_luma / pred_regular_chroma
 *                                                   libavcodec/vvc/vvc_inter.c:545-639,764-811
 *   dmvr_mv_refine + parametric_mv_refine           vvc_inter.c:642-748
 *   luma_prof_uni / luma_prof_bi (affine, PROF)     vvc_inter.c:368-446
 *   pred_gpm_blk                                    vvc_inter.c:466-521
 *   emulated_edge / _dmvr / _bilinear               vvc_inter.c:33-110 (+ videodsp_template.c:26-105)
 * and the table entries those call:
 *   put / put_uni / put_uni_w (pixels,h,v,hv; 8-tap luma, 4-tap chroma)
 *                                                   libavcodec/h26x/h2656_inter_template.c:29-577
 *   avg, w_avg, put_gpm, bdof_fetch_samples, fetch_samples, prof_grad_filter, apply_prof*,
 *   apply_bdof, dmvr*                               libavcodec/vvc/vvc_inter_template.c:25-436
 *   sad, pad_int16                                  libavcodec/vvc/vvcdsp.c:29-65
 *
 * Written per sample on clamped reference coordinates instead of through padded scratch copies:
 * ff_emulated_edge_mc materialises exactly "coordinate clamped to a window" (picture, or for DMVR
 * the unrefined block's window intersected with the picture).
 */
#include "vvc_oracle.h"
#include "vvc_tables_c.h"

#define TS 128                      /* MAX_PB_SIZE: row pitch of the reference's int16 tiles */
#define RING (TS + 32)              /* PROF_TEMP_OFFSET: room for the one-sample ring, vvc_inter.c:30 */

typedef struct Win {
    const pel *p;
    ptrdiff_t  pitch;
    int        xlo, xhi, ylo, yhi;  /* coordinates are clamped into this window before the fetch */
} Win;

static inline int wget(const Win *w, int x, int y)
{
    x = o_clip3(x, w->xlo, w->xhi);
    y = o_clip3(y, w->ylo, w->yhi);
    return w->p[(ptrdiff_t)y * w->pitch + x];
}

static Win pic_win(const VVCCudaFrame *f, int slot, int c)
{
    const OPlane pl = o_plane(f, c, slot);
    Win w = { pl.p, pl.pitch, 0, pl.w - 1, 0, pl.h - 1 };
    return w;
}

/* emulated_edge_dmvr (vvc_inter.c:60-89): the window of the UNREFINED block, cut to the picture */
static Win dmvr_win(Win pic, int x_sb, int y_sb, int bw, int bh, int before, int after)
{
    const int pw = pic.xhi + 1, ph = pic.yhi + 1;
    const int sx = o_min(o_max(x_sb - before, 0), pw - 1);
    const int sy = o_min(o_max(y_sb - before, 0), ph - 1);
    const int ww = o_max(o_min(pw, x_sb + bw + after) - sx, 1);
    const int hh = o_max(o_min(ph, y_sb + bh + after) - sy, 1);
    Win w = pic;
    w.xlo = sx; w.xhi = sx + ww - 1;
    w.ylo = sy; w.yhi = sy + hh - 1;
    return w;
}

/* One sample of the separable interpolation before the final rounding, h2656_inter_template.c:
 * put_pixels :29, put_luma_h/v/hv :97-150, put_chroma_h/v/hv :342-395.  taps = 8 (luma) or 4. */
static int interp(const Win *r, int x, int y, int mx, int my, const int8_t *hf, const int8_t *vf, int taps, int bd)
{
    const int before = taps / 2 - 1;
    if (!mx && !my)
        return wget(r, x, y) << (14 - bd);
    if (!my) {
        int s = 0;
        for (int k = 0; k < taps; k++)
            s += hf[k] * wget(r, x + k - before, y);
        return s >> (bd - 8);
    }
    if (!mx) {
        int s = 0;
        for (int k = 0; k < taps; k++)
            s += vf[k] * wget(r, x, y + k - before);
        return s >> (bd - 8);
    }
    {
        int s = 0;
        for (int j = 0; j < taps; j++) {
            int t = 0;
            for (int k = 0; k < taps; k++)
                t += hf[k] * wget(r, x + k - before, y + j - before);
            s += vf[j] * (int16_t)(t >> (bd - 8));       /* first stage lives in int16_t tmp_array */
        }
        return s >> 6;
    }
}

typedef struct Weights { int on, denom, w0, w1, o0, o1; } Weights;

/* derive_weight (vvc_inter.c:148-177) for plane c */
static Weights bi_weights(const VVCCudaPB *pb, const VVCCudaWP *wp, int c)
{
    Weights w = { 0 };
    const int weighted = (pb->flags & VVC_CUDA_PB_WEIGHTED) && !(pb->flags & VVC_CUDA_PB_DMVR);
    static const int bcw_lut[5] = { 4, 5, 3, 10, -2 };
    if (pb->bcw_idx) {
        w.on = 1; w.denom = 2; w.w1 = bcw_lut[pb->bcw_idx]; w.w0 = 8 - w.w1;
    } else if (weighted) {
        const VVCCudaWP *e = &wp[pb->wp];
        w.on = 1; w.denom = e->log2_denom[c > 0];
        w.w0 = e->weight[0][c]; w.w1 = e->weight[1][c];
        w.o0 = e->offset[0][c]; w.o1 = e->offset[1][c];
    }
    return w;
}

/* avg / w_avg, vvc_inter_template.c:25-57 */
static inline int combine_bi(int a, int b, const Weights *w, int bd)
{
    if (!w->on) {
        const int shift = o_max(3, 15 - bd);
        return o_clip_pel((a + b + (1 << (shift - 1))) >> shift, bd);
    } else {
        const int shift = w->denom + o_max(3, 15 - bd);
        const int offset = (((w->o0 + w->o1) << (bd - 8)) + 1) << (shift - 1);
        return o_clip_pel((a * w->w0 + b * w->w1 + offset) >> shift, bd);
    }
}

/* put_uni_* / put_uni_w_* final rounding, h2656_inter_template.c:154-334 */
static inline int finish_uni(int val, const VVCCudaPB *pb, const VVCCudaWP *wp, int lx, int c, int bd)
{
    if (pb->flags & VVC_CUDA_PB_WEIGHTED) {
        const VVCCudaWP *e = &wp[pb->wp];
        const int shift = e->log2_denom[c > 0] + 14 - bd;
        const int ox = e->offset[lx][c] * (1 << (bd - 8));
        return o_clip_pel(((val * e->weight[lx][c] + (1 << (shift - 1))) >> shift) + ox, bd);
    }
    return o_clip_pel((val + (1 << (13 - bd))) >> (14 - bd), bd);
}

/* ---- DMVR: 8.5.3, dmvr_mv_refine vvc_inter.c:685-748 --------------------------------------------- */
/* dmvr / dmvr_h / dmvr_v / dmvr_hv, vvc_inter_template.c:324-415 */
static int bilinear(const Win *r, int x, int y, int mx, int my, int bd)
{
    const int8_t *fx = vvct_dmvr_filters[mx], *fy = vvct_dmvr_filters[my];
    const int shift1 = bd - 6, off1 = 1 << (shift1 - 1);
    if (!mx && !my) {
        const int s = wget(r, x, y);
        return bd > 10 ? (s + (1 << (bd - 11))) >> (bd - 10) : s << (10 - bd);
    }
    if (!my)
        return (fx[0] * wget(r, x, y) + fx[1] * wget(r, x + 1, y) + off1) >> shift1;
    if (!mx)
        return (fy[0] * wget(r, x, y) + fy[1] * wget(r, x, y + 1) + off1) >> shift1;
    {
        const int t0 = (int16_t)((fx[0] * wget(r, x, y)     + fx[1] * wget(r, x + 1, y)     + off1) >> shift1);
        const int t1 = (int16_t)((fx[0] * wget(r, x, y + 1) + fx[1] * wget(r, x + 1, y + 1) + off1) >> shift1);
        return (fy[0] * t0 + fy[1] * t1 + 8) >> 4;
    }
}

/* vvc_sad, vvcdsp.c:49-65 */
static int sad(const int16_t *a, const int16_t *b, int dx, int dy, int w, int h)
{
    int s = 0;
    dx -= 2; dy -= 2;
    a += (2 + dy) * TS + 2 + dx;
    b += (2 - dy) * TS + 2 - dx;
    for (int y = 0; y < h; y += 2, a += 2 * TS, b += 2 * TS)
        for (int x = 0; x < w; x++)
            s += o_abs(a[x] - b[x]);
    return s;
}

/* parametric_mv_refine, vvc_inter.c:642-681 */
static int parametric(const int *s, int stride)
{
    const int sm = s[-stride], sc = s[0], sp = s[stride];
    int denom = ((sm + sp) - (sc << 1)) << 3;
    if (!denom)
        return 0;
    if (sm == sc)
        return -8;
    if (sp == sc)
        return 8;
    {
        int num = (sm - sp) * 16, neg = 0, q = 0;
        if (num < 0) { num = -num; neg = 1; }
        for (int i = 0; i < 3; i++) {
            q <<= 1;
            if (num >= denom) { num -= denom; q++; }
            denom >>= 1;
        }
        return neg ? -q : q;
    }
}

static void dmvr_refine(const VVCCudaFrame *refs, const VVCCudaPB *pb, int mv[2][2], int *bdof, VVCCudaDmvrOut *out)
{
    static _Thread_local int16_t t[2][(16 + 4) * TS];
    const int w = pb->w, h = pb->h, bd = refs->bit_depth;
    int s[5][5], min_dx = 2, min_dy = 2, min_sad;
    for (int i = 0; i < 2; i++) {
        const Win r = pic_win(refs, pb->ref[i], 0);
        const int mx = mv[i][0] & 15, my = mv[i][1] & 15;
        const int ox = pb->x0 + (mv[i][0] >> 4) - 2, oy = pb->y0 + (mv[i][1] >> 4) - 2;
        for (int y = 0; y < h + 4; y++)
            for (int x = 0; x < w + 4; x++)
                t[i][y * TS + x] = (int16_t)bilinear(&r, ox + x, oy + y, mx, my, bd);
    }
    min_sad = sad(t[0], t[1], 2, 2, w, h);
    min_sad -= min_sad >> 2;
    s[2][2] = min_sad;
    if (min_sad >= w * h) {
        int dmv[2];
        for (int dy = 0; dy < 5; dy++)
            for (int dx = 0; dx < 5; dx++) {
                if (dx == 2 && dy == 2)
                    continue;
                s[dy][dx] = sad(t[0], t[1], dx, dy, w, h);
                if (s[dy][dx] < min_sad) {
                    min_sad = s[dy][dx]; min_dx = dx; min_dy = dy;
                }
            }
        dmv[0] = (min_dx - 2) * 16;
        dmv[1] = (min_dy - 2) * 16;
        if (min_dx != 0 && min_dx != 4 && min_dy != 0 && min_dy != 4) {
            dmv[0] += parametric(&s[min_dy][min_dx], 1);
            dmv[1] += parametric(&s[min_dy][min_dx], 5);
        }
        for (int i = 0; i < 2; i++) {
            mv[i][0] = o_clip3(mv[i][0] + (1 - 2 * i) * dmv[0], -(1 << 17), (1 << 17) - 1);   /* ff_vvc_clip_mv */
            mv[i][1] = o_clip3(mv[i][1] + (1 - 2 * i) * dmv[1], -(1 << 17), (1 << 17) - 1);
        }
    }
    if (min_sad < 2 * w * h)
        *bdof = 0;
    if (out) {
        memcpy(out->mv, mv, sizeof(out->mv));
        out->min_sad = min_sad;
        out->bdof_applied = *bdof;
    }
}

/* ---- BDOF: apply_bdof, vvc_inter_template.c:237-317 ------------------------------------------- */
/* pad_int16, vvcdsp.c:29-47 */
static void pad16(int16_t *t, int pitch, int w, int h)
{
    for (int y = 0; y < h; y++) {
        t[y * pitch - 1] = t[y * pitch];
        t[y * pitch + w] = t[y * pitch + w - 1];
    }
    memcpy(t - 1 - pitch, t - 1, (w + 2) * sizeof(int16_t));
    memcpy(t - 1 + h * pitch, t - 1 + (h - 1) * pitch, (w + 2) * sizeof(int16_t));
}

static inline int vsign(int v) { return v < 0 ? -1 : !!v; }

static void bdof(pel *dst, ptrdiff_t dp, int16_t *s0, int16_t *s1, int w, int h, int bd)
{
    enum { GP = 18 };
    int16_t gh[2][GP * GP], gv[2][GP * GP];
    int16_t *s[2] = { s0, s1 };
    for (int i = 0; i < 2; i++) {
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int16_t *p = s[i] + y * TS + x;
                gh[i][(y + 1) * GP + x + 1] = (int16_t)((p[1] >> 6) - (p[-1] >> 6));
                gv[i][(y + 1) * GP + x + 1] = (int16_t)((p[TS] >> 6) - (p[-TS] >> 6));
            }
        pad16(gh[i] + GP + 1, GP, w, h);
        pad16(gv[i] + GP + 1, GP, w, h);
        pad16(s[i], TS, w, h);
    }
    for (int by = 0; by < h; by += 4)
        for (int bx = 0; bx < w; bx += 4) {
            int sgx2 = 0, sgy2 = 0, sgxgy = 0, sgxdi = 0, sgydi = 0, vx, vy;
            for (int y = 0; y < 6; y++)
                for (int x = 0; x < 6; x++) {
                    const int ti = (by + y - 1) * TS + bx + x - 1, gi = (by + y) * GP + bx + x;
                    const int diff = (s0[ti] >> 4) - (s1[ti] >> 4);
                    const int th = (gh[0][gi] + gh[1][gi]) >> 1, tv = (gv[0][gi] + gv[1][gi]) >> 1;
                    sgx2 += o_abs(th);
                    sgy2 += o_abs(tv);
                    sgxgy += vsign(tv) * th;
                    sgxdi += -vsign(th) * diff;
                    sgydi += -vsign(tv) * diff;
                }
            vx = sgx2 > 0 ? o_clip3((sgxdi * 4) >> o_ilog2(sgx2), -15, 15) : 0;
            vy = sgy2 > 0 ? o_clip3(((sgydi * 4) - ((vx * sgxgy) >> 1)) >> o_ilog2(sgy2), -15, 15) : 0;
            for (int y = 0; y < 4; y++)
                for (int x = 0; x < 4; x++) {
                    const int ti = (by + y) * TS + bx + x, gi = (by + y + 1) * GP + bx + x + 1;
                    const int off = vx * (gh[0][gi] - gh[1][gi]) + vy * (gv[0][gi] - gv[1][gi]);
                    const int shift = 15 - bd;
                    dst[(by + y) * dp + bx + x] = (pel)o_clip_pel((s0[ti] + (1 << (shift - 1)) + s1[ti] + off) >> shift, bd);
                }
        }
}

/* bdof_fetch_samples / fetch_samples, vvc_inter_template.c:101-133: the ring of integer samples */
static void fetch_ring(int16_t *t, const Win *r, int ox, int oy, int mx, int my, int w, int h, int bd)
{
    const int xo = (mx >> 3) - 1, yo = (my >> 3) - 1;
    for (int y = 0; y < h + 2; y++)
        for (int x = 0; x < w + 2; x++)
            if (y == 0 || y == h + 1 || x == 0 || x == w + 1)
                t[(y - 1) * TS + x - 1] = (int16_t)(wget(r, ox + xo + x, oy + yo + y) << (14 - bd));
}

/* gradient + displacement of one PROF sample, apply_prof* vvc_inter_template.c:160-235 */
static inline int prof_val(const int16_t *t, int x, int y, const int16_t *dx, const int16_t *dy, int bd)
{
    const int16_t *p = t + y * TS + x;
    const int limit = 1 << o_max(13, bd + 1);
    const int16_t g_h = (int16_t)((p[1] >> 6) - (p[-1] >> 6)), g_v = (int16_t)((p[TS] >> 6) - (p[-TS] >> 6));
    const int di = g_h * dx[y * 4 + x] + g_v * dy[y * 4 + x];
    return p[0] + o_clip3(di, -limit, limit - 1);
}

/* ---- one record --------------------------------------------------------------------------------- */
static void predict_pb(const VVCCudaFrame *dst, const VVCCudaFrame *refs, const VVCCudaPB *pb,
                       const VVCCudaWP *wp, const VVCCudaProf *prof, VVCCudaDmvrOut *out)
{
    static _Thread_local int16_t tile[2][RING + 17 * TS];
    const int bd = dst->bit_depth, w = pb->w, h = pb->h;
    const int gpm = pb->flags & VVC_CUDA_PB_GPM;
    const int bi = gpm || pb->pred_flag == 3;
    int mv[2][2], orig[2][2], do_bdof = !!(pb->flags & VVC_CUDA_PB_BDOF);
    memcpy(mv, pb->mv, sizeof(mv));
    memcpy(orig, pb->mv, sizeof(mv));
    if ((pb->flags & VVC_CUDA_PB_DMVR) && (pb->planes & VVC_CUDA_PB_LUMA))
        dmvr_refine(refs, pb, mv, &do_bdof, out);
    const int dmvr = pb->flags & VVC_CUDA_PB_DMVR;

    for (int c = 0; c < (dst->chroma_format_idc ? 3 : 1); c++) {
        if (!(pb->planes & (c ? VVC_CUDA_PB_CHROMA : VVC_CUDA_PB_LUMA)))
            continue;
        const int sh = c ? 1 : 0, taps = c ? 4 : 8;
        const int bw = w >> sh, bh = h >> sh, x0 = pb->x0 >> sh, y0 = pb->y0 >> sh;
        const OPlane dp = o_plane(dst, c, pb->pic);
        pel *d = dp.p + (ptrdiff_t)y0 * dp.pitch + x0;
        Win r[2];
        int ox[2], oy[2], mx[2], my[2];
        const int8_t *hf[2], *vf[2];
        for (int i = 0; i < 2; i++) {
            if (!gpm && !(pb->pred_flag & (1 << i)))
                continue;
            const int filt = gpm ? 0 : pb->filt;
            r[i] = pic_win(refs, pb->ref[i], c);
            ox[i] = x0 + (mv[i][0] >> (4 + sh));
            oy[i] = y0 + (mv[i][1] >> (4 + sh));
            mx[i] = c ? mv[i][0] & 31 : mv[i][0] & 15;          /* av_mod_uintp2(mv, 4 + hs) << (1 - hs), hs = 1 */
            my[i] = c ? mv[i][1] & 31 : mv[i][1] & 15;
            hf[i] = c ? vvct_chroma_mc_filters[0][mx[i]] : vvct_luma_mc_filters[filt][mx[i]];
            vf[i] = c ? vvct_chroma_mc_filters[0][my[i]] : vvct_luma_mc_filters[filt][my[i]];
            if (dmvr)
                r[i] = dmvr_win(r[i], x0 + (orig[i][0] >> (4 + sh)), y0 + (orig[i][1] >> (4 + sh)), bw, bh,
                                c ? 1 : 3, c ? 2 : 4);
        }
        if (!bi) {
            const int lx = pb->pred_flag - 1;
            const int use_prof = !c && (pb->flags & (lx ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0));
            if (use_prof) {
                int16_t *t = tile[0] + RING;
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++)
                        t[y * TS + x] = (int16_t)interp(&r[lx], ox[lx] + x, oy[lx] + y, mx[lx], my[lx], hf[lx], vf[lx], 8, bd);
                fetch_ring(t, &r[lx], ox[lx], oy[lx], mx[lx], my[lx], 4, 4, bd);
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++)
                        d[y * dp.pitch + x] = (pel)finish_uni(prof_val(t, x, y, prof[pb->prof].diff_mv_x[lx], prof[pb->prof].diff_mv_y[lx], bd),
                                                              pb, wp, lx, 0, bd);
            } else {
                for (int y = 0; y < bh; y++)
                    for (int x = 0; x < bw; x++)
                        d[y * dp.pitch + x] = (pel)finish_uni(interp(&r[lx], ox[lx] + x, oy[lx] + y, mx[lx], my[lx], hf[lx], vf[lx], taps, bd),
                                                              pb, wp, lx, c, bd);
            }
            continue;
        }
        /* two predictions into int16 tiles, as inter.put stores them */
        for (int i = 0; i < 2; i++) {
            int16_t *t = tile[i] + RING;
            const int use_prof = !c && !gpm && (pb->flags & (i ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0));
            for (int y = 0; y < bh; y++)
                for (int x = 0; x < bw; x++)
                    t[y * TS + x] = (int16_t)interp(&r[i], ox[i] + x, oy[i] + y, mx[i], my[i], hf[i], vf[i], taps, bd);
            if (use_prof) {                                   /* luma_prof_bi, vvc_inter.c:410-446 */
                int16_t o[16];
                fetch_ring(t, &r[i], ox[i], oy[i], mx[i], my[i], 4, 4, bd);
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++)
                        o[y * 4 + x] = (int16_t)prof_val(t, x, y, prof[pb->prof].diff_mv_x[i], prof[pb->prof].diff_mv_y[i], bd);
                for (int y = 0; y < 4; y++)
                    for (int x = 0; x < 4; x++)
                        t[y * TS + x] = o[y * 4 + x];
            }
            if (!c && do_bdof && !gpm)
                fetch_ring(t, &r[i], ox[i], oy[i], mx[i], my[i], bw, bh, bd);
        }
        if (gpm) {                                            /* put_gpm, vvc_inter_template.c:78-98 */
            const int shift = o_max(5, 17 - bd);
            const int sx = pb->gpm_step_x << sh, sy = pb->gpm_step_y << sh;
            const uint8_t *wt = &vvct_gpm_weights[0][0] + pb->gpm_weights;
            for (int y = 0; y < bh; y++)
                for (int x = 0; x < bw; x++) {
                    const int a = tile[0][RING + y * TS + x], b = tile[1][RING + y * TS + x];
                    const int g = wt[y * sy + x * sx];
                    d[y * dp.pitch + x] = (pel)o_clip_pel((a * g + b * (8 - g) + (1 << (shift - 1))) >> shift, bd);
                }
        } else if (!c && do_bdof) {
            bdof(d, dp.pitch, tile[0] + RING, tile[1] + RING, bw, bh, bd);
        } else {
            const Weights wt = bi_weights(pb, wp, c);
            for (int y = 0; y < bh; y++)
                for (int x = 0; x < bw; x++)
                    d[y * dp.pitch + x] = (pel)combine_bi(tile[0][RING + y * TS + x], tile[1][RING + y * TS + x], &wt, bd);
        }
    }
}

void vvco_inter_frame(const VVCCudaFrame *dst, const VVCCudaFrame *refs, const VVCCudaPB *pbs, int n_pbs,
                      const VVCCudaWP *wp, const VVCCudaProf *prof, VVCCudaDmvrOut *dmvr_out)
{
    for (int i = 0; i < n_pbs; i++)
        predict_pb(dst, refs, &pbs[i], wp, prof, dmvr_out ? &dmvr_out[i] : NULL);
}
