/*
 * Oracle: deblocking of a whole picture (TEST INFRASTRUCTURE, see vvc_oracle.h).
 *
 * Restates the reference's pixel work, in place and in the reference's order (all vertical
 * edges CTU by CTU in raster order, then all horizontal edges):
 *   walk            ff_vvc_deblock_vertical / _horizontal   libavcodec/vvc/vvc_filter.c:861-1003
 *   luma decisions  vvc_loop_filter_luma                    libavcodec/vvc/vvc_filter_template.c:546-631
 *   long filter     loop_filter_luma_large                  libavcodec/vvc/vvc_filter_template.c:466-544
 *   strong / weak   loop_filter_luma_strong / _weak         libavcodec/h26x/h2656_deblock_template.c:25-81
 *   chroma          vvc_loop_filter_chroma                  libavcodec/vvc/vvc_filter_template.c:681-754
 *                   loop_filter_chroma_strong(_one_side)    :633-679,  loop_filter_chroma_weak  h2656_deblock_template.c:83-99
 * The per-segment (tc, beta, max lengths) arrive precomputed in the edge maps, exactly as the
 * reference driver hands them to the DSP entries.
 */
#include "vvc_oracle.h"

/* sample k positions across the edge: k = 0 is Q0, k = -1 is P0 */
#define AT(base, k) ((base)[(k) * xs])

static void luma_long(pel *pix, ptrdiff_t xs, ptrdiff_t ys, int tc, int lp, int lq)
{
    static const uint8_t w3[3] = { 53, 32, 11 }, w5[5] = { 58, 45, 32, 19, 6 }, w7[7] = { 59, 50, 41, 32, 23, 14, 5 };
    static const uint8_t k3[3] = { 6, 4, 2 }, k5[5] = { 6, 5, 4, 3, 2 }, k7[7] = { 6, 5, 4, 3, 2, 1, 1 };
    const uint8_t *wp = lp == 3 ? w3 : lp == 5 ? w5 : w7, *kp = lp == 3 ? k3 : lp == 5 ? k5 : k7;
    const uint8_t *wq = lq == 3 ? w3 : lq == 5 ? w5 : w7, *kq = lq == 3 ? k3 : lq == 5 ? k5 : k7;

    for (int line = 0; line < 4; line++, pix += ys) {
        int p[8], q[8], m;
        for (int i = 0; i < 8; i++) { p[i] = AT(pix, -1 - i); q[i] = AT(pix, i); }
        if (lp == 5 && lq == 5)
            m = (p[4] + p[3] + 2 * (p[2] + p[1] + p[0] + q[0] + q[1] + q[2]) + q[3] + q[4] + 8) >> 4;
        else if (lp == lq)
            m = (p[6] + p[5] + p[4] + p[3] + p[2] + p[1] + 2 * (p[0] + q[0]) + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
        else if (lp + lq == 12)
            m = (p[5] + p[4] + p[3] + p[2] + 2 * (p[1] + p[0] + q[0] + q[1]) + q[2] + q[3] + q[4] + q[5] + 8) >> 4;
        else if (lp + lq == 8)
            m = (p[3] + p[2] + p[1] + p[0] + q[0] + q[1] + q[2] + q[3] + 4) >> 3;
        else if (lq == 7)
            m = (2 * (p[2] + p[1] + p[0] + q[0]) + p[0] + p[1] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + 8) >> 4;
        else
            m = (p[6] + p[5] + p[4] + p[3] + p[2] + p[1] + 2 * (q[2] + q[1] + q[0] + p[0]) + q[0] + q[1] + 8) >> 4;
        {
            const int ref = (p[lp] + p[lp - 1] + 1) >> 1;
            for (int i = 0; i < lp; i++) {
                const int lim = (tc * kp[i]) >> 1;
                AT(pix, -1 - i) = (pel)(p[i] + o_clip3(((m * wp[i] + ref * (64 - wp[i]) + 32) >> 6) - p[i], -lim, lim));
            }
        }
        {
            const int ref = (q[lq] + q[lq - 1] + 1) >> 1;
            for (int i = 0; i < lq; i++) {
                const int lim = (tc * kq[i]) >> 1;
                AT(pix, i) = (pel)(q[i] + o_clip3(((m * wq[i] + ref * (64 - wq[i]) + 32) >> 6) - q[i], -lim, lim));
            }
        }
    }
}

static void luma_strong(pel *pix, ptrdiff_t xs, ptrdiff_t ys, int tc)
{
    for (int line = 0; line < 4; line++, pix += ys) {
        const int p3 = AT(pix, -4), p2 = AT(pix, -3), p1 = AT(pix, -2), p0 = AT(pix, -1);
        const int q0 = AT(pix, 0), q1 = AT(pix, 1), q2 = AT(pix, 2), q3 = AT(pix, 3);
        AT(pix, -1) = (pel)(p0 + o_clip3(((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3) - p0, -3 * tc, 3 * tc));
        AT(pix, -2) = (pel)(p1 + o_clip3(((p2 + p1 + p0 + q0 + 2) >> 2) - p1, -2 * tc, 2 * tc));
        AT(pix, -3) = (pel)(p2 + o_clip3(((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3) - p2, -tc, tc));
        AT(pix, 0)  = (pel)(q0 + o_clip3(((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3) - q0, -3 * tc, 3 * tc));
        AT(pix, 1)  = (pel)(q1 + o_clip3(((p0 + q0 + q1 + q2 + 2) >> 2) - q1, -2 * tc, 2 * tc));
        AT(pix, 2)  = (pel)(q2 + o_clip3(((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3) - q2, -tc, tc));
    }
}

static void luma_weak(pel *pix, ptrdiff_t xs, ptrdiff_t ys, int tc, int np, int nq, int bd)
{
    const int half = tc >> 1;
    for (int line = 0; line < 4; line++, pix += ys) {
        const int p2 = AT(pix, -3), p1 = AT(pix, -2), p0 = AT(pix, -1);
        const int q0 = AT(pix, 0), q1 = AT(pix, 1), q2 = AT(pix, 2);
        int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
        if (o_abs(delta) >= 10 * tc)
            continue;
        delta = o_clip3(delta, -tc, tc);
        AT(pix, -1) = (pel)o_clip_pel(p0 + delta, bd);
        AT(pix, 0)  = (pel)o_clip_pel(q0 - delta, bd);
        if (np > 1)
            AT(pix, -2) = (pel)o_clip_pel(p1 + o_clip3((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -half, half), bd);
        if (nq > 1)
            AT(pix, 1)  = (pel)o_clip_pel(q1 + o_clip3((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -half, half), bd);
    }
}

static inline int curv(const pel *l, ptrdiff_t xs, int a, int b, int c) { return o_abs(AT(l, a) - 2 * AT(l, b) + AT(l, c)); }

/* one 4-line luma segment */
static void luma_segment(pel *pix, ptrdiff_t xs, ptrdiff_t ys, int tc_in, int beta_in, int lp, int lq,
                         int hor_ctu_edge, int bd)
{
    const int tc = tc_in << (bd - 10), beta = beta_in << (bd - 8);
    pel *l0 = pix, *l3 = pix + 3 * ys;
    if (!tc)
        return;
    {
        const int dp0 = curv(l0, xs, -3, -2, -1), dq0 = curv(l0, xs, 2, 1, 0);
        const int dp3 = curv(l3, xs, -3, -2, -1), dq3 = curv(l3, xs, 2, 1, 0);
        const int d0 = dp0 + dq0, d3 = dp3 + dq3;
        const int tc25 = (tc * 5 + 1) >> 1;
        const int big_p = lp > 3 && !hor_ctu_edge, big_q = lq > 3;

        if (big_p || big_q) {
            const int dp0l = big_p ? (dp0 + curv(l0, xs, -6, -5, -4) + 1) >> 1 : dp0;
            const int dq0l = big_q ? (dq0 + curv(l0, xs, 5, 4, 3) + 1) >> 1 : dq0;
            const int dp3l = big_p ? (dp3 + curv(l3, xs, -6, -5, -4) + 1) >> 1 : dp3;
            const int dq3l = big_q ? (dq3 + curv(l3, xs, 5, 4, 3) + 1) >> 1 : dq3;
            const int d0l = dp0l + dq0l, d3l = dp3l + dq3l;
            lp = big_p ? lp : 3;             /* sticky: the normal decision below sees these (:591-592) */
            lq = big_q ? lq : 3;
            if (d0l + d3l < beta) {
                const int b53 = (beta * 3) >> 5, b4 = beta >> 4;
                int sp0 = o_abs(AT(l0, -4) - AT(l0, -1)) + (lp == 7 ? o_abs(AT(l0, -8) - AT(l0, -7) - AT(l0, -6) + AT(l0, -5)) : 0);
                int sq0 = o_abs(AT(l0, 0) - AT(l0, 3))   + (lq == 7 ? o_abs(AT(l0, 4) - AT(l0, 5) - AT(l0, 6) + AT(l0, 7)) : 0);
                int sp3 = o_abs(AT(l3, -4) - AT(l3, -1)) + (lp == 7 ? o_abs(AT(l3, -8) - AT(l3, -7) - AT(l3, -6) + AT(l3, -5)) : 0);
                int sq3 = o_abs(AT(l3, 0) - AT(l3, 3))   + (lq == 7 ? o_abs(AT(l3, 4) - AT(l3, 5) - AT(l3, 6) + AT(l3, 7)) : 0);
                if (big_p) {
                    sp0 = (sp0 + o_abs(AT(l0, -4) - AT(l0, -1 - lp)) + 1) >> 1;
                    sp3 = (sp3 + o_abs(AT(l3, -4) - AT(l3, -1 - lp)) + 1) >> 1;
                }
                if (big_q) {
                    sq0 = (sq0 + o_abs(AT(l0, 3) - AT(l0, lq)) + 1) >> 1;
                    sq3 = (sq3 + o_abs(AT(l3, 3) - AT(l3, lq)) + 1) >> 1;
                }
                if (sp0 + sq0 < b53 && o_abs(AT(l0, -1) - AT(l0, 0)) < tc25 &&
                    sp3 + sq3 < b53 && o_abs(AT(l3, -1) - AT(l3, 0)) < tc25 &&
                    (d0l << 1) < b4 && (d3l << 1) < b4) {
                    luma_long(pix, xs, ys, tc, lp, lq);
                    return;
                }
            }
        }
        if (d0 + d3 < beta) {
            if (lp > 2 && lq > 2 &&
                o_abs(AT(l0, -4) - AT(l0, -1)) + o_abs(AT(l0, 3) - AT(l0, 0)) < (beta >> 3) && o_abs(AT(l0, -1) - AT(l0, 0)) < tc25 &&
                o_abs(AT(l3, -4) - AT(l3, -1)) + o_abs(AT(l3, 3) - AT(l3, 0)) < (beta >> 3) && o_abs(AT(l3, -1) - AT(l3, 0)) < tc25 &&
                (d0 << 1) < (beta >> 2) && (d3 << 1) < (beta >> 2)) {
                luma_strong(pix, xs, ys, tc);
            } else {
                int np = 1, nq = 1;
                if (lp > 1 && lq > 1) {
                    const int side = (beta + (beta >> 1)) >> 3;
                    if (dp0 + dp3 < side) np = 2;
                    if (dq0 + dq3 < side) nq = 2;
                }
                luma_weak(pix, xs, ys, tc, np, nq, bd);
            }
        }
    }
}

/* one chroma segment of `lines` lines (2 when the edge direction is subsampled, else 4) */
static void chroma_segment(pel *pix, ptrdiff_t xs, ptrdiff_t ys, int tc_in, int beta_in, int lp, int lq,
                           int lines, int bd)
{
    const int tc = tc_in << (bd - 10), beta = beta_in << (bd - 8);
    if (!tc || !lp || !lq)
        return;
    if (lq == 3) {
        const pel *l0 = pix, *l1 = pix + (lines == 2 ? 1 : 3) * ys;   /* second decision line (:714-719) */
        const int tc25 = (tc * 5 + 1) >> 1;
        const int one = lp == 1;
        const int p0 = AT(l0, -1), p1 = AT(l0, -2), p2 = one ? p1 : AT(l0, -3), p3 = one ? p1 : AT(l0, -4);
        const int n0 = AT(l1, -1), n1 = AT(l1, -2), n2 = one ? n1 : AT(l1, -3), n3 = one ? n1 : AT(l1, -4);
        const int d0 = o_abs(p2 - 2 * p1 + p0) + curv(l0, xs, 2, 1, 0);
        const int d1 = o_abs(n2 - 2 * n1 + n0) + curv(l1, xs, 2, 1, 0);
        int strong = 0;
        if (d0 + d1 < beta) {
            const int ok0 = (d0 << 1) < (beta >> 2) && o_abs(p3 - p0) + o_abs(AT(l0, 0) - AT(l0, 3)) < (beta >> 3) && o_abs(p0 - AT(l0, 0)) < tc25;
            const int ok1 = (d1 << 1) < (beta >> 2) && o_abs(n3 - n0) + o_abs(AT(l1, 0) - AT(l1, 3)) < (beta >> 3) && o_abs(n0 - AT(l1, 0)) < tc25;
            strong = ok0 && ok1;
        }
        if (!strong)
            lp = lq = 1;
    }
    for (int line = 0; line < lines; line++, pix += ys) {
        const int p3 = AT(pix, -4), p2 = AT(pix, -3), p1 = AT(pix, -2), p0 = AT(pix, -1);
        const int q0 = AT(pix, 0), q1 = AT(pix, 1), q2 = AT(pix, 2), q3 = AT(pix, 3);
        if (lp == 3 && lq == 3) {
            AT(pix, -1) = (pel)o_clip3((p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
            AT(pix, -2) = (pel)o_clip3((2 * p3 + p2 + 2 * p1 + p0 + q0 + q1 + 4) >> 3, p1 - tc, p1 + tc);
            AT(pix, -3) = (pel)o_clip3((3 * p3 + 2 * p2 + p1 + p0 + q0 + 4) >> 3, p2 - tc, p2 + tc);
            AT(pix, 0)  = (pel)o_clip3((p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
            AT(pix, 1)  = (pel)o_clip3((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
            AT(pix, 2)  = (pel)o_clip3((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
        } else if (lq == 3) {
            AT(pix, -1) = (pel)o_clip3((3 * p1 + 2 * p0 + q0 + q1 + q2 + 4) >> 3, p0 - tc, p0 + tc);
            AT(pix, 0)  = (pel)o_clip3((2 * p1 + p0 + 2 * q0 + q1 + q2 + q3 + 4) >> 3, q0 - tc, q0 + tc);
            AT(pix, 1)  = (pel)o_clip3((p1 + p0 + q0 + 2 * q1 + q2 + 2 * q3 + 4) >> 3, q1 - tc, q1 + tc);
            AT(pix, 2)  = (pel)o_clip3((p0 + q0 + q1 + 2 * q2 + 3 * q3 + 4) >> 3, q2 - tc, q2 + tc);
        } else {
            const int delta = o_clip3((((q0 - p0) * 4) + p1 - q1 + 4) >> 3, -tc, tc);
            AT(pix, -1) = (pel)o_clip_pel(p0 + delta, bd);
            AT(pix, 0)  = (pel)o_clip_pel(q0 - delta, bd);
        }
    }
}

static void copy_frame(const VVCCudaFrame *d, const VVCCudaFrame *s)
{
    const int planes = s->chroma_format_idc ? 3 : 1;
    if (d->data[0] == s->data[0])
        return;
    for (int k = 0; k < s->batch; k++)
        for (int c = 0; c < planes; c++) {
            const OPlane sp = o_plane(s, c, k), dp = o_plane(d, c, k);
            for (int y = 0; y < sp.h; y++)
                memcpy(dp.p + y * dp.pitch, sp.p + y * sp.pitch, sp.w * sizeof(pel));
        }
}

/* dir 1: vertical edges, dir 0: horizontal edges; in place on dst after copying src. */
void vvco_deblock_frame(const VVCCudaFrame *dstf, const VVCCudaFrame *srcf, const VVCCudaDeblockMaps *maps, int dir)
{
    const int ctb = 1 << srcf->ctb_log2, cols = o_ctb_cols(srcf), rows = o_ctb_rows(srcf);
    const int planes = srcf->chroma_format_idc ? 3 : 1, bd = srcf->bit_depth;

    copy_frame(dstf, srcf);
    for (int k = 0; k < srcf->batch; k++)
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++)
                for (int c = 0; c < planes; c++) {
                    const OPlane pl = o_plane(dstf, c, k);
                    const int hs = c ? srcf->hshift : 0, vs = c ? srcf->vshift : 0;
                    const int grid = c ? 8 : 4;
                    const int seg = c ? 4 >> (dir ? vs : hs) : 4;       /* samples of one segment along the edge */
                    const int x0 = (cx * ctb) >> hs, y0 = (cy * ctb) >> vs;
                    const int x1 = o_min(x0 + (ctb >> hs), pl.w), y1 = o_min(y0 + (ctb >> vs), pl.h);
                    const VVCCudaDbkEdge *map = maps->edge[dir][c] + (size_t)k * maps->size[dir][c];
                    const int mp = maps->pitch[dir][c];
                    if (dir) {
                        for (int y = y0; y < y1; y += seg)
                            for (int x = o_max(x0, grid); x < x1; x += grid) {
                                const VVCCudaDbkEdge e = map[(y / seg) * mp + x / grid];
                                pel *pix = pl.p + y * pl.pitch + x;
                                if (c) chroma_segment(pix, 1, pl.pitch, e.tc, e.beta, e.max_len & 15, e.max_len >> 4, seg, bd);
                                else   luma_segment(pix, 1, pl.pitch, e.tc, e.beta, e.max_len & 15, e.max_len >> 4, 0, bd);
                            }
                    } else {
                        for (int y = o_max(y0, grid); y < y1; y += grid)
                            for (int x = x0; x < x1; x += seg) {
                                const VVCCudaDbkEdge e = map[(y / grid) * mp + x / seg];
                                pel *pix = pl.p + y * pl.pitch + x;
                                const int ctu_edge = !((y << vs) & (ctb - 1));
                                if (c) chroma_segment(pix, pl.pitch, 1, e.tc, e.beta, e.max_len & 15, e.max_len >> 4, seg, bd);
                                else   luma_segment(pix, pl.pitch, 1, e.tc, e.beta, e.max_len & 15, e.max_len >> 4, ctu_edge, bd);
                            }
                    }
                }
}
