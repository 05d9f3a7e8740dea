/*
 * Oracle: derivation of the deblocking parameters of a picture (TEST INFRASTRUCTURE, see vvc_oracle.h).
 *
 * Restates, on the list inputs of vvc_cuda_deblock_params_frame (include/vvcdsp_cuda.h):
 *   boundary_strength :308-370, derive_max_filter_length_luma :373-397, vvc_deblock_subblock_bs_vertical / _horizontal
 *   :399-470, vvc_deblock_bs_luma_vertical / _horizontal :472-634, vvc_deblock_bs_chroma_vertical / _horizontal :636-754,
 *   vvc_deblock_bs :756-781, max_filter_length_chroma :793-812, TC_CALC :823-826, get_qp_y / get_qp_c :829-852 and the
 *   per-edge part of ff_vvc_deblock_vertical / _horizontal :861-1003           libavcodec/vvc/vvc_filter.c
 *   vvc_loop_ladf_level :788-804                                               libavcodec/vvc/vvc_filter_template.c
 * The lists are first scattered into per-4x4 tables (what the reference's parser keeps in fc->tab.*), then every
 * transform block writes the strengths of its left / upper edge and of its sub-block edges, then every edge segment gets
 * (tc, beta, maximum lengths) as the reference's filter loop computes them per call.
 */
#include "vvc_oracle.h"

typedef struct Unit {
    int16_t tbx[2], tby[2];         /* transform block origin per tree (luma samples), -1 = none */
    uint8_t tbw[2], tbh[2];         /* its luma size (log2) */
    uint8_t cbf[3], joint, pcm[2];
    int8_t  qp[3];
    int16_t cbx, cby;
    uint8_t cbw, cbh, sb;           /* coding block size (log2), merge-subblock / affine */
    uint8_t pred, ciip;
    int16_t ref[2];
    int32_t mv[2][2];
} Unit;

static const uint16_t tc_table[66] = {
      0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
      0,   0,   3,   4,   4,   4,   4,   5,   5,   5,   5,   7,   7,   8,   9,  10,
     10,  11,  13,  14,  15,  17,  19,  21,  24,  25,  29,  33,  36,  41,  45,  51,
     57,  64,  71,  80,  89, 100, 112, 125, 141, 157, 177, 198, 222, 250, 280, 314,
    352, 395,
};
static const uint8_t beta_table[64] = {
      0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,   0,
      6,   7,   8,   9,  10,  11,  12,  13,  14,  15,  16,  17,  18,  20,  22,  24,
     26,  28,  30,  32,  34,  36,  38,  40,  42,  44,  46,  48,  50,  52,  54,  56,
     58,  60,  62,  64,  66,  68,  70,  72,  74,  76,  78,  80,  82,  84,  86,  88,
};

static int far(const int32_t a[2], const int32_t b[2]) { return o_abs(a[0] - b[0]) >= 8 || o_abs(a[1] - b[1]) >= 8; }

/* boundary_strength: motion discontinuity between two inter units */
static int motion_bs(const Unit *c, const Unit *n)
{
    if (c->pred == 3 && n->pred == 3) {
        if (c->ref[0] == n->ref[0] && c->ref[0] == c->ref[1] && n->ref[0] == n->ref[1])
            return (far(n->mv[0], c->mv[0]) || far(n->mv[1], c->mv[1])) && (far(n->mv[1], c->mv[0]) || far(n->mv[0], c->mv[1]));
        if (n->ref[0] == c->ref[0] && n->ref[1] == c->ref[1])
            return far(n->mv[0], c->mv[0]) || far(n->mv[1], c->mv[1]);
        if (n->ref[1] == c->ref[0] && n->ref[0] == c->ref[1])
            return far(n->mv[1], c->mv[0]) || far(n->mv[0], c->mv[1]);
        return 1;
    }
    if (c->pred != 3 && n->pred != 3) {
        const int lc = (c->pred & 1) ? 0 : 1, ln = (n->pred & 1) ? 0 : 1;
        return c->ref[lc] == n->ref[ln] ? far(c->mv[lc], n->mv[ln]) : 1;
    }
    return 1;
}

typedef struct Tabs {
    Unit    *u;
    uint8_t *bs[3];         /* strength of the edge on the left (vertical pass) / upper (horizontal pass) side of a unit */
    uint8_t *lp, *lq;       /* luma maximum filter lengths */
    int      uw, uh;
} Tabs;

#define U(t, x, y) (&(t)->u[((y) >> 2) * (t)->uw + ((x) >> 2)])
#define IDX(t, x, y) (((y) >> 2) * (t)->uw + ((x) >> 2))

static void luma_block(Tabs *t, const VVCCudaFrame *f, const VVCCudaDbkCtb *ctbs, int x0, int y0, int w, int h, int vertical)
{
    const Unit *q0 = U(t, x0, y0);
    const int is_intra = q0->pred == 0;
    const int cb_side = vertical ? 1 << q0->cbw : 1 << q0->cbh;
    const int has_sb = !is_intra && q0->sb && cb_side > 8;
    const int ctb_mask = (1 << f->ctb_log2) - 1;
    const int pos = vertical ? x0 : y0, len = vertical ? h : w;
    int edge = pos > 0;
    if (edge && !(pos & ctb_mask)) {
        const VVCCudaDbkCtb *cb = &ctbs[(y0 >> f->ctb_log2) * o_ctb_cols(f) + (x0 >> f->ctb_log2)];
        if (vertical ? cb->no_left : cb->no_top)
            edge = 0;
    }
    if (edge) {
        const int off = (vertical ? q0->cbx : q0->cby) - pos;          /* <= 0: distance to the coding block's own edge */
        for (int i = 0; i < len; i += 4) {
            const int qx = vertical ? x0 : x0 + i, qy = vertical ? y0 + i : y0;
            const int px = vertical ? qx - 1 : qx, py = vertical ? qy : qy - 1;
            const Unit *q = U(t, qx, qy), *p = U(t, px, py);
            int bs, lp, lq;
            if (p->pcm[0] && q->pcm[0])                                         bs = 0;
            else if (q->pred == 0 || p->pred == 0 || q->ciip || p->ciip)       bs = 2;
            else if (q->cbf[0] || p->cbf[0])                                   bs = 1;
            else if (off && ((off % 8) || !has_sb))                            bs = 0;
            else                                                               bs = motion_bs(q, p);
            t->bs[0][IDX(t, qx, qy)] = (uint8_t)bs;
            const int size_p = 1 << (vertical ? p->tbw[0] : p->tbh[0]), size_q = 1 << (vertical ? q->tbw[0] : q->tbh[0]);
            if (size_p <= 4 || size_q <= 4) {
                lp = lq = 1;
            } else {
                lp = size_p >= 32 ? 7 : 3;
                lq = size_q >= 32 ? 7 : 3;
            }
            if (has_sb) lq = o_min(5, lq);
            if (p->sb)  lp = o_min(5, lp);
            t->lp[IDX(t, qx, qy)] = (uint8_t)lp; t->lq[IDX(t, qx, qy)] = (uint8_t)lq;
        }
    }
    if (!is_intra && q0->sb) {
        /* edges of the 8x8 sub-block grid of the coding block that fall inside this transform block */
        const int cb0 = vertical ? q0->cbx : q0->cby, across = vertical ? w : h;
        for (int j = 0; j < len; j += 4)
            for (int i = 8 - ((pos - cb0) % 8); i < across; i += 8) {
                const int qx = vertical ? x0 + i : x0 + j, qy = vertical ? y0 + j : y0 + i;
                const Unit *q = U(t, qx, qy), *p = U(t, vertical ? qx - 1 : qx, vertical ? qy : qy - 1);
                const int ml = (i == 4 || i == across - 4) ? 1 : (i == 8 || i == across - 8) ? 2 : 3;
                t->bs[0][IDX(t, qx, qy)] = (uint8_t)motion_bs(q, p);
                t->lp[IDX(t, qx, qy)] = t->lq[IDX(t, qx, qy)] = (uint8_t)ml;
            }
    }
}

static void chroma_block(Tabs *t, const VVCCudaFrame *f, const VVCCudaDbkCtb *ctbs, int x0, int y0, int w, int h, int vertical)
{
    const int ctb_mask = (1 << f->ctb_log2) - 1;
    const int pos = vertical ? x0 : y0, len = vertical ? h : w;
    const int grid = 8 << (vertical ? f->hshift : f->vshift);
    int edge = pos > 0 && !(pos & (grid - 1));
    if (edge && !(pos & ctb_mask)) {
        const VVCCudaDbkCtb *cb = &ctbs[(y0 >> f->ctb_log2) * o_ctb_cols(f) + (x0 >> f->ctb_log2)];
        if (vertical ? cb->no_left : cb->no_top)
            edge = 0;
    }
    if (!edge)
        return;
    for (int i = 0; i < len; i += 4) {
        const int qx = vertical ? x0 : x0 + i, qy = vertical ? y0 + i : y0;
        const Unit *q = U(t, qx, qy), *p = U(t, vertical ? qx - 1 : qx, vertical ? qy : qy - 1);
        for (int c = 1; c <= 2; c++) {
            int bs = 0;
            if (p->pcm[1] && q->pcm[1])                                        bs = 0;
            else if (q->pred == 0 || p->pred == 0 || q->ciip || p->ciip)       bs = 2;
            else if (p->cbf[c] | q->cbf[c] | p->joint | q->joint)              bs = 1;
            t->bs[c][IDX(t, qx, qy)] = (uint8_t)bs;
        }
    }
}

void vvco_deblock_params_frame(const VVCCudaFrame *f, const VVCCudaDbkTU *tus, int n_tus, const VVCCudaDbkMvf *mvfs, int n_mvfs,
                               const VVCCudaDbkCtb *ctbs, const VVCCudaDbkParams *prm, const VVCCudaDeblockMaps *maps, int dir)
{
    const int vertical = dir, planes = f->chroma_format_idc ? 3 : 1;
    Tabs t;
    t.uw = (f->width + 3) >> 2; t.uh = (f->height + 3) >> 2;
    const size_t nu = (size_t)t.uw * t.uh;
    t.u = malloc(nu * sizeof(Unit));
    for (int c = 0; c < 3; c++) t.bs[c] = malloc(nu);
    t.lp = malloc(nu); t.lq = malloc(nu);
    for (int k = 0; k < f->batch; k++) {
        memset(t.u, 0, nu * sizeof(Unit));
        for (size_t i = 0; i < nu; i++)
            t.u[i].tbx[0] = t.u[i].tbx[1] = t.u[i].tby[0] = t.u[i].tby[1] = -1;
        for (int c = 0; c < 3; c++) memset(t.bs[c], 0, nu);
        memset(t.lp, 0, nu); memset(t.lq, 0, nu);
        /* ---- the parser's side tables ---- */
        for (int i = 0; i < n_tus; i++) {
            const VVCCudaDbkTU *tu = &tus[i];
            if (tu->pic != k) continue;
            const int w = 1 << tu->log2_w, h = 1 << tu->log2_h;
            for (int y = tu->y0; y < tu->y0 + h && y < f->height; y += 4)
                for (int x = tu->x0; x < tu->x0 + w && x < f->width; x += 4) {
                    Unit *u = U(&t, x, y);
                    if (tu->planes & VVC_CUDA_DBK_TU_LUMA) {
                        u->tbx[0] = (int16_t)tu->x0; u->tby[0] = (int16_t)tu->y0; u->tbw[0] = tu->log2_w; u->tbh[0] = tu->log2_h;
                        u->cbf[0] = !!(tu->flags & VVC_CUDA_DBK_CBF_Y); u->pcm[0] = !!(tu->flags & VVC_CUDA_DBK_BDPCM_Y);
                        u->qp[0] = tu->qp[0];
                        u->cbx = (int16_t)(tu->x0 - 4 * tu->cu_dx); u->cby = (int16_t)(tu->y0 - 4 * tu->cu_dy);
                        u->cbw = tu->cb_log2_w; u->cbh = tu->cb_log2_h; u->sb = !!(tu->cu_flags & VVC_CUDA_DBK_CU_SUBBLOCK);
                    }
                    if (tu->planes & VVC_CUDA_DBK_TU_CHROMA) {
                        u->tbx[1] = (int16_t)tu->x0; u->tby[1] = (int16_t)tu->y0; u->tbw[1] = tu->log2_w; u->tbh[1] = tu->log2_h;
                        u->cbf[1] = !!(tu->flags & VVC_CUDA_DBK_CBF_CB); u->cbf[2] = !!(tu->flags & VVC_CUDA_DBK_CBF_CR);
                        u->joint = !!(tu->flags & VVC_CUDA_DBK_JOINT); u->pcm[1] = !!(tu->flags & VVC_CUDA_DBK_BDPCM_C);
                        u->qp[1] = tu->qp[1]; u->qp[2] = tu->qp[2];
                    }
                }
        }
        for (int i = 0; i < n_mvfs; i++) {
            const VVCCudaDbkMvf *m = &mvfs[i];
            if (m->pic != k) continue;
            for (int y = m->y0; y < m->y0 + 4 * m->h4 && y < f->height; y += 4)
                for (int x = m->x0; x < m->x0 + 4 * m->w4 && x < f->width; x += 4) {
                    Unit *u = U(&t, x, y);
                    u->pred = m->pred_flag; u->ciip = m->ciip_flag;
                    u->ref[0] = m->ref_pic[0]; u->ref[1] = m->ref_pic[1];
                    memcpy(u->mv, m->mv, sizeof(u->mv));
                }
        }
        /* ---- strengths and luma lengths, one transform block at a time (vvc_deblock_bs) ---- */
        const VVCCudaDbkCtb *cb = ctbs + (size_t)k * o_ctb_cols(f) * o_ctb_rows(f);
        for (int y = 0; y < f->height; y += 4)
            for (int x = 0; x < f->width; x += 4) {
                const Unit *u = U(&t, x, y);
                if (u->tbx[0] == x && u->tby[0] == y)
                    luma_block(&t, f, cb, x, y, 1 << u->tbw[0], 1 << u->tbh[0], vertical);
                if (planes == 3 && u->tbx[1] == x && u->tby[1] == y)
                    chroma_block(&t, f, cb, x, y, 1 << u->tbw[1], 1 << u->tbh[1], vertical);
            }
        /* ---- per edge segment: QP, beta, tc, lengths (the filter loops of ff_vvc_deblock_vertical / _horizontal) ---- */
        const OPlane Y = o_plane(f, 0, k);
        for (int c = 0; c < planes; c++) {
            VVCCudaDbkEdge *map = (VVCCudaDbkEdge *)maps->edge[dir][c] + (size_t)k * maps->size[dir][c];
            const int pitch = maps->pitch[dir][c];
            memset(map, 0, (size_t)maps->size[dir][c] * sizeof(*map));
            const int grid = c ? 8 << (vertical ? f->hshift : f->vshift) : 4;
            for (int y = 0; y < f->height; y += 4)
                for (int x = 0; x < f->width; x += 4) {
                    const int pos = vertical ? x : y;
                    if (!pos || (pos & (grid - 1)))
                        continue;
                    const int bs = t.bs[c][IDX(&t, x, y)];
                    if (!bs)
                        continue;
                    const Unit *q = U(&t, x, y), *p = U(&t, vertical ? x - 1 : x, vertical ? y : y - 1);
                    const VVCCudaDbkCtb *ct = &cb[(y >> f->ctb_log2) * o_ctb_cols(f) + (x >> f->ctb_log2)];
                    int qp, lp, lq;
                    if (!c) {
                        qp = (p->qp[0] + q->qp[0] + 1) >> 1;
                        if (prm->ladf_enabled) {
                            const pel *s = Y.p + (ptrdiff_t)y * Y.pitch + x;
                            const ptrdiff_t xs = vertical ? 1 : Y.pitch, ys = vertical ? Y.pitch : 1;
                            const int level = (s[-xs] + s[-xs + 3 * ys] + s[0] + s[3 * ys]) >> 2;
                            int off = prm->ladf_lowest_interval_qp_offset;
                            for (int i = 0; i < prm->num_ladf_intervals - 1 && level > prm->ladf_interval_lower_bound[i + 1]; i++)
                                off = prm->ladf_qp_offset[i];
                            qp += off;
                        }
                        lp = t.lp[IDX(&t, x, y)]; lq = t.lq[IDX(&t, x, y)];
                    } else {
                        qp = (p->qp[c] + q->qp[c] - 2 * prm->qp_bd_offset + 1) >> 1;
                        /* max_filter_length_chroma: transform block sizes in chroma samples across the edge */
                        const int size_p = (1 << (vertical ? p->tbw[1] : p->tbh[1])) >> (vertical ? f->hshift : f->vshift);
                        const int size_q = (1 << (vertical ? q->tbw[1] : q->tbh[1])) >> (vertical ? f->hshift : f->vshift);
                        if (size_p >= 8 && size_q >= 8) {
                            lp = lq = 3;
                            if (!vertical && !(y & ((1 << f->ctb_log2) - 1)))
                                lp = 1;
                        } else {
                            lp = lq = bs == 2;
                        }
                    }
                    VVCCudaDbkEdge e;
                    e.beta = beta_table[o_clip3(qp + ct->beta_offset[c], 0, 63)];
                    e.tc = tc_table[o_clip3(qp + 2 * (bs - 1) + (ct->tc_offset[c] & -2), 0, 63 + 2)];
                    e.max_len = (uint8_t)(lp | (lq << 4));
                    const int hs = c ? f->hshift : 0, vs = c ? f->vshift : 0;
                    const int cx = x >> hs, cy = y >> vs;                              /* position in plane c */
                    const int gs = c ? 8 : 4, seg = c ? 4 >> (vertical ? vs : hs) : 4;
                    if (vertical) map[(size_t)(cy / seg) * pitch + cx / gs] = e;
                    else          map[(size_t)(cy / gs) * pitch + cx / seg] = e;
                }
        }
    }
    free(t.u); for (int c = 0; c < 3; c++) free(t.bs[c]);
    free(t.lp); free(t.lq);
}
