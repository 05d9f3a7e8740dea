/*
 * intra.c - CPU oracle of the intra leaf predictors and the CIIP blend (TEST INFRASTRUCTURE, see
 * vvc_oracle.h).  Restates, against VVCCudaIntraPB / VVCCudaCiip of include/vvcdsp_cuda.h:
 *   pred_planar :686, pred_dc / pred_dc_val :826-864, pred_v :866, pred_h :877,
 *   pred_angular_v :894, pred_angular_h :950, pred_mip :773 (mip_downsampling :708, mip_reduced_pred :728,
 *   mip_upsampling_1d :749)                                   libavcodec/vvc/vvc_intra_template.c
 *   ff_vvc_intra_pred_angle_derive :661, ff_vvc_intra_inv_angle_derive :684, ff_vvc_nscale_derive :538,
 *   ff_vvc_get_mip_size_id :529                               libavcodec/vvc/vvc_intra.c
 *   put_ciip                                                  libavcodec/vvc/vvc_inter_template.c:60-76
 * Written per output sample (what each sample depends on) rather than as in-place passes.
 */
#include "vvc_oracle.h"
#include "vvc_tables_c.h"

static int pred_angle(int mode)
{
    static const int angles[] = { 0, 1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 23, 26, 29,
                                  32, 35, 39, 45, 51, 57, 64, 73, 86, 102, 128, 171, 256, 341, 512 };
    int idx = mode > 34 ? mode - 50 : mode > 0 ? 18 - mode : 18 - 2 - mode;      /* INTRA_DIAG 34, VERT 50, HORZ 18 */
    int sign = 1;
    if (idx < 0) { idx = -idx; sign = -1; }
    return sign * angles[idx];
}

static int inv_angle_of(int angle)
{
    const float f = 32 * 512.0 / angle;           /* the reference's float expression, vvc_intra.c:683-690 */
    return (int)(f < 0 ? -(-f + 0.5) : (f + 0.5));
}

static int nscale_of(int w, int h, int mode)
{
    const int side = mode >= 50 ? h : w;
    return o_min(2, o_ilog2(side) - o_ilog2(3 * inv_angle_of(pred_angle(mode)) - 2) + 8);
}

static void angular(pel *dst, ptrdiff_t pitch, const pel *top0, const pel *left0, const VVCCudaIntraPB *b, int bd, int vertical)
{
    const int w = b->w, h = b->h, ref_idx = b->ref_idx, is_luma = !b->c_idx, pdpc = b->flags & VVC_CUDA_INTRA_PDPC;
    const int angle = pred_angle(b->mode);
    /* main reference: top for the vertical family, left for the horizontal one; roles of x and y swap */
    const pel *mainr = (vertical ? top0 : left0) - (1 + ref_idx), *side = vertical ? left0 : top0;
    const int inv_angle = pdpc ? inv_angle_of(angle) : 0, nscale = pdpc ? nscale_of(w, h, b->mode) : 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int along = vertical ? x : y, across = vertical ? y : x;   /* position along / across the main reference */
            const int pos = (across + 1 + ref_idx) * angle, idx = (pos >> 5) + ref_idx, fact = pos & 31;
            const pel *p = mainr + along + idx;
            int pred;
            if (!fact && (!is_luma || !b->filter_flag)) {
                pred = p[1];
            } else if (is_luma) {
                const int8_t *f = vvct_intra_luma_filter[b->filter_flag][fact];
                pred = o_clip_pel((p[0] * f[0] + p[1] * f[1] + p[2] * f[2] + p[3] * f[3] + 32) >> 6, bd);
            } else {
                pred = ((32 - fact) * p[1] + fact * p[2] + 16) >> 5;
            }
            if (pdpc) {
                if (vertical) {
                    if (x < o_min(w, 3 << nscale)) {
                        const int l = side[y + ((256 + (x + 1) * inv_angle) >> 9)], wl = 32 >> ((x << 1) >> nscale);
                        pred = o_clip_pel(pred + (((l - pred) * wl + 32) >> 6), bd);
                    }
                } else if (y < (3 << nscale)) {
                    const int t = side[x + ((256 + (y + 1) * inv_angle) >> 9)], wt = 32 >> o_min(31, (y * 2) >> nscale);
                    pred = o_clip_pel(pred + (((t - pred) * wt + 32) >> 6), bd);
                }
            }
            dst[y * pitch + x] = (pel)pred;
        }
}

static void mip(pel *dst, ptrdiff_t pitch, const pel *top, const pel *left, const VVCCudaIntraPB *b, int bd)
{
    const int w = b->w, h = b->h, transposed = !!(b->flags & VVC_CUDA_INTRA_MIP_TRANSPOSED);
    const int size_id = (w == 4 && h == 4) ? 0 : ((w == 4 || h == 4) || (w == 8 && h == 8)) ? 1 : 2;
    const int bsz = size_id ? 4 : 2, psz = size_id == 2 ? 8 : 4, in_size = 2 * bsz - (size_id == 2);
    const uint8_t *matrix = size_id == 0 ? &vvct_mip_matrix_0[b->mode][0][0] : size_id == 1 ? &vvct_mip_matrix_1[b->mode][0][0]
                                                                                              : &vvct_mip_matrix_2[b->mode][0][0];
    const int up_hor = w / psz, up_ver = h / psz;
    int reduced[16], *red_t = transposed ? reduced + bsz : reduced, *red_l = transposed ? reduced : reduced + bsz;
    int small[8][8];
    for (int side = 0; side < 2; side++) {               /* boundary down-sampling */
        const pel *ref = side ? left : top;
        int *out = side ? red_l : red_t;
        const int n = side ? h : w, dwn = n / bsz, lg = o_ilog2(dwn);
        for (int i = 0; i < bsz; i++) {
            int r = 0;
            for (int j = 0; j < dwn; j++)
                r += ref[i * dwn + j];
            out[i] = dwn == 1 ? r : (r + (1 << (lg - 1))) >> lg;
        }
    }
    {
        const int temp0 = reduced[0], off = size_id != 2 ? 0 : 1;
        int ow = size_id != 2 ? (1 << (bd - 1)) - temp0 : reduced[1] - temp0;
        reduced[0] = ow;
        for (int i = 1; i < in_size; i++) {
            reduced[i] = reduced[i + off] - temp0;
            ow += reduced[i];
        }
        ow = 32 - 32 * ow;
        for (int y = 0; y < psz; y++)
            for (int x = 0; x < psz; x++) {
                int pred = 0;
                for (int i = 0; i < in_size; i++)
                    pred += reduced[i] * matrix[(y * psz + x) * in_size + i];
                pred = o_clip3(((pred + ow) >> 6) + temp0, 0, (1 << bd) - 1);
                if (transposed) small[x][y] = pred; else small[y][x] = pred;
            }
    }
    /* reduced prediction sits at ((i + 1) * up - 1); horizontal then vertical linear up-sampling from the boundaries */
    for (int j = 0; j < psz; j++)
        for (int i = 0; i < psz; i++)
            dst[((j + 1) * up_ver - 1) * pitch + (i + 1) * up_hor - 1] = (pel)small[j][i];
    if (up_hor > 1)
        for (int j = 0; j < psz; j++) {
            pel *row = dst + ((j + 1) * up_ver - 1) * pitch;
            int before = left[(j + 1) * up_ver - 1];
            for (int i = 0; i < psz; i++) {
                const int after = row[(i + 1) * up_hor - 1];
                for (int k = 1; k < up_hor; k++)
                    row[i * up_hor + k - 1] = (pel)(((up_hor - k) * before + k * after + up_hor / 2) / up_hor);
                before = after;
            }
        }
    if (up_ver > 1)
        for (int x = 0; x < w; x++) {
            int before = top[x];
            for (int j = 0; j < psz; j++) {
                const int after = dst[((j + 1) * up_ver - 1) * pitch + x];
                for (int k = 1; k < up_ver; k++)
                    dst[(j * up_ver + k - 1) * pitch + x] = (pel)(((up_ver - k) * before + k * after + up_ver / 2) / up_ver);
                before = after;
            }
        }
}

/* one leaf predictor: dst = the block's first sample, top / left as IntraEdgeParams carries them */
static void leaf_predict(pel *dst, ptrdiff_t pitch, const pel *top, const pel *left, const VVCCudaIntraPB *b, int bd)
{
    const int w = b->w, h = b->h;
    switch (b->kind) {
    case VVC_CUDA_INTRA_PLANAR: {
        const int lw = o_ilog2(w), lh = o_ilog2(h);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int pv = ((h - 1 - y) * top[x] + (y + 1) * left[h]) << lw;
                const int ph = ((w - 1 - x) * left[y] + (x + 1) * top[w]) << lh;
                dst[y * pitch + x] = (pel)((pv + ph + w * h) >> (lw + lh + 1));
            }
        break;
    }
    case VVC_CUDA_INTRA_DC: {
        unsigned offset = w == h ? (unsigned)w << 1 : (unsigned)o_max(w, h);
        const int shift = o_ilog2(offset);
        int sum = 0;
        if (w >= h) for (int k = 0; k < w; k++) sum += top[k];
        if (w <= h) for (int k = 0; k < h; k++) sum += left[k];
        const pel dc = (pel)((sum + (offset >> 1)) >> shift);
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                dst[y * pitch + x] = dc;
        break;
    }
    case VVC_CUDA_INTRA_VERT:
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                dst[y * pitch + x] = top[x];
        break;
    case VVC_CUDA_INTRA_HORZ:
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
                dst[y * pitch + x] = left[y];
        break;
    case VVC_CUDA_INTRA_ANGULAR_V: angular(dst, pitch, top, left, b, bd, 1); break;
    case VVC_CUDA_INTRA_ANGULAR_H: angular(dst, pitch, top, left, b, bd, 0); break;
    default:                       mip(dst, pitch, top, left, b, bd); break;
    }
}

void vvco_intra_leaf_frame(const VVCCudaFrame *f, const VVCCudaIntraPB *pbs, int n, const uint16_t *edges)
{
    for (int i = 0; i < n; i++) {
        const VVCCudaIntraPB *b = &pbs[i];
        const OPlane pl = o_plane(f, b->c_idx, b->pic);
        leaf_predict(pl.p + b->y0 * pl.pitch + b->x0, pl.pitch, edges + b->top, edges + b->left, b, f->bit_depth);
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * intra.intra_pred with its edge preparation, and intra.intra_cclm_pred, on VVCCudaIntraBlk records.
 *   intra_pred :595-683, prepare_intra_edge_params :467-592, ref_filter :450-465, intra_cclm_pred :29-375
 *                                                                             libavcodec/vvc/vvc_intra_template.c
 *   ff_vvc_wide_angle_mode_mapping :693-715, ff_vvc_need_pdpc :557-573, ff_vvc_ref_filter_flag_derive :655-659
 *                                                                             libavcodec/vvc/vvc_intra.c
 * ---------------------------------------------------------------------------------------------------------------- */
enum { E_NEG = 80, E_LEN = E_NEG + 192 };      /* reference lines: indices -(64 + 3) .. 2 * 64 + 16 * 2 + 1 */

static int wide_angle_mode(const VVCCudaIntraBlk *b)
{
    const int isp = (b->flags & VVC_CUDA_INTRA_F_ISP) && !b->c_idx;
    const int nw = isp ? b->cb_w : b->w, nh = isp ? b->cb_h : b->h;
    const int ratio = o_abs(o_ilog2(nw) - o_ilog2(nh));
    const int hi = ratio > 1 ? 8 + 2 * ratio : 8, lo = ratio > 1 ? 60 - 2 * ratio : 60;
    int m = b->pred_mode;
    if (nw > nh && m >= 2 && m < hi)
        m += 65;
    else if (nh > nw && m <= 66 && m > lo)
        m -= 67;
    return m;
}

static int smoothing_mode(int mode)
{
    static const int modes[] = { -14, -12, -10, -6, 0, 2, 34, 66, 72, 76, 78, 80 };
    for (unsigned i = 0; i < sizeof(modes) / sizeof(modes[0]); i++)
        if (modes[i] == mode)
            return 1;
    return 0;
}

static int pdpc_needed(int w, int h, int bdpcm, int mode, int ref_idx)
{
    if (w < 4 || h < 4 || ref_idx || bdpcm)
        return 0;
    if (mode == 0 || mode == 1 || mode == 18 || mode == 50)
        return 1;
    if (mode > 18 && mode < 50)
        return 0;
    return nscale_of(w, h, mode) >= 0;
}

static void intra_pred_block(const VVCCudaFrame *f, const VVCCudaIntraBlk *b)
{
    const int bd = f->bit_depth, c_idx = b->c_idx, x0 = b->x0, y0 = b->y0, w = b->w, h = b->h;
    const OPlane pl = o_plane(f, c_idx, b->pic);
    const int is_mip = b->kind == VVC_CUDA_INTRA_KIND_MIP;
    const int isp = (b->flags & VVC_CUDA_INTRA_F_ISP) != 0, ref_idx = c_idx ? 0 : b->ref_idx, rl = -1 - ref_idx;
    const int mode = is_mip ? 0 : wide_angle_mode(b);
    const int pdpc = !is_mip && pdpc_needed(w, h, (b->flags & VVC_CUDA_INTRA_F_BDPCM) != 0, mode, ref_idx);
    const int rff = is_mip ? 0 : smoothing_mode(mode);
    const int smooth = rff && !ref_idx && w * h > 32 && !c_idx && !isp;
    pel raw_l[E_LEN], raw_t[E_LEN], flt_l[E_LEN], flt_t[E_LEN];
    pel *left = raw_l + E_NEG, *top = raw_t + E_NEG;
    int n_left, n_top, refw = 0, refh = 0, angle = 0;
    memset(raw_l, 0, sizeof(raw_l)); memset(raw_t, 0, sizeof(raw_t));
    memset(flt_l, 0, sizeof(flt_l)); memset(flt_t, 0, sizeof(flt_t));
    /* how many samples of each line the predictor (and the smoothing filter, one more) reads */
    if (is_mip || mode == 0)      { n_left = h + 1 + smooth; n_top = w + 1 + smooth; }
    else if (mode == 1)           { n_left = h; n_top = w; }
    else if (mode == 50)          { n_left = pdpc ? h : 1; n_top = w; }
    else if (mode == 18)          { n_left = h; n_top = pdpc ? w : 1; }
    else {
        refw = (isp && !c_idx) ? b->cb_w + w : 2 * w;
        refh = (isp && !c_idx) ? b->cb_h + h : 2 * h;
        n_top = refw; n_left = refh;
        angle = pred_angle(mode);
    }
#define PIC(xx, yy) pl.p[(ptrdiff_t)(y0 + (yy)) * pl.pitch + x0 + (xx)]
    const int got_l = o_min(n_left, b->avail_left), got_t = o_min(n_top, b->avail_top);
    for (int i = 0; i < got_l; i++) left[i] = PIC(rl, i);
    for (int i = 0; i < got_t; i++) top[i] = PIC(i, rl);
    for (int i = -1; i >= rl; i--) {
        if (b->flags & VVC_CUDA_INTRA_F_UP_LEFT) { left[i] = PIC(rl, i); top[i] = PIC(i, rl); }
        else if (got_l)                          left[i] = top[i] = left[0];
        else if (got_t)                          left[i] = top[i] = top[0];
        else                                     left[i] = top[i] = (pel)(1 << (bd - 1));
    }
#undef PIC
    for (int i = got_t; i < n_top; i++)  top[i] = top[got_t - 1];
    for (int i = got_l; i < n_left; i++) left[i] = left[got_l - 1];
    if (smooth) {
        /* [1 2 1] over both lines; the angular modes keep their last sample, planar reads one sample more instead */
        const int keep_last = !(is_mip || mode == 0);
        pel *fl = flt_l + E_NEG, *ft = flt_t + E_NEG;
        fl[-1] = ft[-1] = (pel)((left[0] + 2 * left[-1] + top[0] + 2) >> 2);
        for (int i = 0; i < n_left - keep_last; i++) fl[i] = (pel)((left[i - 1] + 2 * left[i] + left[i + 1] + 2) >> 2);
        for (int i = 0; i < n_top - keep_last; i++)  ft[i] = (pel)((top[i - 1] + 2 * top[i] + top[i + 1] + 2) >> 2);
        if (keep_last) { ft[n_top - 1] = top[n_top - 1]; fl[n_left - 1] = left[n_left - 1]; }
        left = fl; top = ft;
    }
    VVCCudaIntraPB leaf;
    memset(&leaf, 0, sizeof(leaf));
    leaf.w = (uint8_t)w; leaf.h = (uint8_t)h; leaf.c_idx = (uint8_t)c_idx; leaf.ref_idx = (uint8_t)ref_idx;
    if (!is_mip && mode != 0 && mode != 1) {
        if (!c_idx && !(rff || ref_idx || isp)) {
            static const int thres[] = { 24, 14, 2, 0, 0 };
            const int dist = o_min(o_abs(mode - 50), o_abs(mode - 18)), ntbs = (o_ilog2(w) + o_ilog2(h)) >> 1;
            leaf.filter_flag = dist > thres[ntbs - 2];
        }
        if (mode != 50 && mode != 18) {
            /* the part of the main reference beyond the samples fetched: projected from the other line for negative
             * angles, the last sample repeated for positive ones */
            const int vertical = mode >= 34;
            pel *mainr = vertical ? top : left;
            const pel *side = vertical ? left : top;
            const int n_main = vertical ? refw : refh, across = vertical ? h : w, along = vertical ? w : h;
            if (angle < 0) {
                const int inv = inv_angle_of(angle);
                for (int k = -across; k < 0; k++)
                    mainr[k - (ref_idx + 1)] = side[-1 - ref_idx + o_min((k * inv + 256) >> 9, across)];
            } else {
                for (int i = n_main; i <= n_main + o_max(1, along / across) * ref_idx + 1; i++)
                    mainr[i] = mainr[n_main - 1];
            }
        }
    }
    pel *dst = pl.p + (ptrdiff_t)y0 * pl.pitch + x0;
    if (is_mip) {
        leaf.kind = VVC_CUDA_INTRA_MIP; leaf.mode = (int8_t)b->pred_mode;
        leaf.flags = (b->flags & VVC_CUDA_INTRA_F_MIP_TRANSP) ? VVC_CUDA_INTRA_MIP_TRANSPOSED : 0;
    } else {
        leaf.kind = mode == 0 ? VVC_CUDA_INTRA_PLANAR : mode == 1 ? VVC_CUDA_INTRA_DC : mode == 50 ? VVC_CUDA_INTRA_VERT
                  : mode == 18 ? VVC_CUDA_INTRA_HORZ : mode >= 34 ? VVC_CUDA_INTRA_ANGULAR_V : VVC_CUDA_INTRA_ANGULAR_H;
        leaf.mode = (int8_t)mode;
        leaf.flags = pdpc ? VVC_CUDA_INTRA_PDPC : 0;
    }
    leaf_predict(dst, pl.pitch, top, left, &leaf, bd);
    if (pdpc && (mode == 0 || mode == 1 || mode == 18 || mode == 50)) {
        /* position-dependent filtering of the four non-angular modes (:653-681) */
        const int scale = (o_ilog2(w) + o_ilog2(h) - 2) >> 2;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++) {
                const int val = dst[y * pl.pitch + x];
                int l, t, wl, wt;
                if (mode == 0 || mode == 1) {
                    l = left[y]; t = top[x];
                    wl = 32 >> o_min((x << 1) >> scale, 31);
                    wt = 32 >> o_min((y << 1) >> scale, 31);
                } else {
                    l = left[y] - left[-1] + val; t = top[x] - top[-1] + val;
                    wl = mode == 50 ? 32 >> o_min((x << 1) >> scale, 31) : 0;
                    wt = mode == 18 ? 32 >> o_min((y << 1) >> scale, 31) : 0;
                }
                dst[y * pl.pitch + x] = (pel)o_clip_pel(val + ((wl * (l - val) + wt * (t - val) + 32) >> 6), bd);
            }
    }
}

static void cclm_block(const VVCCudaFrame *f, const VVCCudaIntraBlk *b)
{
    const int bd = f->bit_depth, hs = f->hshift, vs = f->vshift, w = b->w, h = b->h, x = b->x0, y = b->y0;
    const int x0 = x << hs, y0 = y << vs;                         /* luma position */
    const int at = !!(b->flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_T), al = !!(b->flags & VVC_CUDA_INTRA_F_LUMA_AVAIL_L);
    const int colloc = !!(b->flags & VVC_CUDA_INTRA_F_COLLOCATED);
    const OPlane Y = o_plane(f, 0, b->pic);
    OPlane C[2] = { o_plane(f, 1, b->pic), o_plane(f, 2, b->pic) };
    int a[2] = { 0, 0 }, k[2] = { 0, 0 }, bb[2] = { 1 << (bd - 1), 1 << (bd - 1) };
    pel dsy[64 * 64];
#define L(xx, yy) ((int)Y.p[(ptrdiff_t)(yy) * Y.pitch + (xx)])
    if (!at && !al) {
        for (int c = 0; c < 2; c++)
            for (int i = 0; i < h; i++)
                for (int j = 0; j < w; j++)
                    C[c].p[(ptrdiff_t)(y + i) * C[c].pitch + x + j] = (pel)(1 << (bd - 1));
        return;
    }
    /* luma at chroma resolution (cclm_get_luma_rec_pixels :282-335) */
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            const int lx = x0 + (j << hs), ly = y0 + (i << vs);
            int v;
            if (!hs && !vs) {
                v = L(lx, ly);
            } else {
                const int xl = (j || al) ? lx - 1 : lx;
                if (!vs)
                    v = (L(xl, ly) + 2 * L(lx, ly) + L(lx + 1, ly) + 2) >> 2;
                else if (colloc)
                    v = (L(xl, ly) + L(lx, (i || at) ? ly - 1 : ly) + 4 * L(lx, ly) + L(lx + 1, ly) + L(lx, ly + 1) + 4) >> 3;
                else
                    v = (L(xl, ly) + L(xl, ly + 1) + 2 * L(lx, ly) + 2 * L(lx, ly + 1) + L(lx + 1, ly) + L(lx + 1, ly + 1) + 4) >> 3;
            }
            dsy[i * w + j] = (pel)v;
        }
    /* the neighbour positions that enter the model (cclm_get_select_pos :61-88) */
    const int mode = b->pred_mode, is4 = !at || !al || mode != 81;
    int n[2], cnt[2] = { 0, 0 }, pos[2][4];
    if (mode == 81) { n[0] = at ? w : 0; n[1] = al ? h : 0; }
    else {
        n[0] = (at && mode == 83) ? o_min(w + o_min(w, h), b->avail_top) : 0;
        n[1] = (al && mode == 82) ? o_min(h + o_min(w, h), b->avail_left) : 0;
    }
    if (n[0] || n[1]) {
        int sel[3][8] = { { 0 } }, mn[3], mx[3];
        for (int s = 0; s < 2; s++) {
            const int start = n[s] >> (2 + is4), step = o_max(1, n[s] >> (1 + is4));
            cnt[s] = o_min(n[s], (1 + is4) << 1);
            for (int c = 0; c < cnt[s]; c++)
                pos[s][c] = start + c * step;
        }
        const int ctu_top = !(y0 & ((1 << f->ctb_log2) - 1));
        for (int i = 0; i < cnt[0]; i++) {                         /* above (cclm_select_luma :99-139) */
            const int px = pos[0][i] << hs, lx = x0 + px;
            int v;
            if (!hs && !vs) {
                v = L(x0 + pos[0][i], y0 - at);
            } else {
                const int xl = (px || al) ? lx - 1 : lx;
                if (vs && !ctu_top) {
                    const int ly = y0 - 2;
                    if (colloc) v = (L(lx, ly - 1) + L(xl, ly) + 4 * L(lx, ly) + L(lx + 1, ly) + L(lx, ly + 1) + 4) >> 3;
                    else        v = (L(xl, ly) + L(xl, ly + 1) + 2 * (L(lx, ly) + L(lx, ly + 1)) + L(lx + 1, ly) + L(lx + 1, ly + 1) + 4) >> 3;
                } else {
                    const int ly = y0 - 1;
                    v = (L(xl, ly) + 2 * L(lx, ly) + L(lx + 1, ly) + 2) >> 2;
                }
            }
            sel[0][i] = v;
            for (int c = 0; c < 2; c++)
                sel[1 + c][i] = C[c].p[(ptrdiff_t)(y - 1) * C[c].pitch + x + pos[0][i]];
        }
        for (int i = 0; i < cnt[1]; i++) {                         /* left (:141-166) */
            int v;
            if (!hs && !vs) {
                v = L(x0 - al, y0 + pos[1][i]);
            } else {
                const int ly = y0 + (pos[1][i] << vs), lx = x0 - (1 + hs) * al, xl = lx - al;
                if (!vs)         v = (L(xl, ly) + 2 * L(lx, ly) + L(lx + 1, ly) + 2) >> 2;
                else if (colloc) v = (L(xl, ly) + L(lx, (pos[1][i] || at) ? ly - 1 : ly) + 4 * L(lx, ly) + L(lx + 1, ly) + L(lx, ly + 1) + 4) >> 3;
                else             v = (L(xl, ly) + L(xl, ly + 1) + 2 * L(lx, ly) + 2 * L(lx, ly + 1) + L(lx + 1, ly) + L(lx + 1, ly + 1) + 4) >> 3;
            }
            sel[0][cnt[0] + i] = v;
            for (int c = 0; c < 2; c++)
                sel[1 + c][cnt[0] + i] = C[c].p[(ptrdiff_t)(y + pos[1][i]) * C[c].pitch + x - 1];
        }
        if (cnt[0] + cnt[1] == 2)
            for (int c = 0; c < 3; c++) {
                sel[c][3] = sel[c][0]; sel[c][2] = sel[c][1]; sel[c][0] = sel[c][1]; sel[c][1] = sel[c][3];
            }
        {   /* two smallest / two largest luma values by the reference's compare-exchange network (:203-226) */
            int lo[2] = { 0, 2 }, hi[2] = { 1, 3 }, t;
#define SWAP(p, q) do { t = (p); (p) = (q); (q) = t; } while (0)
            if (sel[0][lo[0]] > sel[0][lo[1]]) SWAP(lo[0], lo[1]);
            if (sel[0][hi[0]] > sel[0][hi[1]]) SWAP(hi[0], hi[1]);
            if (sel[0][lo[0]] > sel[0][hi[1]]) { SWAP(lo[0], hi[0]); SWAP(lo[1], hi[1]); }
            if (sel[0][lo[1]] > sel[0][hi[0]]) SWAP(lo[1], hi[0]);
#undef SWAP
            for (int c = 0; c < 3; c++) {
                mx[c] = (sel[c][hi[0]] + sel[c][hi[1]] + 1) >> 1;
                mn[c] = (sel[c][lo[0]] + sel[c][lo[1]] + 1) >> 1;
            }
        }
        const int diff = mx[0] - mn[0];
        for (int c = 0; c < 2; c++) {
            if (!diff) { a[c] = k[c] = 0; bb[c] = mn[c + 1]; continue; }
            static const int sig[] = { 0, 7, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1, 1, 1, 0 };
            const int diffc = mx[c + 1] - mn[c + 1];
            int lx = o_ilog2(diff);
            const int norm = ((diff << 4) >> lx) & 15;
            lx += norm ? 1 : 0;
            const int ly = o_abs(diffc) > 0 ? o_ilog2(o_abs(diffc)) + 1 : 0;
            const int v = sig[norm] | 8;
            a[c] = (diffc * v + ((1 << ly) >> 1)) >> ly;
            k[c] = o_max(1, 3 + lx - ly);
            if (3 + lx - ly < 1)
                a[c] = o_sign(a[c]) * 15;
            bb[c] = mn[c + 1] - ((a[c] * mn[0]) >> k[c]);
        }
    }
#undef L
    for (int c = 0; c < 2; c++)
        for (int i = 0; i < h; i++)
            for (int j = 0; j < w; j++)
                C[c].p[(ptrdiff_t)(y + i) * C[c].pitch + x + j] = (pel)o_clip_pel(((dsy[i * w + j] * a[c]) >> k[c]) + bb[c], bd);
}

void vvco_intra_pred_frame(const VVCCudaFrame *f, const VVCCudaIntraBlk *blks, int n)
{
    for (int i = 0; i < n; i++) {
        if (blks[i].kind == VVC_CUDA_INTRA_KIND_CCLM)
            cclm_block(f, &blks[i]);
        else
            intra_pred_block(f, &blks[i]);
    }
}

void vvco_itx_frame_q(const VVCCudaFrame *f, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs, int range);

/* prediction and residual alternating wave by wave (any order that respects the dependencies gives the same picture:
 * the tests hand the oracle one block per wave in decoding order) */
void vvco_intra_recon_frame(const VVCCudaFrame *f, const VVCCudaIntraBlk *blks, const int32_t *blk_end, const VVCCudaCoeffs *co,
                            const VVCCudaTB *tbs, const int32_t *tb_end, int n_waves, int range)
{
    int b0 = 0, t0 = 0;
    for (int g = 0; g < n_waves; g++) {
        vvco_intra_pred_frame(f, blks + b0, blk_end[g] - b0);
        if (tb_end[g] > t0) {
            VVCCudaCoeffs c = *co;
            if (c.quant)
                c.quant += t0;
            vvco_itx_frame_q(f, &c, tbs + t0, tb_end[g] - t0, range);
        }
        b0 = blk_end[g]; t0 = tb_end[g];
    }
}

void vvco_ciip_frame(const VVCCudaFrame *dst, const VVCCudaFrame *inter, const VVCCudaCiip *blocks, int n)
{
    for (int i = 0; i < n; i++) {
        const VVCCudaCiip *b = &blocks[i];
        const OPlane d = o_plane(dst, b->c_idx, b->pic), s = o_plane(inter, b->c_idx, b->pic);
        const int wi = b->intra_weight;
        for (int y = 0; y < b->h; y++)
            for (int x = 0; x < b->w; x++) {
                pel *p = &d.p[(b->y0 + y) * d.pitch + b->x0 + x];
                *p = (pel)((*p * wi + s.p[(b->y0 + y) * s.pitch + b->x0 + x] * (4 - wi) + 2) >> 2);
            }
    }
}
