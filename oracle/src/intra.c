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

void vvco_intra_leaf_frame(const VVCCudaFrame *f, const VVCCudaIntraPB *pbs, int n, const uint16_t *edges)
{
    const int bd = f->bit_depth;
    for (int i = 0; i < n; i++) {
        const VVCCudaIntraPB *b = &pbs[i];
        const OPlane pl = o_plane(f, b->c_idx, b->pic);
        pel *dst = pl.p + b->y0 * pl.pitch + b->x0;
        const pel *top = edges + b->top, *left = edges + b->left;
        const int w = b->w, h = b->h;
        switch (b->kind) {
        case VVC_CUDA_INTRA_PLANAR: {
            const int lw = o_ilog2(w), lh = o_ilog2(h);
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    const int pv = ((h - 1 - y) * top[x] + (y + 1) * left[h]) << lw;
                    const int ph = ((w - 1 - x) * left[y] + (x + 1) * top[w]) << lh;
                    dst[y * pl.pitch + x] = (pel)((pv + ph + w * h) >> (lw + lh + 1));
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
                    dst[y * pl.pitch + x] = dc;
            break;
        }
        case VVC_CUDA_INTRA_VERT:
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++)
                    dst[y * pl.pitch + x] = top[x];
            break;
        case VVC_CUDA_INTRA_HORZ:
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++)
                    dst[y * pl.pitch + x] = left[y];
            break;
        case VVC_CUDA_INTRA_ANGULAR_V: angular(dst, pl.pitch, top, left, b, bd, 1); break;
        case VVC_CUDA_INTRA_ANGULAR_H: angular(dst, pl.pitch, top, left, b, bd, 0); break;
        default:                       mip(dst, pl.pitch, top, left, b, bd); break;
        }
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
