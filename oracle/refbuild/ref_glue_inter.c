/*
 * TEST INFRASTRUCTURE - inter prediction stage through the UNMODIFIED reference entries
 *   inter.put / put_uni / put_uni_w / avg / w_avg / put_gpm / bdof_fetch_samples / fetch_samples /
 *   apply_prof / apply_prof_uni(_w) / apply_bdof / sad / dmvr, VideoDSPContext.emulated_edge_mc and
 *   ff_vvc_clip_mv,
 * called in the order of the reference's INTER driver (libavcodec/vvc/vvc_inter.c): the driver's
 * functions are static and take decoder contexts, so their call sequence (which entry, which
 * scratch tile, which edge-emulation window) is re-expressed here per VVCCudaPB record; every
 * sample is produced by the reference's own code.
 *
 *   emulated_edge / _dmvr / _bilinear   vvc_inter.c:33-110
 *   luma_mc_uni / luma_mc_bi            :222-296      chroma_mc_uni / chroma_mc_bi  :298-366
 *   luma_prof_uni / luma_prof_bi        :368-446      pred_gpm_blk                  :466-521
 *   dmvr_mv_refine                      :685-748      (parametric_mv_refine :642-681 is host
 *   arithmetic of the driver, not a table entry: its 30 lines are restated below)
 */
#include <stdint.h>
#include <string.h>
#include "libavutil/common.h"
#include "libavcodec/videodsp.h"
#include "libavcodec/vvc/vvcdsp.h"
#include "libavcodec/vvc/vvc_data.h"
#include "libavcodec/vvc/vvc_ctu.h"
#include "libavcodec/vvc/vvc_mvs.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

#define PB_STRIDE 128                 /* MAX_PB_SIZE */
#define TEMP_OFFSET (PB_STRIDE + 32)  /* PROF_TEMP_OFFSET */
#define EMU_STRIDE 160                /* EDGE_EMU_BUFFER_STRIDE, in samples */

typedef struct Ref {
    const uint8_t *data;
    ptrdiff_t      stride;
    int            w, h;
} Ref;

static Ref ref_plane(const VVCCudaFrame *f, int slot, int c)
{
    Ref r;
    r.data   = (const uint8_t *)f->data[c] + (ptrdiff_t)slot * f->batch_stride[c];
    r.stride = f->stride[c];
    r.w      = c ? f->width >> f->hshift : f->width;
    r.h      = c ? f->height >> f->vshift : f->height;
    return r;
}

static VideoDSPContext g_vdsp[2];
static int g_vdsp_ready[2];
static VideoDSPContext *vdsp(int bd)
{
    const int s = bd > 8;
    if (!g_vdsp_ready[s]) {
        ff_videodsp_init(&g_vdsp[s], bd);
        g_vdsp_ready[s] = 1;
    }
    return &g_vdsp[s];
}

/* emulated_edge(), vvc_inter.c:33-58 */
static void edge(uint8_t *buf, const uint8_t **src, ptrdiff_t *stride, const Ref *r, int x_off, int y_off,
                 int bw, int bh, int before, int after, int bd)
{
    if (x_off < before || y_off < before || x_off >= r->w - bw - after || y_off >= r->h - bh - after) {
        const ptrdiff_t es = EMU_STRIDE * 2;
        const int offset = before * *stride + before * 2, buf_offset = before * es + before * 2;
        vdsp(bd)->emulated_edge_mc(buf, *src - offset, es, *stride, bw + before + after, bh + before + after,
                                   x_off - before, y_off - before, r->w, r->h);
        *src = buf + buf_offset;
        *stride = es;
    }
}

/* emulated_edge_dmvr(), vvc_inter.c:60-89 */
static void edge_dmvr(uint8_t *buf, const uint8_t **src, ptrdiff_t *stride, const Ref *r, int x_sb, int y_sb,
                      int x_off, int y_off, int bw, int bh, int before, int after, int bd)
{
    if (x_off < before || y_off < before || x_off >= r->w - bw - after || y_off >= r->h - bh - after ||
        x_off != x_sb || y_off != y_sb) {
        const ptrdiff_t es = EMU_STRIDE * 2;
        const int offset = before * *stride + before * 2, buf_offset = before * es + before * 2;
        const int start_x = FFMIN(FFMAX(x_sb - before, 0), r->w - 1);
        const int start_y = FFMIN(FFMAX(y_sb - before, 0), r->h - 1);
        const int width   = FFMAX(FFMIN(r->w, x_sb + bw + after) - start_x, 1);
        const int height  = FFMAX(FFMIN(r->h, y_sb + bh + after) - start_y, 1);
        vdsp(bd)->emulated_edge_mc(buf, *src - offset, es, *stride, bw + before + after, bh + before + after,
                                   x_off - start_x - before, y_off - start_y - before, width, height);
        *src = buf + buf_offset;
        *stride = es;
    }
}

/* parametric_mv_refine(), vvc_inter.c:642-681 (driver arithmetic) */
static int parametric_mv_refine(const int *sad, const int stride)
{
    const int sad_minus = sad[-stride], sad_center = sad[0], sad_plus = sad[stride];
    int dmvc, denom = ((sad_minus + sad_plus) - (sad_center << 1)) << 3;
    if (!denom)
        return 0;
    if (sad_minus == sad_center)
        return -8;
    if (sad_plus == sad_center)
        return 8;
    {
        int num = (sad_minus - sad_plus) * (1 << 4), sign_num = 0, quotient = 0, counter = 3;
        if (num < 0) { num = -num; sign_num = 1; }
        while (counter > 0) {
            counter--;
            quotient <<= 1;
            if (num >= denom) { num -= denom; quotient++; }
            denom >>= 1;
        }
        dmvc = sign_num ? -quotient : quotient;
    }
    return dmvc;
}

typedef struct Scratch {
    int16_t tmp[3][PB_STRIDE * PB_STRIDE];       /* lc->tmp, tmp1, tmp2 (vvc_ctu.h:376-378) */
    uint8_t edge_emu[EMU_STRIDE * 2 * (PB_STRIDE + 8)];
} Scratch;

static int bi_weight(int *denom, int *w0, int *w1, int *o0, int *o1, const VVCCudaPB *pb, const VVCCudaWP *wp, int c)
{
    static const int bcw_w_lut[] = { 4, 5, 3, 10, -2 };
    const int weight_flag = (pb->flags & VVC_CUDA_PB_WEIGHTED) && !(pb->flags & VVC_CUDA_PB_DMVR);
    if (!weight_flag && !pb->bcw_idx)
        return 0;
    if (pb->bcw_idx) {
        *denom = 2; *w1 = bcw_w_lut[pb->bcw_idx]; *w0 = 8 - *w1; *o0 = *o1 = 0;
    } else {
        const VVCCudaWP *e = &wp[pb->wp];
        *denom = e->log2_denom[c > 0];
        *w0 = e->weight[0][c]; *w1 = e->weight[1][c]; *o0 = e->offset[0][c]; *o1 = e->offset[1][c];
    }
    return 1;
}

static void predict(Scratch *sc, const VVCCudaFrame *dst, const VVCCudaFrame *refs, const VVCCudaPB *pb,
                    const VVCCudaWP *wp, const VVCCudaProf *prof, VVCCudaDmvrOut *out)
{
    const int bd = dst->bit_depth;
    const VVCDSPContext *dsp = vvcref_dsp(bd);
    const int gpm = pb->flags & VVC_CUDA_PB_GPM, dmvr_flag = pb->flags & VVC_CUDA_PB_DMVR;
    Mv mv[2] = { { pb->mv[0][0], pb->mv[0][1] } , { pb->mv[1][0], pb->mv[1][1] } }, orig[2];
    int sb_bdof_flag = !!(pb->flags & VVC_CUDA_PB_BDOF);
    memcpy(orig, mv, sizeof(mv));

    if (dmvr_flag && (pb->planes & VVC_CUDA_PB_LUMA)) {          /* dmvr_mv_refine, :685-748 */
        const int block_w = pb->w, block_h = pb->h, sr_range = 2;
        int sad[5][5], min_dx = 2, min_dy = 2, min_sad, dx = 2, dy = 2;
        for (int i = 0; i < 2; i++) {
            const Ref r = ref_plane(refs, pb->ref[i], 0);
            const int pred_w = block_w + 2 * sr_range, pred_h = block_h + 2 * sr_range;
            const int mx = mv[i].x & 0xf, my = mv[i].y & 0xf;
            const int ox = pb->x0 + (mv[i].x >> 4) - sr_range, oy = pb->y0 + (mv[i].y >> 4) - sr_range;
            ptrdiff_t src_stride = r.stride;
            const uint8_t *src = r.data + oy * src_stride + ox * 2;
            edge(sc->edge_emu, &src, &src_stride, &r, ox, oy, pred_w, pred_h, 0, 1, bd);   /* emulated_edge_bilinear */
            dsp->inter.dmvr[!!my][!!mx](sc->tmp[i], src, src_stride, pred_h, mx, my, pred_w);
        }
        min_sad = dsp->inter.sad(sc->tmp[0], sc->tmp[1], dx, dy, block_w, block_h);
        min_sad -= min_sad >> 2;
        sad[dy][dx] = min_sad;
        if (min_sad >= block_w * block_h) {
            int dmv[2];
            for (dy = 0; dy < 5; dy++)
                for (dx = 0; dx < 5; dx++)
                    if (dx != sr_range || dy != sr_range) {
                        sad[dy][dx] = dsp->inter.sad(sc->tmp[0], sc->tmp[1], dx, dy, block_w, block_h);
                        if (sad[dy][dx] < min_sad) {
                            min_sad = sad[dy][dx]; min_dx = dx; min_dy = dy;
                        }
                    }
            dmv[0] = (min_dx - sr_range) * (1 << 4);
            dmv[1] = (min_dy - sr_range) * (1 << 4);
            if (min_dx != 0 && min_dx != 4 && min_dy != 0 && min_dy != 4) {
                dmv[0] += parametric_mv_refine(&sad[min_dy][min_dx], 1);
                dmv[1] += parametric_mv_refine(&sad[min_dy][min_dx], 5);
            }
            for (int i = 0; i < 2; i++) {
                mv[i].x += (1 - 2 * i) * dmv[0];
                mv[i].y += (1 - 2 * i) * dmv[1];
                ff_vvc_clip_mv(&mv[i]);
            }
        }
        if (min_sad < 2 * block_w * block_h)
            sb_bdof_flag = 0;
        if (out) {
            for (int i = 0; i < 2; i++) { out->mv[i][0] = mv[i].x; out->mv[i][1] = mv[i].y; }
            out->min_sad = min_sad;
            out->bdof_applied = sb_bdof_flag;
        }
    }

    for (int c = 0; c < (dst->chroma_format_idc ? 3 : 1); c++) {
        if (!(pb->planes & (c ? VVC_CUDA_PB_CHROMA : VVC_CUDA_PB_LUMA)))
            continue;
        const int hs = c ? 1 : 0, before = c ? 1 : 3, after = c ? 2 : 4;
        const int block_w = pb->w >> hs, block_h = pb->h >> hs, x_off = pb->x0 >> hs, y_off = pb->y0 >> hs;
        const int idx = av_log2(block_w) - 1;
        const ptrdiff_t dst_stride = dst->stride[c];
        uint8_t *d = (uint8_t *)dst->data[c] + (ptrdiff_t)pb->pic * dst->batch_stride[c] + y_off * dst_stride + x_off * 2;
        const int use_bdof = !c && sb_bdof_flag && !gpm;
        int16_t *tmp[2] = { sc->tmp[0] + use_bdof * TEMP_OFFSET, sc->tmp[1] + use_bdof * TEMP_OFFSET };
        const int filt = gpm ? 0 : pb->filt;
        const int bi = gpm || pb->pred_flag == 3;

        for (int i = 0; i < 2; i++) {
            if (!gpm && !(pb->pred_flag & (1 << i)))
                continue;
            const Ref r = ref_plane(refs, pb->ref[i], c);
            const int mx = c ? av_mod_uintp2(mv[i].x, 4 + hs) << (1 - hs) : mv[i].x & 0xf;
            const int my = c ? av_mod_uintp2(mv[i].y, 4 + hs) << (1 - hs) : mv[i].y & 0xf;
            const int ox = x_off + (mv[i].x >> (4 + hs)), oy = y_off + (mv[i].y >> (4 + hs));
            const int8_t *hf = c ? ff_vvc_inter_chroma_filters[0][mx] : ff_vvc_inter_luma_filters[filt][mx];
            const int8_t *vf = c ? ff_vvc_inter_chroma_filters[0][my] : ff_vvc_inter_luma_filters[filt][my];
            ptrdiff_t src_stride = r.stride;
            const uint8_t *src = r.data + oy * src_stride + ox * 2;
            const int prof_flag = !c && !gpm && (pb->flags & (i ? VVC_CUDA_PB_PROF1 : VVC_CUDA_PB_PROF0));
            const int16_t *dmx = prof_flag ? prof[pb->prof].diff_mv_x[i] : NULL, *dmy = prof_flag ? prof[pb->prof].diff_mv_y[i] : NULL;
            int denom = 0, wx = 0, wox = 0;
            const int weight_uni = !bi && (pb->flags & VVC_CUDA_PB_WEIGHTED);
            if (weight_uni) {
                denom = wp[pb->wp].log2_denom[c > 0]; wx = wp[pb->wp].weight[i][c]; wox = wp[pb->wp].offset[i][c];
            }
            if (dmvr_flag && bi && !gpm)
                edge_dmvr(sc->edge_emu, &src, &src_stride, &r, x_off + (orig[i].x >> (4 + hs)), y_off + (orig[i].y >> (4 + hs)),
                          ox, oy, block_w, block_h, before, after, bd);
            else
                edge(sc->edge_emu, &src, &src_stride, &r, ox, oy, block_w, block_h, before, after, bd);

            if (!bi) {
                if (prof_flag) {                                     /* luma_prof_uni, :368-408 */
                    uint16_t *prof_tmp = (uint16_t *)sc->tmp[0] + TEMP_OFFSET;
                    dsp->inter.put[0][idx][!!my][!!mx]((int16_t *)prof_tmp, src, src_stride, 4, hf, vf, 4);
                    dsp->inter.fetch_samples((int16_t *)prof_tmp, src, src_stride, mx, my);
                    if (!weight_uni)
                        dsp->inter.apply_prof_uni(d, dst_stride, (int16_t *)prof_tmp, dmx, dmy);
                    else
                        dsp->inter.apply_prof_uni_w(d, dst_stride, (int16_t *)prof_tmp, dmx, dmy, denom, wx, wox);
                } else if (weight_uni) {                             /* luma_mc_uni :222-251, chroma_mc_uni :298-327 */
                    dsp->inter.put_uni_w[c > 0][idx][!!my][!!mx](d, dst_stride, src, src_stride, block_h, denom, wx, wox, hf, vf, block_w);
                } else {
                    dsp->inter.put_uni[c > 0][idx][!!my][!!mx](d, dst_stride, src, src_stride, block_h, hf, vf, block_w);
                }
            } else if (prof_flag) {                                  /* luma_prof_bi, :410-446 */
                uint16_t *prof_tmp = (uint16_t *)sc->tmp[2] + TEMP_OFFSET;
                dsp->inter.put[0][idx][!!my][!!mx]((int16_t *)prof_tmp, src, src_stride, 4, hf, vf, 4);
                dsp->inter.fetch_samples((int16_t *)prof_tmp, src, src_stride, mx, my);
                dsp->inter.apply_prof(tmp[i], (int16_t *)prof_tmp, dmx, dmy);
            } else {                                                 /* luma_mc_bi :253-296, chroma_mc_bi :329-366, luma_mc/chroma_mc :179-220 */
                dsp->inter.put[c > 0][idx][!!my][!!mx](tmp[i], src, src_stride, block_h, hf, vf, block_w);
                if (use_bdof)
                    dsp->inter.bdof_fetch_samples(tmp[i], src, src_stride, mx, my, block_w, block_h);
            }
        }
        if (!bi)
            continue;
        if (gpm) {                                                   /* pred_gpm_blk, :466-521 */
            const uint8_t *weights = &ff_vvc_gpm_weights[0][0] + pb->gpm_weights;
            dsp->inter.put_gpm(d, dst_stride, block_w, block_h, tmp[0], tmp[1], weights, pb->gpm_step_x << hs, pb->gpm_step_y << hs);
        } else if (use_bdof) {
            dsp->inter.apply_bdof(d, dst_stride, tmp[0], tmp[1], block_w, block_h);
        } else {
            int denom, w0, w1, o0, o1;
            if (bi_weight(&denom, &w0, &w1, &o0, &o1, pb, wp, c))
                dsp->inter.w_avg(d, dst_stride, tmp[0], tmp[1], block_w, block_h, denom, w0, w1, o0, o1);
            else
                dsp->inter.avg(d, dst_stride, tmp[0], tmp[1], block_w, block_h);
        }
    }
}

void vvcref_inter_frame(const VVCCudaFrame *dst, const VVCCudaFrame *refs, const VVCCudaPB *pbs, int n_pbs,
                        const VVCCudaWP *wp, const VVCCudaProf *prof, VVCCudaDmvrOut *dmvr_out)
{
    static _Thread_local Scratch sc;
    for (int i = 0; i < n_pbs; i++)
        predict(&sc, dst, refs, &pbs[i], wp, prof, dmvr_out ? &dmvr_out[i] : NULL);
}
