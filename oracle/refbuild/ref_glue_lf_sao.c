/*
 * TEST INFRASTRUCTURE - frame-level deblocking and SAO through the UNMODIFIED reference
 * table entries (lf.filter_luma/chroma, sao.band_filter/edge_filter/edge_restore).
 *
 * The walks reproduce the reference drivers:
 *   ff_vvc_deblock_vertical / _horizontal   libavcodec/vvc/vvc_filter.c:861-1003
 *       (8-sample calls: 2 luma segments or 8 / (4 >> shift) chroma segments per call; x = 0 and
 *        y = 0 skipped; in place; CTU raster order)
 *   ff_vvc_sao_filter                       libavcodec/vvc/vvc_filter.c:154-298
 *       (band in one call; edge from a copy with a 1-sample halo of pre-SAO samples, fixed
 *        source stride 2*128+64 bytes, then edge_restore[restore])
 * The per-segment tc/beta/max-length values and the SAO edge flags come from our descriptors;
 * all pixel arithmetic is the reference's.
 */
#include <stdint.h>
#include <string.h>
#include "libavcodec/avcodec.h"
#include "libavcodec/vvc/vvc_ctu.h"
#include "libavcodec/vvc/vvcdsp.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

typedef uint16_t pixel;

static void copy_frame(const VVCCudaFrame *d, const VVCCudaFrame *s)
{
    const int planes = s->chroma_format_idc ? 3 : 1;
    if (d->data[0] == s->data[0])
        return;
    for (int k = 0; k < s->batch; k++)
        for (int c = 0; c < planes; c++) {
            const int w = c ? s->width >> s->hshift : s->width, h = c ? s->height >> s->vshift : s->height;
            for (int y = 0; y < h; y++)
                memcpy((uint8_t *)d->data[c] + k * d->batch_stride[c] + y * d->stride[c],
                       (const uint8_t *)s->data[c] + k * s->batch_stride[c] + y * s->stride[c], w * sizeof(pixel));
        }
}

void vvcref_deblock_frame(const VVCCudaFrame *dstf, const VVCCudaFrame *srcf, const VVCCudaDeblockMaps *maps, int dir)
{
    const VVCDSPContext *dsp = vvcref_dsp(srcf->bit_depth);
    const int ctb = 1 << srcf->ctb_log2;
    const int cols = (srcf->width + ctb - 1) >> srcf->ctb_log2, rows = (srcf->height + ctb - 1) >> srcf->ctb_log2;
    const int planes = srcf->chroma_format_idc ? 3 : 1;
    const uint8_t no_p[4] = { 0 }, no_q[4] = { 0 };

    copy_frame(dstf, srcf);
    for (int k = 0; k < srcf->batch; k++)
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++)
                for (int c = 0; c < planes; c++) {
                    const int hs = c ? srcf->hshift : 0, vs = c ? srcf->vshift : 0;
                    const int pw = srcf->width >> hs, ph = srcf->height >> vs;
                    uint8_t *plane = (uint8_t *)dstf->data[c] + k * dstf->batch_stride[c];
                    const ptrdiff_t stride = dstf->stride[c];
                    const int grid = c ? 8 : 4;
                    const int shift = c ? (dir ? vs : hs) : 0;
                    const int seg = c ? 4 >> shift : 4;
                    const int nseg = 8 / seg;
                    const int x0 = (cx * ctb) >> hs, y0 = (cy * ctb) >> vs;
                    const int x1 = x0 + (ctb >> hs) < pw ? x0 + (ctb >> hs) : pw;
                    const int y1 = y0 + (ctb >> vs) < ph ? y0 + (ctb >> vs) : ph;
                    const VVCCudaDbkEdge *map = maps->edge[dir][c] + (size_t)k * maps->size[dir][c];
                    const int mp = maps->pitch[dir][c];

                    if (dir) {
                        for (int y = y0; y < y1; y += 8)
                            for (int x = x0 ? x0 : grid; x < x1; x += grid) {
                                int32_t beta[4] = { 0 }, tc[4] = { 0 };
                                uint8_t lp[4] = { 0 }, lq[4] = { 0 };
                                int any = 0;
                                for (int i = 0; i < nseg; i++) {
                                    if (y + i * seg >= y1)
                                        continue;
                                    const VVCCudaDbkEdge e = map[((y + i * seg) / seg) * mp + x / grid];
                                    tc[i] = e.tc; beta[i] = e.beta; lp[i] = e.max_len & 15; lq[i] = e.max_len >> 4;
                                    any |= e.tc != 0;
                                }
                                if (!any)
                                    continue;
                                if (!c) dsp->lf.filter_luma[1](plane + y * stride + x * 2, stride, beta, tc, no_p, no_q, lp, lq, 0);
                                else    dsp->lf.filter_chroma[1](plane + y * stride + x * 2, stride, beta, tc, no_p, no_q, lp, lq, shift);
                            }
                    } else {
                        for (int y = y0; y < y1; y += grid) {
                            const int ctu_edge = !((y << vs) % ctb);
                            if (!y)
                                continue;
                            for (int x = x0; x < x1; x += 8) {
                                int32_t beta[4] = { 0 }, tc[4] = { 0 };
                                uint8_t lp[4] = { 0 }, lq[4] = { 0 };
                                int any = 0;
                                for (int i = 0; i < nseg; i++) {
                                    if (x + i * seg >= x1)
                                        continue;
                                    const VVCCudaDbkEdge e = map[(y / grid) * mp + (x + i * seg) / seg];
                                    tc[i] = e.tc; beta[i] = e.beta; lp[i] = e.max_len & 15; lq[i] = e.max_len >> 4;
                                    any |= e.tc != 0;
                                }
                                if (!any)
                                    continue;
                                if (!c) dsp->lf.filter_luma[0](plane + y * stride + x * 2, stride, beta, tc, no_p, no_q, lp, lq, ctu_edge);
                                else    dsp->lf.filter_chroma[0](plane + y * stride + x * 2, stride, beta, tc, no_p, no_q, lp, lq, shift);
                            }
                        }
                    }
                }
}

#define SAO_STRIDE (2 * MAX_PB_SIZE + AV_INPUT_BUFFER_PADDING_SIZE)   /* bytes, vvc_filter.c:236 */

void vvcref_sao_frame(const VVCCudaFrame *dstf, const VVCCudaFrame *srcf, const VVCCudaSAOCtb *ctbs)
{
    static const uint8_t sao_tab[16] = { 0, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8 };
    const VVCDSPContext *dsp = vvcref_dsp(srcf->bit_depth);
    const int ctb = 1 << srcf->ctb_log2;
    const int cols = (srcf->width + ctb - 1) >> srcf->ctb_log2, rows = (srcf->height + ctb - 1) >> srcf->ctb_log2;
    const int planes = srcf->chroma_format_idc ? 3 : 1;
    /* (128 + 2) rows of SAO_STRIDE bytes + slack, 32-byte aligned like lc->sao_buffer */
    static _Thread_local uint8_t buffer[SAO_STRIDE * (128 + 4) + 64] __attribute__((aligned(32)));

    copy_frame(dstf, srcf);
    for (int k = 0; k < srcf->batch; k++)
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++) {
                const VVCCudaSAOCtb *p = &ctbs[(size_t)k * cols * rows + cy * cols + cx];
                SAOParams sao;
                int edges[4] = { cx == 0, cy == 0, cx == cols - 1, cy == rows - 1 };
                uint8_t vert_edge[2]  = { p->no_filter & 1, (p->no_filter >> 1) & 1 };
                uint8_t horiz_edge[2] = { (p->no_filter >> 2) & 1, (p->no_filter >> 3) & 1 };
                uint8_t diag_edge[4]  = { (p->no_filter >> 4) & 1, (p->no_filter >> 5) & 1, (p->no_filter >> 6) & 1, (p->no_filter >> 7) & 1 };
                memset(&sao, 0, sizeof(sao));
                for (int c = 0; c < 3; c++) {
                    sao.band_position[c] = p->band_position[c];
                    sao.eo_class[c] = p->eo_class[c];
                    sao.type_idx[c] = p->type_idx[c];
                    memcpy(sao.offset_val[c], p->offset_val[c], sizeof(sao.offset_val[c]));
                }
                for (int c = 0; c < planes; c++) {
                    const int hs = c ? srcf->hshift : 0, vs = c ? srcf->vshift : 0;
                    const int pw = srcf->width >> hs, ph = srcf->height >> vs;
                    const int x0 = (cx * ctb) >> hs, y0 = (cy * ctb) >> vs;
                    const int w = (ctb >> hs) < pw - x0 ? (ctb >> hs) : pw - x0;
                    const int h = (ctb >> vs) < ph - y0 ? (ctb >> vs) : ph - y0;
                    const int tab = sao_tab[(FFALIGN(w, 8) >> 3) - 1];
                    const uint8_t *splane = (const uint8_t *)srcf->data[c] + k * srcf->batch_stride[c];
                    uint8_t *dplane = (uint8_t *)dstf->data[c] + k * dstf->batch_stride[c];
                    const ptrdiff_t ss = srcf->stride[c], ds = dstf->stride[c];
                    uint8_t *dst = dplane + y0 * ds + x0 * 2;
                    const uint8_t *src = splane + y0 * ss + x0 * 2;

                    if (p->type_idx[c] == SAO_BAND) {
                        dsp->sao.band_filter[tab](dst, src, ds, ss, sao.offset_val[c], sao.band_position[c], w, h);
                    } else if (p->type_idx[c] == SAO_EDGE) {
                        uint8_t *buf = buffer + SAO_STRIDE + AV_INPUT_BUFFER_PADDING_SIZE;
                        /* CTB plus the 1-sample ring that exists inside the picture, from the pre-SAO picture */
                        for (int y = -1; y <= h; y++) {
                            if (y0 + y < 0 || y0 + y >= ph)
                                continue;
                            for (int x = -1; x <= w; x++) {
                                if (x0 + x < 0 || x0 + x >= pw)
                                    continue;
                                ((pixel *)(buf + y * SAO_STRIDE))[x] = ((const pixel *)(splane + (y0 + y) * ss))[x0 + x];
                            }
                        }
                        dsp->sao.edge_filter[tab](dst, buf, ds, sao.offset_val[c], sao.eo_class[c], w, h);
                        dsp->sao.edge_restore[p->restore ? 1 : 0](dst, buf, ds, SAO_STRIDE, &sao, edges, w, h, c,
                                                                  vert_edge, horiz_edge, diag_edge);
                    }
                }
            }
}
