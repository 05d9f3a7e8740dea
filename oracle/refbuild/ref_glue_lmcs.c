/*
 * TEST INFRASTRUCTURE - LMCS through the UNMODIFIED reference entry lmcs.filter, called per CTU like
 * ff_vvc_lmcs_filter (libavcodec/vvc/vvc_filter.c:1322-1332) or per rectangle like predict_inter
 * (libavcodec/vvc/vvc_inter.c:888-891).
 */
#include <stdint.h>
#include "libavcodec/vvc/vvcdsp.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

void vvcref_lmcs_frame(const VVCCudaFrame *f, const uint16_t *lut, const uint8_t *ctb_enable)
{
    const VVCDSPContext *dsp = vvcref_dsp(f->bit_depth);
    const int ctb = 1 << f->ctb_log2;
    const int cols = (f->width + ctb - 1) >> f->ctb_log2, rows = (f->height + ctb - 1) >> f->ctb_log2;
    for (int k = 0; k < f->batch; k++)
        for (int cy = 0; cy < rows; cy++)
            for (int cx = 0; cx < cols; cx++) {
                const int x = cx * ctb, y = cy * ctb;
                const int w = ctb < f->width - x ? ctb : f->width - x, h = ctb < f->height - y ? ctb : f->height - y;
                if (ctb_enable && !ctb_enable[(size_t)k * cols * rows + cy * cols + cx])
                    continue;
                dsp->lmcs.filter((uint8_t *)f->data[0] + k * f->batch_stride[0] + y * f->stride[0] + x * 2,
                                 f->stride[0], w, h, (const uint8_t *)lut);
            }
}

void vvcref_lmcs_rects(const VVCCudaFrame *f, const uint16_t *lut, const VVCCudaRect *r, int n)
{
    const VVCDSPContext *dsp = vvcref_dsp(f->bit_depth);
    for (int i = 0; i < n; i++)
        dsp->lmcs.filter((uint8_t *)f->data[0] + r[i].pic * f->batch_stride[0] + r[i].y * f->stride[0] + r[i].x * 2,
                         f->stride[0], r[i].w, r[i].h, (const uint8_t *)lut);
}
