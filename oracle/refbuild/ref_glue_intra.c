/*
 * TEST INFRASTRUCTURE - intra leaf predictors and the CIIP blend through the UNMODIFIED reference table
 * entries intra.pred_planar / pred_dc / pred_v / pred_h / pred_angular_v / pred_angular_h / pred_mip
 * (libavcodec/vvc/vvc_intra_template.c:686-1015) and inter.put_ciip (vvc_inter_template.c:60-76), one call
 * per record, with the edge pointers IntraEdgeParams would carry.
 */
#include <stdint.h>
#include "libavcodec/vvc/vvcdsp.h"
#include "vvcdsp_cuda.h"

const VVCDSPContext *vvcref_dsp(int bit_depth);

void vvcref_intra_leaf_frame(const VVCCudaFrame *f, const VVCCudaIntraPB *pbs, int n, const uint16_t *edges)
{
    const VVCDSPContext *dsp = vvcref_dsp(f->bit_depth);
    for (int i = 0; i < n; i++) {
        const VVCCudaIntraPB *b = &pbs[i];
        uint8_t *dst = (uint8_t *)f->data[b->c_idx] + b->pic * f->batch_stride[b->c_idx] + b->y0 * f->stride[b->c_idx] + b->x0 * 2;
        const ptrdiff_t stride = f->stride[b->c_idx] / 2;       /* the leaf predictors take the stride in samples */
        const uint8_t *top = (const uint8_t *)(edges + b->top), *left = (const uint8_t *)(edges + b->left);
        switch (b->kind) {
        case VVC_CUDA_INTRA_PLANAR: dsp->intra.pred_planar(dst, top, left, b->w, b->h, stride); break;
        case VVC_CUDA_INTRA_DC:     dsp->intra.pred_dc(dst, top, left, b->w, b->h, stride); break;
        case VVC_CUDA_INTRA_VERT:   dsp->intra.pred_v(dst, top, b->w, b->h, stride); break;
        case VVC_CUDA_INTRA_HORZ:   dsp->intra.pred_h(dst, left, b->w, b->h, stride); break;
        case VVC_CUDA_INTRA_ANGULAR_V:
            dsp->intra.pred_angular_v(dst, top, left, b->w, b->h, stride, b->c_idx, b->mode, b->ref_idx, b->filter_flag,
                                      b->flags & VVC_CUDA_INTRA_PDPC);
            break;
        case VVC_CUDA_INTRA_ANGULAR_H:
            dsp->intra.pred_angular_h(dst, top, left, b->w, b->h, stride, b->c_idx, b->mode, b->ref_idx, b->filter_flag,
                                      b->flags & VVC_CUDA_INTRA_PDPC);
            break;
        default:
            dsp->intra.pred_mip(dst, top, left, b->w, b->h, stride, b->mode, !!(b->flags & VVC_CUDA_INTRA_MIP_TRANSPOSED));
            break;
        }
    }
}

void vvcref_ciip_frame(const VVCCudaFrame *dst, const VVCCudaFrame *inter, const VVCCudaCiip *blocks, int n)
{
    const VVCDSPContext *dsp = vvcref_dsp(dst->bit_depth);
    for (int i = 0; i < n; i++) {
        const VVCCudaCiip *b = &blocks[i];
        const int c = b->c_idx;
        dsp->inter.put_ciip((uint8_t *)dst->data[c] + b->pic * dst->batch_stride[c] + b->y0 * dst->stride[c] + b->x0 * 2, dst->stride[c],
                            b->w, b->h,
                            (const uint8_t *)inter->data[c] + b->pic * inter->batch_stride[c] + b->y0 * inter->stride[c] + b->x0 * 2,
                            inter->stride[c], b->intra_weight);
    }
}
