// Host <-> device staging used by the *_host entries and the per-call table shims.
// The reference hands the DSP layer caller-owned host buffers (SURVEY.md 8(b) "ownership"); these
// helpers copy them into the context's device staging area and back without retaining pointers.
#include "common.cuh"

static inline int plane_w(const VVCCudaFrame *f, int c) { return c ? f->width >> f->hshift : f->width; }
static inline int plane_h(const VVCCudaFrame *f, int c) { return c ? f->height >> f->vshift : f->height; }
static inline int plane_count(const VVCCudaFrame *f) { return f->chroma_format_idc ? 3 : 1; }

size_t vvc_stage_frame_size(const VVCCudaFrame *f)
{
    size_t total = 0;
    for (int c = 0; c < plane_count(f); c++)
        total += align_up((size_t)plane_w(f, c) * sizeof(pel), 256) * plane_h(f, c) * f->batch;
    return total;
}

void vvc_stage_frame_layout(const VVCCudaFrame *host, void *dbase, VVCCudaFrame *dev)
{
    uint8_t *at = (uint8_t *)dbase;
    *dev = *host;
    for (int c = 0; c < 3; c++) {
        dev->data[c] = NULL; dev->stride[c] = 0; dev->batch_stride[c] = 0;
    }
    for (int c = 0; c < plane_count(host); c++) {
        const size_t pitch = align_up((size_t)plane_w(host, c) * sizeof(pel), 256);
        dev->data[c]         = at;
        dev->stride[c]       = (ptrdiff_t)pitch;
        dev->batch_stride[c] = (ptrdiff_t)(pitch * plane_h(host, c));
        at += pitch * plane_h(host, c) * host->batch;
    }
}

static int copy_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src, cudaMemcpyKind kind)
{
    for (int c = 0; c < plane_count(src); c++)
        for (int k = 0; k < src->batch; k++) {
            const size_t row = (size_t)plane_w(src, c) * sizeof(pel);
            if ((size_t)dst->stride[c] == row && (size_t)src->stride[c] == row) {      // both sides dense: one linear copy
                VVC_TRY(ctx, cudaMemcpyAsync((uint8_t *)dst->data[c] + k * dst->batch_stride[c],
                                             (const uint8_t *)src->data[c] + k * src->batch_stride[c],
                                             row * plane_h(src, c), kind, ctx->stream));
                continue;
            }
            VVC_TRY(ctx, cudaMemcpy2DAsync((uint8_t *)dst->data[c] + k * dst->batch_stride[c], dst->stride[c],
                                           (const uint8_t *)src->data[c] + k * src->batch_stride[c], src->stride[c],
                                           row, plane_h(src, c), kind, ctx->stream));
        }
    return 0;
}

int vvc_stage_frame_h2d(VVCCudaCtx *ctx, const VVCCudaFrame *dev, const VVCCudaFrame *host)
{
    return copy_frame(ctx, dev, host, cudaMemcpyHostToDevice);
}

int vvc_stage_frame_d2h(VVCCudaCtx *ctx, const VVCCudaFrame *host, const VVCCudaFrame *dev)
{
    return copy_frame(ctx, host, dev, cudaMemcpyDeviceToHost);
}
