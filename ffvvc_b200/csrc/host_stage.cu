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

// ---- pre-padded reference planes (VVC_CUDA_OPT_REF_PAD) ------------------------------------------------------------
// What the reference fabricates per block when a motion vector points outside the picture (ff_emulated_edge_mc through
// emulated_edge, libavcodec/vvc/vvc_inter.c:33-58: every coordinate clamped to the picture) is stored once per picture
// instead: a margin of replicated samples around every plane.  One thread per margin sample; the interior is not touched.
namespace {

__global__ void __launch_bounds__(256) pad_frame_kernel(pel *plane, int pitch, long long bstride, int w, int h, int mx, int my)
{
    pel *pic = plane + blockIdx.y * bstride;
    const int W = w + 2 * mx;
    // margin samples in row-major order of the padded plane: my full rows above, then (left, right) strips of the h picture
    // rows, then my full rows below
    const long long top = (long long)my * W, sides = (long long)h * 2 * mx, total = 2 * top + sides;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        int x, y;
        if (i < top)              { y = (int)(i / W) - my;              x = (int)(i % W) - mx; }
        else if (i < top + sides) { const long long j = i - top; y = (int)(j / (2 * mx)); const int k = (int)(j % (2 * mx)); x = k < mx ? k - mx : w + k - mx; }
        else                      { const long long j = i - top - sides; y = h + (int)(j / W); x = (int)(j % W) - mx; }
        pic[(long long)y * pitch + x] = pic[(long long)min(max(y, 0), h - 1) * pitch + min(max(x, 0), w - 1)];
    }
}

}  // namespace

extern "C" int vvc_cuda_pad_frame(VVCCudaCtx *ctx, const VVCCudaFrame *f, int pad)
{
    if (ctx->err)
        return ctx->err;
    if (!f || pad < 0 || (pad & 3))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "pad_frame: null frame or a margin that is not a multiple of 4");
    if (!pad)
        return VVC_CUDA_OK;
    const int planes = f->chroma_format_idc ? 3 : 1;
    for (int c = 0; c < planes; c++) {
        const int w = c ? f->width >> f->hshift : f->width, h = c ? f->height >> f->vshift : f->height;
        const int mx = c ? pad >> f->hshift : pad, my = c ? pad >> f->vshift : pad;
        if (f->stride[c] / 2 < w + 2 * mx)
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "pad_frame: the row pitch of plane %d leaves no room for a margin of %d", c, mx);
        const long long total = 2ll * my * (w + 2 * mx) + 2ll * mx * h;
        const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        pad_frame_kernel<<<dim3(grid, f->batch), 256, 0, ctx->stream>>>((pel *)f->data[c], (int)(f->stride[c] / 2), f->batch_stride[c] / 2,
                                                                        w, h, mx, my);
        VVC_LAUNCHED(ctx);
    }
    return VVC_CUDA_OK;
}
