// LMCS stage for sm_100a: luma[x] = lut[luma[x]], in place.
//
// Replaces lmcs.filter (lmcs_filter_luma, libavcodec/vvc/vvc_filter_template.c:25-36) as driven per
// CTU by ff_vvc_lmcs_filter (libavcodec/vvc/vvc_filter.c:1322-1332) and per inter CU by
// predict_inter (libavcodec/vvc/vvc_inter.c:888-891).
//
// B200 design: streaming, HBM bound (2 B read + 2 B write per luma sample).  The LUT (2 KB at 10 bit,
// 8 KB at 12 bit) is staged once per CTA into shared memory; every thread maps 8 samples per 128-bit
// load/store.  The LUT is stored as one 32-bit word per entry so the 8 look-ups of a thread hit
// 8 independent banks instead of pairing two entries per bank.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void stage_lut(unsigned *s_lut, const uint16_t *lut, int n)
{
    for (int i = threadIdx.x; i < n; i += kThreads)
        s_lut[i] = __ldg(lut + i);
    __syncthreads();
}

__device__ __forceinline__ uint4 map8(uint4 v, const unsigned *s_lut)
{
    uint4 o;
    o.x = s_lut[v.x & 0xffff] | (s_lut[v.x >> 16] << 16);
    o.y = s_lut[v.y & 0xffff] | (s_lut[v.y >> 16] << 16);
    o.z = s_lut[v.z & 0xffff] | (s_lut[v.z >> 16] << 16);
    o.w = s_lut[v.w & 0xffff] | (s_lut[v.w >> 16] << 16);
    return o;
}

// grid.x = CTB-row strips of 8 rows; each CTA walks 8 rows x the whole picture width
__global__ void __launch_bounds__(kThreads) lmcs_frame_kernel(pel *plane, int pitch, long long bstride, int w, int h,
                                                              int ctb_log2, int ctb_cols, int ctb_rows,
                                                              const uint16_t *lut, int lut_n, const uint8_t *enable)
{
    extern __shared__ unsigned s_lut[];
    stage_lut(s_lut, lut, lut_n);
    const int k = blockIdx.y;
    pel *pic = plane + k * bstride;
    const int y0 = blockIdx.x * 8;
    const int groups = w >> 3;                           // 8-sample groups per row (w is a multiple of 8)
    for (int i = threadIdx.x; i < groups * 8; i += kThreads) {
        const int r = i / groups, g = i - r * groups;
        const int y = y0 + r, x = g << 3;
        if (y >= h)
            break;
        if (enable && !enable[((long long)k * ctb_rows + (y >> ctb_log2)) * ctb_cols + (x >> ctb_log2)])
            continue;
        uint4 *p = reinterpret_cast<uint4 *>(pic + (long long)y * pitch + x);
        *p = map8(*p, s_lut);
    }
}

__global__ void __launch_bounds__(kThreads) lmcs_rects_kernel(pel *plane, int pitch, long long bstride,
                                                              const uint16_t *lut, int lut_n,
                                                              const VVCCudaRect *rects, int n)
{
    extern __shared__ unsigned s_lut[];
    stage_lut(s_lut, lut, lut_n);
    for (int ri = blockIdx.x; ri < n; ri += gridDim.x) {
        const VVCCudaRect r = rects[ri];
        pel *base = plane + r.pic * bstride + (long long)r.y * pitch + r.x;
        const int quads = (r.w + 3) >> 2;
        for (int i = threadIdx.x; i < quads * r.h; i += kThreads) {
            const int y = i / quads, x = (i - y * quads) << 2;
            pel *p = base + (long long)y * pitch + x;
            if (x + 3 < r.w && !((r.x + x) & 3)) {
                uint2 v = *reinterpret_cast<uint2 *>(p);
                v.x = s_lut[v.x & 0xffff] | (s_lut[v.x >> 16] << 16);
                v.y = s_lut[v.y & 0xffff] | (s_lut[v.y >> 16] << 16);
                *reinterpret_cast<uint2 *>(p) = v;
            } else {
                for (int e = 0; e < 4 && x + e < r.w; e++)
                    p[e] = (pel)s_lut[p[e]];
            }
        }
    }
}

// lmcs_derive_chroma_scale (libavcodec/vvc/vvc_intra_template.c:389-428) for a list of VPDUs: one warp per VPDU sums the
// reconstructed luma column left of it and the row above it (lmcs_sum_samples :377-387 repeats the last sample past the
// picture edge, which is what the clamped index does), averages, and searches the pivots.
__global__ void __launch_bounds__(kThreads) lmcs_chroma_scale_kernel(const pel *plane, int pitch, long long bstride, int w, int h,
                                                                     int size, int bd, const VVCCudaLmcsVpdu *vpdus, int n,
                                                                     const VVCCudaLmcsParams *lp, uint16_t *scales)
{
    const int i = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n)
        return;
    const VVCCudaLmcsVpdu v = vpdus[i];
    const pel *pic = plane + v.pic * bstride;
    int sum = 0, cnt = 0;
    if (v.avail_l) {
        for (int j = lane; j < size; j += 32)
            sum += pic[(long long)min(v.y + j, h - 1) * pitch + v.x - 1];
        cnt = size;
    }
    if (v.avail_t) {
        for (int j = lane; j < size; j += 32)
            sum += pic[(long long)(v.y - 1) * pitch + min(v.x + j, w - 1)];
        cnt += size;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane)
        return;
    const int luma = cnt ? (sum + (cnt >> 1)) >> d_ilog2(cnt) : 1 << (bd - 1);
    int k = lp->min_bin_idx;
    while (k <= lp->max_bin_idx && luma >= lp->pivot[k + 1])
        k++;
    scales[i] = lp->chroma_scale_coeff[min(k, 15)];
}

// whole_rows: the CTB kernel walks rows with 128-bit accesses and needs the width to be a multiple of 8;
// the rectangle kernel checks alignment per rectangle
int check(VVCCudaCtx *ctx, const VVCCudaFrame *f, const void *lut, bool whole_rows)
{
    if (!f || !lut)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "lmcs: null argument");
    if ((f->bit_depth != 10 && f->bit_depth != 12) || (whole_rows && (f->width & 7)) || !frame_vec_ok(f))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "lmcs: unsupported picture format");
    return 0;
}

}  // namespace

extern "C" int vvc_cuda_lmcs_frame(VVCCudaCtx *ctx, const VVCCudaFrame *f, const uint16_t *lut, const uint8_t *ctb_enable)
{
    if (ctx->err)
        return ctx->err;
    if (check(ctx, f, lut, true))
        return ctx->err;
    const int n = 1 << f->bit_depth;
    lmcs_frame_kernel<<<dim3(ceil_div(f->height, 8), f->batch), kThreads, n * sizeof(unsigned), ctx->stream>>>(
        (pel *)f->data[0], (int)(f->stride[0] / 2), f->batch_stride[0] / 2, f->width, f->height, f->ctb_log2,
        ceil_div(f->width, 1 << f->ctb_log2), ceil_div(f->height, 1 << f->ctb_log2), lut, n, ctb_enable);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_lmcs_rects(VVCCudaCtx *ctx, const VVCCudaFrame *f, const uint16_t *lut,
                                   const VVCCudaRect *rects, int n_rects)
{
    if (ctx->err)
        return ctx->err;
    if (check(ctx, f, lut, false) || !rects)
        return ctx->err ? ctx->err : vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "lmcs: null rects");
    if (n_rects <= 0)
        return VVC_CUDA_OK;
    const int n = 1 << f->bit_depth;
    const int grid = n_rects < 148 * 8 ? n_rects : 148 * 8;
    lmcs_rects_kernel<<<grid, kThreads, n * sizeof(unsigned), ctx->stream>>>(
        (pel *)f->data[0], (int)(f->stride[0] / 2), f->batch_stride[0] / 2, lut, n, rects, n_rects);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_lmcs_chroma_scale(VVCCudaCtx *ctx, const VVCCudaFrame *f, const VVCCudaLmcsVpdu *vpdus, int n,
                                          const VVCCudaLmcsParams *params, uint16_t *scales)
{
    if (ctx->err)
        return ctx->err;
    if (!f || !vpdus || !params || !scales || n < 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "lmcs_chroma_scale: null argument");
    if (f->bit_depth != 10 && f->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "lmcs_chroma_scale: unsupported picture format");
    if (!n)
        return VVC_CUDA_OK;
    const int ctb = 1 << f->ctb_log2;
    lmcs_chroma_scale_kernel<<<ceil_div(n, kThreads / 32), kThreads, 0, ctx->stream>>>(
        (const pel *)f->data[0], (int)(f->stride[0] / 2), f->batch_stride[0] / 2, f->width, f->height, ctb < 64 ? ctb : 64,
        f->bit_depth, vpdus, n, params, scales);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_lmcs_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *f, const uint16_t *lut, const uint8_t *ctb_enable)
{
    if (ctx->err)
        return ctx->err;
    if (!f || !lut)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "lmcs_host: null argument");
    // only luma travels
    VVCCudaFrame luma = *f;
    luma.chroma_format_idc = 0;
    const int n_ctb = ceil_div(f->width, 1 << f->ctb_log2) * ceil_div(f->height, 1 << f->ctb_log2) * f->batch;
    const size_t fsz = align_up(vvc_stage_frame_size(&luma), 256);
    const size_t lsz = align_up(sizeof(uint16_t) << f->bit_depth, 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, fsz + lsz + align_up(n_ctb, 256));
    if (!base)
        return ctx->err;
    VVCCudaFrame d;
    vvc_stage_frame_layout(&luma, base, &d);
    uint16_t *dlut = (uint16_t *)(base + fsz);
    uint8_t *den = ctb_enable ? base + fsz + lsz : NULL;
    if (vvc_stage_frame_h2d(ctx, &d, &luma))
        return ctx->err;
    VVC_TRY(ctx, cudaMemcpyAsync(dlut, lut, sizeof(uint16_t) << f->bit_depth, cudaMemcpyHostToDevice, ctx->stream));
    if (den)
        VVC_TRY(ctx, cudaMemcpyAsync(den, ctb_enable, n_ctb, cudaMemcpyHostToDevice, ctx->stream));
    if (vvc_cuda_lmcs_frame(ctx, &d, dlut, den))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, &luma, &d))
        return ctx->err;
    return vvc_cuda_sync(ctx);
}
