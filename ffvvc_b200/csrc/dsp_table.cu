// ff_vvc_dsp_init_cuda(): the drop-in override of the reference's VVCDSPContext (include/vvcdsp_table.h).
// Host-pointer entries with the reference's signatures; each one stages its operands into the process-wide
// context's device buffer, runs the same kernels as the batched path, copies the result back and returns.
#include <mutex>
#include <stdlib.h>
#include "common.cuh"
#include "vvcdsp_table.h"

namespace {

std::mutex g_mu;
VVCCudaCtx *g_ctx;
int g_err;
char g_msg[256] = "ok";

void latch(int code, const char *msg)
{
    if (!g_err) {
        g_err = code;
        snprintf(g_msg, sizeof(g_msg), "%s", msg);
    }
}

// called with g_mu held
VVCCudaCtx *table_ctx()
{
    if (!g_ctx && !g_err) {
        const char *dev = getenv("VVC_CUDA_DEVICE");
        const int rc = vvc_cuda_ctx_create(&g_ctx, dev ? atoi(dev) : 0, NULL);
        if (rc)
            latch(rc, "ff_vvc_dsp_init_cuda: no usable CUDA device (there is no CPU fallback behind the CUDA table)");
    }
    return g_ctx;
}

void finish(VVCCudaCtx *ctx)
{
    if (ctx->err) {
        latch(ctx->err, ctx->msg);
        ctx->err = VVC_CUDA_OK;                 // the table keeps working for later, independent calls
    }
}

// A dummy 8x8 luma-only device picture for calls that never touch the picture (STORE_RESIDUAL blocks).
VVCCudaFrame dummy_frame(void *dev, int bd)
{
    VVCCudaFrame f;
    memset(&f, 0, sizeof(f));
    f.width = 8; f.height = 8; f.bit_depth = bd; f.chroma_format_idc = 0; f.ctb_log2 = 7; f.batch = 1;
    f.data[0] = dev; f.stride[0] = 256; f.batch_stride[0] = 256 * 8;
    return f;
}

// ---- itx.itx[trh][trv][log2 w][log2 h] (vvcdsp.c:140-195): in place on row-major int32 coeffs[h][w] ----
void itx_call(int trh, int trv, int lw, int lh, int *coeffs, size_t nzw, size_t nzh, intptr_t range, intptr_t bd)
{
    std::lock_guard<std::mutex> lock(g_mu);
    VVCCudaCtx *ctx = table_ctx();
    if (!ctx)
        return;
    const size_t n = (size_t)1 << (lw + lh);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 4096 + 256 + n * sizeof(int32_t));
    if (!base)
        return finish(ctx);
    VVCCudaTB tb;
    memset(&tb, 0, sizeof(tb));
    tb.log2_w = (uint8_t)lw; tb.log2_h = (uint8_t)lh; tb.trh = (uint8_t)trh; tb.trv = (uint8_t)trv;
    tb.nzw = (uint8_t)nzw; tb.nzh = (uint8_t)nzh; tb.flags = VVC_CUDA_TB_STORE_RESIDUAL;
    const VVCCudaFrame f = dummy_frame(base, (int)bd);
    VVCCudaTB *dtb = (VVCCudaTB *)(base + 4096);
    int32_t *dco = (int32_t *)(base + 4096 + 256);
    cudaMemcpyAsync(dtb, &tb, sizeof(tb), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dco, coeffs, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (!vvc_cuda_itx_frame(ctx, &f, dco, dtb, 1, (int)range)) {
        cudaMemcpyAsync(coeffs, dco, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        vvc_cuda_sync(ctx);
    }
    finish(ctx);
}

template <int TRH, int TRV, int LW, int LH>
void itx_cell(int *coeffs, size_t nzw, size_t nzh, intptr_t range, intptr_t bd)
{
    itx_call(TRH, TRV, LW, LH, coeffs, nzw, nzh, range, bd);
}

// the cells the reference installs: DCT2 for 2..64, DST7 / DCT8 for 4..32 in their dimension; 1-D cells are
// 16 or more long (DST7 / DCT8: 16, 32) and carry DCT2 as the type of their unit dimension
bool cell_valid(int trh, int trv, int lw, int lh)
{
    if (!lw && !lh)
        return false;
    if (!lw || !lh) {
        const int l = lw + lh, tr = lw ? trh : trv, other = lw ? trv : trh;
        return other == 0 && l >= 4 && (tr == 0 || l <= 5);
    }
    if (trh && (lw < 2 || lw > 5)) return false;
    if (trv && (lh < 2 || lh > 5)) return false;
    return true;
}

template <int TRH, int TRV, int LW, int LH>
void install_cell(VVCItxDSPContext *t)
{
    if (cell_valid(TRH, TRV, LW, LH))
        t->itx[TRH][TRV][LW][LH] = itx_cell<TRH, TRV, LW, LH>;
}
template <int TRH, int TRV, int LW>
void install_row(VVCItxDSPContext *t)
{
    install_cell<TRH, TRV, LW, 0>(t); install_cell<TRH, TRV, LW, 1>(t); install_cell<TRH, TRV, LW, 2>(t);
    install_cell<TRH, TRV, LW, 3>(t); install_cell<TRH, TRV, LW, 4>(t); install_cell<TRH, TRV, LW, 5>(t);
    install_cell<TRH, TRV, LW, 6>(t);
}
template <int TRH, int TRV>
void install_type(VVCItxDSPContext *t)
{
    install_row<TRH, TRV, 0>(t); install_row<TRH, TRV, 1>(t); install_row<TRH, TRV, 2>(t); install_row<TRH, TRV, 3>(t);
    install_row<TRH, TRV, 4>(t); install_row<TRH, TRV, 5>(t); install_row<TRH, TRV, 6>(t);
}

int ilog2_exact(int v)
{
    int l = 0;
    while ((1 << l) < v)
        l++;
    return (1 << l) == v ? l : -1;
}

// ---- itx.transform_bdpcm (vvcdsp_template.c:76-95): in place ----
void bdpcm_entry(int *coeffs, int width, int height, int vertical, int log2_transform_range)
{
    std::lock_guard<std::mutex> lock(g_mu);
    VVCCudaCtx *ctx = table_ctx();
    if (!ctx)
        return;
    const int lw = ilog2_exact(width), lh = ilog2_exact(height);
    if (lw < 0 || lh < 0 || lw > 6 || lh > 6)
        return latch(VVC_CUDA_ERR_ARG, "transform_bdpcm: block size is not a power of two <= 64");
    const size_t n = (size_t)width * height;
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 4096 + 256 + n * sizeof(int32_t));
    if (!base)
        return finish(ctx);
    VVCCudaTB tb;
    memset(&tb, 0, sizeof(tb));
    tb.log2_w = (uint8_t)lw; tb.log2_h = (uint8_t)lh; tb.nzw = (uint8_t)(width > 32 ? 32 : width); tb.nzh = (uint8_t)(height > 32 ? 32 : height);
    tb.flags = VVC_CUDA_TB_STORE_RESIDUAL | VVC_CUDA_TB_TS | (vertical ? VVC_CUDA_TB_BDPCM_VERT : VVC_CUDA_TB_BDPCM);
    const VVCCudaFrame f = dummy_frame(base, 10);
    VVCCudaTB *dtb = (VVCCudaTB *)(base + 4096);
    int32_t *dco = (int32_t *)(base + 4096 + 256);
    cudaMemcpyAsync(dtb, &tb, sizeof(tb), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dco, coeffs, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (!vvc_cuda_itx_frame(ctx, &f, dco, dtb, 1, log2_transform_range)) {
        cudaMemcpyAsync(coeffs, dco, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
        vvc_cuda_sync(ctx);
    }
    finish(ctx);
}

// ---- itx.add_residual (vvcdsp_template.c:32-46): dst = clip_pixel(dst + res), 10-bit table ----
void add_residual_entry(uint8_t *dst, const int *res, int width, int height, ptrdiff_t stride)
{
    std::lock_guard<std::mutex> lock(g_mu);
    VVCCudaCtx *ctx = table_ctx();
    if (!ctx)
        return;
    const int lw = ilog2_exact(width), lh = ilog2_exact(height);
    if (lw < 0 || lh < 0 || lw > 6 || lh > 6)
        return latch(VVC_CUDA_ERR_ARG, "add_residual: block size is not a power of two <= 64");
    // the block as a one-plane host picture; the residual as a transform-skip block (no LFNST / transform)
    VVCCudaFrame f;
    memset(&f, 0, sizeof(f));
    f.width = width; f.height = height; f.bit_depth = 10; f.chroma_format_idc = 0; f.ctb_log2 = 7; f.batch = 1;
    f.data[0] = dst; f.stride[0] = stride; f.batch_stride[0] = stride * height;
    VVCCudaTB tb;
    memset(&tb, 0, sizeof(tb));
    tb.log2_w = (uint8_t)lw; tb.log2_h = (uint8_t)lh; tb.nzw = (uint8_t)(width > 32 ? 32 : width); tb.nzh = (uint8_t)(height > 32 ? 32 : height);
    tb.flags = VVC_CUDA_TB_TS;
    vvc_cuda_itx_frame_host(ctx, &f, const_cast<int *>(res), (size_t)width * height, &tb, 1, 15);
    finish(ctx);
}

// ---- lmcs.filter (vvc_filter_template.c:25-36): dst[x] = lut[dst[x]] in place, any block size ----
void lmcs_entry(uint8_t *dst, ptrdiff_t dst_stride, int width, int height, const uint8_t *lut)
{
    std::lock_guard<std::mutex> lock(g_mu);
    VVCCudaCtx *ctx = table_ctx();
    if (!ctx)
        return;
    VVCCudaFrame f, d;
    memset(&f, 0, sizeof(f));
    f.width = width; f.height = height; f.bit_depth = 10; f.chroma_format_idc = 0; f.ctb_log2 = 7; f.batch = 1;
    f.data[0] = dst; f.stride[0] = dst_stride; f.batch_stride[0] = dst_stride * height;
    const size_t fsz = align_up(vvc_stage_frame_size(&f), 256), lsz = sizeof(uint16_t) << 10;
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, fsz + lsz + 256);
    if (!base)
        return finish(ctx);
    vvc_stage_frame_layout(&f, base, &d);
    uint16_t *dlut = (uint16_t *)(base + fsz);
    VVCCudaRect *drect = (VVCCudaRect *)(base + fsz + lsz);
    VVCCudaRect r;
    memset(&r, 0, sizeof(r));
    r.w = (uint16_t)width; r.h = (uint16_t)height;
    if (!vvc_stage_frame_h2d(ctx, &d, &f)) {
        cudaMemcpyAsync(dlut, lut, lsz, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(drect, &r, sizeof(r), cudaMemcpyHostToDevice, ctx->stream);
        if (!vvc_cuda_lmcs_rects(ctx, &d, dlut, drect, 1) && !vvc_stage_frame_d2h(ctx, &f, &d))
            vvc_cuda_sync(ctx);
    }
    finish(ctx);
}

}  // namespace

extern "C" void ff_vvc_dsp_init_cuda(VVCDSPContext *c, int bit_depth)
{
    if (!c || bit_depth != 10)
        return;
    install_type<0, 0>(&c->itx); install_type<0, 1>(&c->itx); install_type<0, 2>(&c->itx);
    install_type<1, 0>(&c->itx); install_type<1, 1>(&c->itx); install_type<1, 2>(&c->itx);
    install_type<2, 0>(&c->itx); install_type<2, 1>(&c->itx); install_type<2, 2>(&c->itx);
    c->itx.transform_bdpcm = bdpcm_entry;
    c->itx.add_residual = add_residual_entry;
    c->lmcs.filter = lmcs_entry;
}

extern "C" int ff_vvc_dsp_cuda_last_error(void) { std::lock_guard<std::mutex> lock(g_mu); return g_err; }
extern "C" const char *ff_vvc_dsp_cuda_error_string(void) { return g_msg; }
extern "C" void ff_vvc_dsp_cuda_reset_error(void) { std::lock_guard<std::mutex> lock(g_mu); g_err = 0; snprintf(g_msg, sizeof(g_msg), "ok"); }
extern "C" size_t ff_vvc_dsp_cuda_sizeof_table(void) { return sizeof(VVCDSPContext); }
