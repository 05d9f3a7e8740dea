// ff_vvc_dsp_init_cuda(): the drop-in override of the reference's VVCDSPContext (include/vvcdsp_table.h).
// Host-pointer entries with the reference's signatures; each one stages its operands into the process-wide
// context's device buffer, runs the same kernels as the batched path, copies the result back and returns.
#include <mutex>
#include <stdlib.h>
#include "common.cuh"
#include "dsp_block.cuh"
#include "vvcdsp_table.h"

namespace {

std::mutex g_mu;
VVCCudaCtx *g_ctx;
int g_err;
char g_msg[256] = "ok";

// The entries return void like the reference's, so a failure cannot travel up the call: it is latched for
// ff_vvc_dsp_cuda_last_error(), reported once on stderr, and - with VVC_CUDA_TABLE_ABORT=1 in the environment - aborts the
// process, so that a decoder cannot run on over blocks that were never reconstructed.
void latch(int code, const char *msg)
{
    if (!g_err) {
        g_err = code;
        snprintf(g_msg, sizeof(g_msg), "%s", msg);
        fprintf(stderr, "CUDA table error %d: %s\n", code, g_msg);
        const char *ab = getenv("VVC_CUDA_TABLE_ABORT");
        if (ab && atoi(ab))
            abort();
    }
}

// a staged copy of a table entry: failures go through the context's sticky error (and from there into the latch)
void copy(VVCCudaCtx *ctx, void *dst, const void *src, size_t bytes, cudaMemcpyKind kind)
{
    if (!ctx->err)
        vvc_ctx_check(ctx, cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream), "cudaMemcpyAsync (table entry)");
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
    copy(ctx, dtb, &tb, sizeof(tb), cudaMemcpyHostToDevice);
    copy(ctx, dco, coeffs, n * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (!vvc_cuda_itx_frame(ctx, &f, dco, dtb, 1, (int)range)) {
        copy(ctx, coeffs, dco, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
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
    copy(ctx, dtb, &tb, sizeof(tb), cudaMemcpyHostToDevice);
    copy(ctx, dco, coeffs, n * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (!vvc_cuda_itx_frame(ctx, &f, dco, dtb, 1, log2_transform_range)) {
        copy(ctx, coeffs, dco, n * sizeof(int32_t), cudaMemcpyDeviceToHost);
        vvc_cuda_sync(ctx);
    }
    finish(ctx);
}

// ---- itx.add_residual (vvcdsp_template.c:32-46): dst = clip_pixel(dst + res), 10-bit table ----
template <int BD>
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
    f.width = width; f.height = height; f.bit_depth = BD; f.chroma_format_idc = 0; f.ctb_log2 = 7; f.batch = 1;
    f.data[0] = dst; f.stride[0] = stride; f.batch_stride[0] = stride * height;
    VVCCudaTB tb;
    memset(&tb, 0, sizeof(tb));
    tb.log2_w = (uint8_t)lw; tb.log2_h = (uint8_t)lh; tb.nzw = (uint8_t)(width > 32 ? 32 : width); tb.nzh = (uint8_t)(height > 32 ? 32 : height);
    tb.flags = VVC_CUDA_TB_TS;
    vvc_cuda_itx_frame_host(ctx, &f, const_cast<int *>(res), (size_t)width * height, &tb, 1, 15);
    finish(ctx);
}

// ---- lmcs.filter (vvc_filter_template.c:25-36): dst[x] = lut[dst[x]] in place, any block size ----
template <int BD>
void lmcs_entry(uint8_t *dst, ptrdiff_t dst_stride, int width, int height, const uint8_t *lut)
{
    std::lock_guard<std::mutex> lock(g_mu);
    VVCCudaCtx *ctx = table_ctx();
    if (!ctx)
        return;
    VVCCudaFrame f, d;
    memset(&f, 0, sizeof(f));
    f.width = width; f.height = height; f.bit_depth = BD; f.chroma_format_idc = 0; f.ctb_log2 = 7; f.batch = 1;
    f.data[0] = dst; f.stride[0] = dst_stride; f.batch_stride[0] = dst_stride * height;
    const size_t fsz = align_up(vvc_stage_frame_size(&f), 256), lsz = sizeof(uint16_t) << BD;
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
        copy(ctx, dlut, lut, lsz, cudaMemcpyHostToDevice);
        copy(ctx, drect, &r, sizeof(r), cudaMemcpyHostToDevice);
        if (!vvc_cuda_lmcs_rects(ctx, &d, dlut, drect, 1) && !vvc_stage_frame_d2h(ctx, &f, &d))
            vvc_cuda_sync(ctx);
    }
    finish(ctx);
}


// =====================================================================================================================
// Per-call shims of the remaining tables (inter, sao, alf, lf, joint residuals).  Each one stages exactly the samples
// the reference's C entry reads (no more: the caller's buffers end where the reference stops reading), launches the
// block kernel of dsp_block.cu and copies back exactly the samples the C entry writes.
// =====================================================================================================================
struct Call {
    std::unique_lock<std::mutex> lk;
    VVCCudaCtx *ctx;
    uint8_t *base;
    size_t off;
    bool bad;
    explicit Call(size_t bytes) : lk(g_mu), ctx(table_ctx()), base(nullptr), off(0), bad(false)
    {
        if (ctx) {
            base = (uint8_t *)vvc_ctx_dev_stage(ctx, bytes + 16 * 256);
            if (!base) { finish(ctx); bad = true; }
        } else {
            bad = true;
        }
    }
    template <class T> T *take(size_t n) { T *p = (T *)(base + off); off += align_up(n * sizeof(T), 256); return p; }
    void chk(cudaError_t e, const char *what) { if (!bad && vvc_ctx_check(ctx, e, what)) bad = true; }
    void up(void *d, const void *h, size_t n) { chk(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, ctx->stream), "table shim H2D"); }
    void dn(void *h, const void *d, size_t n) { chk(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, ctx->stream), "table shim D2H"); }
    void up2d(void *d, size_t dp, const void *h, size_t hp, size_t wb, size_t rows)
    { chk(cudaMemcpy2DAsync(d, dp, h, hp, wb, rows, cudaMemcpyHostToDevice, ctx->stream), "table shim H2D 2-D"); }
    void dn2d(void *h, size_t hp, const void *d, size_t dp, size_t wb, size_t rows)
    { chk(cudaMemcpy2DAsync(h, hp, d, dp, wb, rows, cudaMemcpyDeviceToHost, ctx->stream), "table shim D2H 2-D"); }
    void launched() { ctx->launches++; chk(cudaGetLastError(), "table shim kernel launch"); }
    void done() { if (ctx) { vvc_cuda_sync(ctx); finish(ctx); } }
    cudaStream_t st() const { return ctx->stream; }
};

constexpr int PB = 128;            // MAX_PB_SIZE

// ---- inter.put / put_uni / put_uni_w [luma, chroma][log2 w - 1][vertical frac][horizontal frac] ----
template <int BD, int CH, int VF, int HF, int MODE>
void mc_call(int16_t *dst16, uint8_t *dst, ptrdiff_t dst_stride, const uint8_t *src, ptrdiff_t src_stride, int height,
             int denom, int wx, int ox, const int8_t *hf, const int8_t *vf, int width)
{
    const int taps = CH ? 4 : 8, before = taps / 2 - 1, after = taps / 2;
    const int x0 = HF ? before : 0, x1 = HF ? after : 0, y0 = VF ? before : 0, y1 = VF ? after : 0;
    const int ww = width + x0 + x1, wh = height + y0 + y1;
    // the window as ONE byte range with the caller's own stride: the reference's loops address src + y * stride + x and
    // work with any stride, also one smaller than the window is wide (checkasm passes 128-sample rows for 128 + 7 columns)
    const size_t span = (size_t)(wh - 1) * src_stride + (size_t)ww * 2;
    const int ss = (int)(src_stride / 2);
    Call c(span + 256 + (size_t)height * PB * 2 + (size_t)width * height * 2);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>(span / 2 + 1);
    c.up(dsrc, src - y0 * src_stride - x0 * 2, span);
    vvcblk::McArgs a;
    memset(&a, 0, sizeof(a));
    a.src = dsrc + y0 * ss + x0; a.sstride = ss; a.w = width; a.h = height; a.taps = taps; a.hfrac = HF; a.vfrac = VF; a.mode = MODE;
    memcpy(a.hf, hf, taps); memcpy(a.vf, vf, taps);
    a.denom = denom; a.wx = wx; a.ox = ox;
    if (MODE == 0) {
        a.dst16 = c.take<int16_t>((size_t)height * PB);
        vvcblk::mc(c.st(), BD, a); c.launched();
        c.dn2d(dst16, PB * 2, a.dst16, PB * 2, width * 2, height);
    } else {
        a.dst = c.take<pel>((size_t)width * height); a.dstride = width;
        vvcblk::mc(c.st(), BD, a); c.launched();
        c.dn2d(dst, dst_stride, a.dst, width * 2, width * 2, height);
    }
    c.done();
}
template <int BD, int CH, int VF, int HF>
void put_entry(int16_t *dst, const uint8_t *src, ptrdiff_t ss, int h, const int8_t *hf, const int8_t *vf, int w)
{ mc_call<BD, CH, VF, HF, 0>(dst, nullptr, 0, src, ss, h, 0, 0, 0, hf, vf, w); }
template <int BD, int CH, int VF, int HF>
void put_uni_entry(uint8_t *dst, ptrdiff_t ds, const uint8_t *src, ptrdiff_t ss, int h, const int8_t *hf, const int8_t *vf, int w)
{ mc_call<BD, CH, VF, HF, 1>(nullptr, dst, ds, src, ss, h, 0, 0, 0, hf, vf, w); }
template <int BD, int CH, int VF, int HF>
void put_uni_w_entry(uint8_t *dst, ptrdiff_t ds, const uint8_t *src, ptrdiff_t ss, int h, int denom, int wx, int ox,
                     const int8_t *hf, const int8_t *vf, int w)
{ mc_call<BD, CH, VF, HF, 2>(nullptr, dst, ds, src, ss, h, denom, wx, ox, hf, vf, w); }

// ---- avg / w_avg / put_gpm / put_ciip ----
template <int BD>
void blend_call(int mode, uint8_t *dst, ptrdiff_t ds, const int16_t *s0, const int16_t *s1, int w, int h, int denom, int w0, int w1,
                int o0, int o1, const uint8_t *weights, int step_x, int step_y, const uint8_t *inter, ptrdiff_t is)
{
    const size_t tile = (size_t)h * PB;
    // GPM weights: the entry reads weights[y * step_y + x * step_x]; steps may be negative
    long long lo = 0, hi = 0;
    if (mode == 2) {
        const long long cx = (long long)(w - 1) * step_x, cy = (long long)(h - 1) * step_y;
        lo = (cx < 0 ? cx : 0) + (cy < 0 ? cy : 0); hi = (cx > 0 ? cx : 0) + (cy > 0 ? cy : 0);
    }
    Call c(2 * tile * 2 + (size_t)w * h * 4 + (size_t)(hi - lo + 1));
    if (c.bad) return c.done();
    vvcblk::BlendArgs a;
    memset(&a, 0, sizeof(a));
    a.w = w; a.h = h; a.mode = mode; a.denom = denom; a.w0 = w0; a.w1 = w1; a.o0 = o0; a.o1 = o1;
    pel *ddst = c.take<pel>((size_t)w * h);
    a.dst = ddst; a.dstride = w;
    if (mode == 3) {
        pel *dint = c.take<pel>((size_t)w * h);
        c.up2d(ddst, w * 2, dst, ds, w * 2, h);
        c.up2d(dint, w * 2, inter, is, w * 2, h);
        a.inter = dint; a.istride = w;
    } else {
        int16_t *d0 = c.take<int16_t>(tile), *d1 = c.take<int16_t>(tile);
        c.up2d(d0, PB * 2, s0, PB * 2, w * 2, h);
        c.up2d(d1, PB * 2, s1, PB * 2, w * 2, h);
        a.src0 = d0; a.src1 = d1;
        if (mode == 2) {
            uint8_t *dw = c.take<uint8_t>((size_t)(hi - lo + 1));
            c.up(dw, weights + lo, (size_t)(hi - lo + 1));
            a.weights = dw - lo; a.step_x = step_x; a.step_y = step_y;
        }
    }
    vvcblk::blend(c.st(), BD, a); c.launched();
    c.dn2d(dst, ds, ddst, w * 2, w * 2, h);
    c.done();
}
template <int BD> void avg_entry(uint8_t *dst, ptrdiff_t ds, const int16_t *s0, const int16_t *s1, int w, int h)
{ blend_call<BD>(0, dst, ds, s0, s1, w, h, 0, 0, 0, 0, 0, nullptr, 0, 0, nullptr, 0); }
template <int BD> void w_avg_entry(uint8_t *dst, ptrdiff_t ds, const int16_t *s0, const int16_t *s1, int w, int h, int denom, int w0, int w1, int o0, int o1)
{ blend_call<BD>(1, dst, ds, s0, s1, w, h, denom, w0, w1, o0, o1, nullptr, 0, 0, nullptr, 0); }
template <int BD> void put_gpm_entry(uint8_t *dst, ptrdiff_t ds, int w, int h, const int16_t *s0, const int16_t *s1, const uint8_t *weights, int sx, int sy)
{ blend_call<BD>(2, dst, ds, s0, s1, w, h, 0, 0, 0, 0, 0, weights, sx, sy, nullptr, 0); }
template <int BD> void put_ciip_entry(uint8_t *dst, ptrdiff_t ds, int w, int h, const uint8_t *inter, ptrdiff_t is, int intra_weight)
{ blend_call<BD>(3, dst, ds, nullptr, nullptr, w, h, 0, intra_weight, 0, 0, 0, nullptr, 0, 0, inter, is); }

// ---- dmvr[my != 0][mx != 0], sad ----
template <int BD, int VF, int HF>
void dmvr_entry(int16_t *dst, const uint8_t *src, ptrdiff_t ss, int height, intptr_t mx, intptr_t my, int width)
{
    const int ww = width + HF, wh = height + VF;
    Call c((size_t)ww * wh * 2 + (size_t)height * PB * 2);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>((size_t)ww * wh);
    int16_t *dd = c.take<int16_t>((size_t)height * PB);
    c.up2d(dsrc, ww * 2, src, ss, ww * 2, wh);
    vvcblk::dmvr(c.st(), BD, dd, dsrc, ww, height, width, HF ? (int)mx : 0, VF ? (int)my : 0); c.launched();
    c.dn2d(dst, PB * 2, dd, PB * 2, width * 2, height);
    c.done();
}
int sad_entry(const int16_t *src0, const int16_t *src1, int dx, int dy, int bw, int bh)
{
    // both tiles are (bw + 4) x (bh + 4) predictions at pitch 128
    const size_t tile = (size_t)(bh + 4) * PB;
    Call c(2 * tile * 2 + 256);
    int out = 0;
    if (c.bad) { c.done(); return 0; }
    int16_t *d0 = c.take<int16_t>(tile), *d1 = c.take<int16_t>(tile);
    int *dout = c.take<int>(1);
    c.up2d(d0, PB * 2, src0, PB * 2, (bw + 4) * 2, bh + 4);
    c.up2d(d1, PB * 2, src1, PB * 2, (bw + 4) * 2, bh + 4);
    vvcblk::sad(c.st(), dout, d0, d1, dx, dy, bw, bh); c.launched();
    c.dn(&out, dout, sizeof(int));
    c.done();
    return out;
}

// ---- fetch_samples / bdof_fetch_samples / prof_grad_filter / apply_prof* / apply_bdof ----
template <int BD>
void bdof_fetch_entry(int16_t *dst, const uint8_t *src, ptrdiff_t ss, int x_frac, int y_frac, int width, int height)
{
    const int w2 = width + 2, h2 = height + 2;
    Call c((size_t)w2 * h2 * 2 + (size_t)(h2 + 1) * PB * 2);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>((size_t)w2 * h2);
    int16_t *dt = c.take<int16_t>((size_t)(h2 + 1) * PB);          // one spare row in front: the ring starts at (-1, -1)
    const uint8_t *s0 = src + ((y_frac >> 3) - 1) * ss + ((x_frac >> 3) - 1) * 2;
    c.up2d(dsrc, w2 * 2, s0, ss, w2 * 2, h2);
    int16_t *org = dt + PB + 1;                                    // tile origin inside the staging tile
    c.up2d(org - PB - 1, PB * 2, dst - PB - 1, PB * 2, w2 * 2, h2); // the entry only writes the ring: keep the interior
    vvcblk::fetch(c.st(), BD, org, dsrc, w2, width, height); c.launched();
    c.dn2d(dst - PB - 1, PB * 2, org - PB - 1, PB * 2, w2 * 2, h2);
    c.done();
}
template <int BD>
void fetch_entry(int16_t *dst, const uint8_t *src, ptrdiff_t ss, int x_frac, int y_frac)
{ bdof_fetch_entry<BD>(dst, src, ss, x_frac, y_frac, 4, 4); }

void prof_grad_entry(int16_t *gh, int16_t *gv, ptrdiff_t gs, const int16_t *src, ptrdiff_t ss, int width, int height, int pad)
{
    // src is read one sample beyond the block on every side; with pad the gradients get a one-sample border
    const int gw = width + 2 * pad, ghh = height + 2 * pad;
    Call c((size_t)(height + 2) * (width + 2) * 2 + 2 * (size_t)gw * ghh * 2);
    if (c.bad) return c.done();
    int16_t *dsrc = c.take<int16_t>((size_t)(height + 2) * (width + 2));
    int16_t *dh = c.take<int16_t>((size_t)gw * ghh), *dv = c.take<int16_t>((size_t)gw * ghh);
    c.up2d(dsrc, (width + 2) * 2, src - ss - 1, ss * 2, (width + 2) * 2, height + 2);
    vvcblk::prof_grad(c.st(), dh, dv, gw, dsrc + (width + 2) + 1, width + 2, width, height, pad); c.launched();
    c.dn2d(gh, gs * 2, dh, gw * 2, gw * 2, ghh);
    c.dn2d(gv, gs * 2, dv, gw * 2, gw * 2, ghh);
    c.done();
}

template <int BD>
void prof_call(int mode, int16_t *dst16, uint8_t *dst, ptrdiff_t ds, const int16_t *src, const int16_t *dmx, const int16_t *dmy,
               int denom, int wx, int ox)
{
    Call c(6 * PB * 2 + 1024);
    if (c.bad) return c.done();
    int16_t *dt = c.take<int16_t>(6 * PB), *ddx = c.take<int16_t>(16), *ddy = c.take<int16_t>(16);
    c.up2d(dt, PB * 2, src - PB - 1, PB * 2, 6 * 2, 6);
    c.up(ddx, dmx, 32); c.up(ddy, dmy, 32);
    vvcblk::ProfArgs a;
    memset(&a, 0, sizeof(a));
    a.src = dt + PB + 1; a.dx = ddx; a.dy = ddy; a.mode = mode; a.denom = denom; a.wx = wx; a.ox = ox;
    if (mode == 0) {
        a.dst16 = c.take<int16_t>(4 * PB);
        vvcblk::prof(c.st(), BD, a); c.launched();
        c.dn2d(dst16, PB * 2, a.dst16, PB * 2, 4 * 2, 4);
    } else {
        a.dst = c.take<pel>(16); a.dstride = 4;
        vvcblk::prof(c.st(), BD, a); c.launched();
        c.dn2d(dst, ds, a.dst, 8, 8, 4);
    }
    c.done();
}
template <int BD> void apply_prof_entry(int16_t *dst, const int16_t *src, const int16_t *dx, const int16_t *dy)
{ prof_call<BD>(0, dst, nullptr, 0, src, dx, dy, 0, 0, 0); }
template <int BD> void apply_prof_uni_entry(uint8_t *dst, ptrdiff_t ds, const int16_t *src, const int16_t *dx, const int16_t *dy)
{ prof_call<BD>(1, nullptr, dst, ds, src, dx, dy, 0, 0, 0); }
template <int BD> void apply_prof_uni_w_entry(uint8_t *dst, ptrdiff_t ds, const int16_t *src, const int16_t *dx, const int16_t *dy, int denom, int wx, int ox)
{ prof_call<BD>(2, nullptr, dst, ds, src, dx, dy, denom, wx, ox); }

template <int BD>
void apply_bdof_entry(uint8_t *dst, ptrdiff_t ds, int16_t *src0, int16_t *src1, int bw, int bh)
{
    // the tiles come with their ring of integer samples (bdof_fetch_samples) and get their borders padded in place
    const size_t tile = (size_t)(bh + 2) * PB + 2;
    Call c(2 * tile * 2 + (size_t)bw * bh * 2);
    if (c.bad) return c.done();
    int16_t *d0 = c.take<int16_t>(tile), *d1 = c.take<int16_t>(tile);
    pel *dd = c.take<pel>((size_t)bw * bh);
    c.up2d(d0, PB * 2, src0 - PB - 1, PB * 2, (bw + 2) * 2, bh + 2);
    c.up2d(d1, PB * 2, src1 - PB - 1, PB * 2, (bw + 2) * 2, bh + 2);
    vvcblk::bdof(c.st(), BD, dd, bw, d0 + PB + 1, d1 + PB + 1, bw, bh); c.launched();
    c.dn2d(dst, ds, dd, bw * 2, bw * 2, bh);
    c.dn2d(src0 - PB - 1, PB * 2, d0, PB * 2, (bw + 2) * 2, bh + 2);      // pad_int16 modified the callers' tiles
    c.dn2d(src1 - PB - 1, PB * 2, d1, PB * 2, (bw + 2) * 2, bh + 2);
    c.done();
}

// ---- sao.band_filter[9] / edge_filter[9] ----
template <int BD>
void sao_band_entry(uint8_t *dst, const uint8_t *src, ptrdiff_t ds, ptrdiff_t ss, const int16_t *offset_val, int left_class, int width, int height)
{
    Call c((size_t)width * height * 4);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>((size_t)width * height), *dd = c.take<pel>((size_t)width * height);
    c.up2d(dsrc, width * 2, src, ss, width * 2, height);
    vvcblk::SaoArgs a;
    memset(&a, 0, sizeof(a));
    a.dst = dd; a.dstride = width; a.src = dsrc; a.sstride = width; a.w = width; a.h = height; a.left_class = left_class;
    memcpy(a.offset_val, offset_val, sizeof(a.offset_val));
    vvcblk::sao(c.st(), BD, a); c.launched();
    c.dn2d(dst, ds, dd, width * 2, width * 2, height);
    c.done();
}
template <int BD>
void sao_edge_entry(uint8_t *dst, const uint8_t *src, ptrdiff_t ds, const int16_t *offset_val, int eo, int width, int height)
{
    const ptrdiff_t ss = 2 * PB + 64;                       // the entry's implicit source stride: (2 * MAX_PB_SIZE + AV_INPUT_BUFFER_PADDING_SIZE) BYTES (h2656_sao_template.c:64)
    const int ww = width + 2, wh = height + 2;
    Call c((size_t)ww * wh * 2 + (size_t)width * height * 2);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>((size_t)ww * wh), *dd = c.take<pel>((size_t)width * height);
    c.up2d(dsrc, ww * 2, src - ss - 2, ss, ww * 2, wh);
    vvcblk::SaoArgs a;
    memset(&a, 0, sizeof(a));
    a.dst = dd; a.dstride = width; a.src = dsrc + ww + 1; a.sstride = ww; a.w = width; a.h = height; a.edge = 1; a.eo = eo;
    memcpy(a.offset_val, offset_val, sizeof(a.offset_val));
    vvcblk::sao(c.st(), BD, a); c.launched();
    c.dn2d(dst, ds, dd, width * 2, width * 2, height);
    c.done();
}

// ---- alf.filter[2] / filter_cc / classify / recon_coeff_and_clip ----
template <int BD, int CHROMA>
void alf_filter_entry(uint8_t *dst, ptrdiff_t ds, const uint8_t *src, ptrdiff_t ss, int width, int height, const int16_t *filter,
                      const int16_t *clip, int vb_pos)
{
    const int r = CHROMA ? 2 : 3, ww = width + 2 * r, wh = height + 2 * r;
    const size_t np = CHROMA ? 6 : (size_t)(width >> 2) * (height >> 2) * 12;
    Call c((size_t)ww * wh * 2 + (size_t)width * height * 2 + np * 4);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>((size_t)ww * wh), *dd = c.take<pel>((size_t)width * height);
    int16_t *df = c.take<int16_t>(np), *dc = c.take<int16_t>(np);
    c.up2d(dsrc, ww * 2, src - r * ss - r * 2, ss, ww * 2, wh);
    c.up(df, filter, np * 2); c.up(dc, clip, np * 2);
    vvcblk::AlfArgs a;
    a.dst = dd; a.dstride = width; a.src = dsrc + r * ww + r; a.sstride = ww; a.w = width; a.h = height; a.chroma = CHROMA; a.vb_pos = vb_pos;
    a.filter = df; a.clip = dc;
    vvcblk::alf_filter(c.st(), BD, a); c.launched();
    c.dn2d(dst, ds, dd, width * 2, width * 2, height);
    c.done();
}
template <int BD>
void alf_cc_entry(uint8_t *dst, ptrdiff_t ds, const uint8_t *luma, ptrdiff_t ls, int width, int height, int hs, int vs, const int16_t *filter, int vb_pos)
{
    const int lw = ((width - 1) << hs) + 3, lh = ((height - 1) << vs) + 4;   // luma rows -1 .. +2, columns -1 .. +1 around the co-located samples
    Call c((size_t)lw * lh * 2 + (size_t)width * height * 2 + 256);
    if (c.bad) return c.done();
    pel *dl = c.take<pel>((size_t)lw * lh), *dd = c.take<pel>((size_t)width * height);
    int16_t *df = c.take<int16_t>(7);
    c.up2d(dl, lw * 2, luma - ls - 2, ls, lw * 2, lh);
    c.up2d(dd, width * 2, dst, ds, width * 2, height);
    c.up(df, filter, 14);
    vvcblk::alf_cc(c.st(), BD, dd, width, dl + lw + 1, lw, width, height, hs, vs, df, vb_pos); c.launched();
    c.dn2d(dst, ds, dd, width * 2, width * 2, height);
    c.done();
}
template <int BD>
void alf_classify_entry(int *class_idx, int *transpose_idx, const uint8_t *src, ptrdiff_t ss, int width, int height, int vb_pos, int *gradient_tmp)
{
    (void)gradient_tmp;                                            // the kernel keeps its Laplacians in registers
    const int ww = width + 6, wh = height + 6, nb = (width >> 2) * (height >> 2);
    Call c((size_t)ww * wh * 2 + (size_t)nb * 8);
    if (c.bad) return c.done();
    pel *dsrc = c.take<pel>((size_t)ww * wh);
    int *dc = c.take<int>(nb), *dt = c.take<int>(nb);
    c.up2d(dsrc, ww * 2, src - 3 * ss - 6, ss, ww * 2, wh);
    vvcblk::alf_classify(c.st(), BD, dc, dt, dsrc + 3 * ww + 3, ww, width, height, vb_pos); c.launched();
    c.dn(class_idx, dc, (size_t)nb * 4); c.dn(transpose_idx, dt, (size_t)nb * 4);
    c.done();
}
template <int BD>
void alf_recon_entry(int16_t *coeff, int16_t *clip, const int *class_idx, const int *transpose_idx, int size, const int16_t *coeff_set,
                     const uint8_t *clip_idx_set, const uint8_t *class_to_filt)
{
    Call c((size_t)size * (12 * 4 + 8) + 25 * 12 * 3 + 1024);
    if (c.bad) return c.done();
    int16_t *dco = c.take<int16_t>((size_t)size * 12), *dcl = c.take<int16_t>((size_t)size * 12);
    int *dc = c.take<int>(size), *dt = c.take<int>(size);
    int16_t *dset = c.take<int16_t>(25 * 12);
    uint8_t *dcs = c.take<uint8_t>(25 * 12), *dmap = c.take<uint8_t>(25);
    c.up(dc, class_idx, (size_t)size * 4); c.up(dt, transpose_idx, (size_t)size * 4);
    c.up(dset, coeff_set, 25 * 12 * 2); c.up(dcs, clip_idx_set, 25 * 12); c.up(dmap, class_to_filt, 25);
    vvcblk::alf_recon(c.st(), BD, dco, dcl, dc, dt, size, dset, dcs, dmap); c.launched();
    c.dn(coeff, dco, (size_t)size * 24); c.dn(clip, dcl, (size_t)size * 24);
    c.done();
}

// ---- lf.filter_luma[2] / filter_chroma[2] / ladf_level[2]: index 0 filters across a horizontal edge ----
template <int BD, int VERT, int CHROMA>
void lf_entry(uint8_t *pix, ptrdiff_t stride, const int32_t *beta, const int32_t *tc, const uint8_t *no_p, const uint8_t *no_q,
              const uint8_t *max_len_p, const uint8_t *max_len_q, int param)
{
    // 8 lines along the edge; across it only as far as the longest filter of the call can read (4, 6 or 8 samples:
    // the caller's picture may end there)
    const int along = 8, n = CHROMA ? (param ? 4 : 2) : 2;
    int rp = 4, rq = 4;
    if (!CHROMA)
        for (int i = 0; i < n; i++) {
            if (!param) rp = max_len_p[i] > 5 ? 8 : max_len_p[i] > 3 ? (rp > 6 ? rp : 6) : rp;     // param: hor_ctu_edge caps the P side at 3
            rq = max_len_q[i] > 5 ? 8 : max_len_q[i] > 3 ? (rq > 6 ? rq : 6) : rq;
        }
    const int w = VERT ? rp + rq : along, h = VERT ? along : rp + rq;
    Call c((size_t)w * h * 2);
    if (c.bad) return c.done();
    pel *d = c.take<pel>((size_t)w * h);
    uint8_t *org = pix - (VERT ? rp * 2 : rp * stride);
    c.up2d(d, w * 2, org, stride, w * 2, h);
    vvcblk::LfArgs a;
    memset(&a, 0, sizeof(a));
    a.pix = VERT ? d + rp : d + rp * w;
    a.xs = VERT ? 1 : w; a.ys = VERT ? w : 1; a.param = param;
    for (int i = 0; i < n; i++) {
        a.beta[i] = beta[i]; a.tc[i] = tc[i]; a.no_p[i] = no_p[i]; a.no_q[i] = no_q[i];
        a.max_len_p[i] = max_len_p[i]; a.max_len_q[i] = max_len_q[i];
    }
    if (CHROMA) vvcblk::lf_chroma(c.st(), BD, a); else vvcblk::lf_luma(c.st(), BD, a);
    c.launched();
    c.dn2d(org, stride, d, w * 2, w * 2, h);
    c.done();
}
template <int BD, int VERT>
void lf_luma_entry(uint8_t *pix, ptrdiff_t stride, const int32_t *beta, const int32_t *tc, const uint8_t *no_p, const uint8_t *no_q,
                   const uint8_t *mp, const uint8_t *mq, int hor_ctu_edge)
{ lf_entry<BD, VERT, 0>(pix, stride, beta, tc, no_p, no_q, mp, mq, hor_ctu_edge); }
template <int BD, int VERT>
void lf_chroma_entry(uint8_t *pix, ptrdiff_t stride, const int32_t *beta, const int32_t *tc, const uint8_t *no_p, const uint8_t *no_q,
                     const uint8_t *mp, const uint8_t *mq, int shift)
{ lf_entry<BD, VERT, 1>(pix, stride, beta, tc, no_p, no_q, mp, mq, shift); }
template <int VERT>
int ladf_entry(const uint8_t *pix, ptrdiff_t stride)
{
    Call c(1024);
    int out = 0;
    if (c.bad) { c.done(); return 0; }
    pel *d = c.take<pel>(4);                                   // P0, P0 three lines on, Q0, Q0 three lines on (:788-794)
    int *dout = c.take<int>(1);
    const ptrdiff_t xs = VERT ? 2 : stride, ys = VERT ? stride : 2;
    c.up(d + 0, pix - xs, 2); c.up(d + 1, pix - xs + 3 * ys, 2); c.up(d + 2, pix, 2); c.up(d + 3, pix + 3 * ys, 2);
    vvcblk::ladf(c.st(), dout, d); c.launched();
    c.dn(&out, dout, sizeof(int));
    c.done();
    return out;
}

// ---- itx.add_residual_joint / pred_residual_joint ----
template <int BD>
void add_residual_joint_entry(uint8_t *dst, const int *res, int w, int h, ptrdiff_t stride, int c_sign, int shift)
{
    Call c((size_t)w * h * 6);
    if (c.bad) return c.done();
    pel *dd = c.take<pel>((size_t)w * h);
    int *dr = c.take<int>((size_t)w * h);
    c.up2d(dd, w * 2, dst, stride, w * 2, h);
    c.up(dr, res, (size_t)w * h * 4);
    vvcblk::residual_joint(c.st(), BD, dd, w, dr, w, h, c_sign, shift, 0); c.launched();
    c.dn2d(dst, stride, dd, w * 2, w * 2, h);
    c.done();
}
void pred_residual_joint_entry(int *buf, int w, int h, int c_sign, int shift)
{
    Call c((size_t)w * h * 4);
    if (c.bad) return c.done();
    int *dr = c.take<int>((size_t)w * h);
    c.up(dr, buf, (size_t)w * h * 4);
    vvcblk::residual_joint(c.st(), 10, nullptr, 0, dr, w, h, c_sign, shift, 1); c.launched();
    c.dn(buf, dr, (size_t)w * h * 4);
    c.done();
}

// ---- intra leaf predictors (vvc_intra_template.c:686-1000): pred_planar / pred_dc / pred_v / pred_h / pred_angular_v /
// pred_angular_h / pred_mip through intra_leaf_kernel.  top / left point into IntraEdgeParams' arrays (6 * MAX_TB_SIZE + 5
// samples with the pointer MAX_TB_SIZE + 3 in, vvcdsp.c:200-207): the range [-67, +200) around each pointer covers every
// sample any predictor reads and lies inside those arrays.  stride is in SAMPLES, as the reference passes it. ----
template <int BD>
void intra_leaf_call(int kind, uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride,
                     int c_idx, int mode, int ref_idx, int filter_flag, int flags)
{
    constexpr int EB = 67, EN = EB + 200;
    Call c((size_t)w * h * 2 + 2 * EN * 2 + 1024);
    if (c.bad) return c.done();
    pel *blk = c.take<pel>((size_t)w * h);
    uint16_t *edges = c.take<uint16_t>(2 * EN);
    VVCCudaIntraPB *dpb = c.take<VVCCudaIntraPB>(1);
    if (top)  c.up(edges, (const uint16_t *)top - EB, EN * 2);
    if (left) c.up(edges + EN, (const uint16_t *)left - EB, EN * 2);
    VVCCudaIntraPB pb;
    memset(&pb, 0, sizeof(pb));
    pb.w = (uint8_t)w; pb.h = (uint8_t)h; pb.c_idx = (uint8_t)c_idx; pb.kind = (uint8_t)kind; pb.mode = (int8_t)mode;
    pb.ref_idx = (uint8_t)ref_idx; pb.filter_flag = (uint8_t)filter_flag; pb.flags = (uint8_t)flags;
    pb.top = EB; pb.left = EN + EB;
    c.up(dpb, &pb, sizeof(pb));
    VVCCudaFrame f;                       // the block as a picture of its own; c_idx only selects luma / chroma arithmetic
    memset(&f, 0, sizeof(f));
    f.width = w; f.height = h; f.bit_depth = BD; f.ctb_log2 = 7; f.batch = 1; f.chroma_format_idc = 3;
    for (int p = 0; p < 3; p++) { f.data[p] = blk; f.stride[p] = w * 2; f.batch_stride[p] = (ptrdiff_t)w * h * 2; }
    if (!c.bad && !vvc_cuda_intra_leaf_frame(c.ctx, &f, dpb, 1, edges))
        c.dn2d(src, stride * 2, blk, w * 2, w * 2, h);
    c.done();
}
template <int BD> void pred_planar_entry(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride)
{ intra_leaf_call<BD>(VVC_CUDA_INTRA_PLANAR, src, top, left, w, h, stride, 0, 0, 0, 0, 0); }
template <int BD> void pred_dc_entry(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride)
{ intra_leaf_call<BD>(VVC_CUDA_INTRA_DC, src, top, left, w, h, stride, 0, 0, 0, 0, 0); }
template <int BD> void pred_v_entry(uint8_t *src, const uint8_t *top, int w, int h, ptrdiff_t stride)
{ intra_leaf_call<BD>(VVC_CUDA_INTRA_VERT, src, top, nullptr, w, h, stride, 0, 0, 0, 0, 0); }
template <int BD> void pred_h_entry(uint8_t *src, const uint8_t *left, int w, int h, ptrdiff_t stride)
{ intra_leaf_call<BD>(VVC_CUDA_INTRA_HORZ, src, nullptr, left, w, h, stride, 0, 0, 0, 0, 0); }
template <int BD> void pred_mip_entry(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride, int mode_id, int is_transpose)
{ intra_leaf_call<BD>(VVC_CUDA_INTRA_MIP, src, top, left, w, h, stride, 0, mode_id, 0, 0, is_transpose ? VVC_CUDA_INTRA_MIP_TRANSPOSED : 0); }
template <int BD, int V> void pred_angular_entry(uint8_t *src, const uint8_t *top, const uint8_t *left, int w, int h, ptrdiff_t stride,
                                                  int c_idx, int mode, int ref_idx, int filter_flag, int need_pdpc)
{ intra_leaf_call<BD>(V ? VVC_CUDA_INTRA_ANGULAR_V : VVC_CUDA_INTRA_ANGULAR_H, src, top, left, w, h, stride, c_idx, mode, ref_idx, filter_flag,
                      need_pdpc ? VVC_CUDA_INTRA_PDPC : 0); }

template <int BD>
void install_blocks(VVCDSPContext *c)
{
#define MC_SET(tbl, fn)                                                                           \
    for (int i = 0; i < 7; i++) {                                                                 \
        c->inter.tbl[0][i][0][0] = fn<BD, 0, 0, 0>; c->inter.tbl[0][i][0][1] = fn<BD, 0, 0, 1>;   \
        c->inter.tbl[0][i][1][0] = fn<BD, 0, 1, 0>; c->inter.tbl[0][i][1][1] = fn<BD, 0, 1, 1>;   \
        c->inter.tbl[1][i][0][0] = fn<BD, 1, 0, 0>; c->inter.tbl[1][i][0][1] = fn<BD, 1, 0, 1>;   \
        c->inter.tbl[1][i][1][0] = fn<BD, 1, 1, 0>; c->inter.tbl[1][i][1][1] = fn<BD, 1, 1, 1>;   \
    }
    MC_SET(put, put_entry) MC_SET(put_uni, put_uni_entry) MC_SET(put_uni_w, put_uni_w_entry)
#undef MC_SET
    c->inter.avg = avg_entry<BD>;               c->inter.w_avg = w_avg_entry<BD>;
    c->inter.put_ciip = put_ciip_entry<BD>;     c->inter.put_gpm = put_gpm_entry<BD>;
    c->inter.fetch_samples = fetch_entry<BD>;   c->inter.bdof_fetch_samples = bdof_fetch_entry<BD>;
    c->inter.prof_grad_filter = prof_grad_entry;
    c->inter.apply_prof = apply_prof_entry<BD>; c->inter.apply_prof_uni = apply_prof_uni_entry<BD>;
    c->inter.apply_prof_uni_w = apply_prof_uni_w_entry<BD>;
    c->inter.apply_bdof = apply_bdof_entry<BD>;
    c->inter.sad = sad_entry;
    c->inter.dmvr[0][0] = dmvr_entry<BD, 0, 0>; c->inter.dmvr[0][1] = dmvr_entry<BD, 0, 1>;
    c->inter.dmvr[1][0] = dmvr_entry<BD, 1, 0>; c->inter.dmvr[1][1] = dmvr_entry<BD, 1, 1>;
    for (int i = 0; i < 9; i++) {
        c->sao.band_filter[i] = sao_band_entry<BD>;
        c->sao.edge_filter[i] = sao_edge_entry<BD>;
    }
    c->alf.filter[0] = alf_filter_entry<BD, 0>; c->alf.filter[1] = alf_filter_entry<BD, 1>;
    c->alf.filter_cc = alf_cc_entry<BD>;
    c->alf.classify = alf_classify_entry<BD>;
    c->alf.recon_coeff_and_clip = alf_recon_entry<BD>;
    c->lf.filter_luma[0] = lf_luma_entry<BD, 0>;     c->lf.filter_luma[1] = lf_luma_entry<BD, 1>;
    c->lf.filter_chroma[0] = lf_chroma_entry<BD, 0>; c->lf.filter_chroma[1] = lf_chroma_entry<BD, 1>;
    c->lf.ladf_level[0] = ladf_entry<0>;             c->lf.ladf_level[1] = ladf_entry<1>;
    c->intra.pred_planar = pred_planar_entry<BD>;  c->intra.pred_dc = pred_dc_entry<BD>;
    c->intra.pred_v = pred_v_entry<BD>;            c->intra.pred_h = pred_h_entry<BD>;
    c->intra.pred_mip = pred_mip_entry<BD>;
    c->intra.pred_angular_v = pred_angular_entry<BD, 1>;  c->intra.pred_angular_h = pred_angular_entry<BD, 0>;
    c->itx.add_residual_joint = add_residual_joint_entry<BD>;
    c->itx.pred_residual_joint = pred_residual_joint_entry;
}

}  // namespace

extern "C" void ff_vvc_dsp_init_cuda(VVCDSPContext *c, int bit_depth)
{
    if (!c || (bit_depth != 10 && bit_depth != 12))
        return;                                  // 8-bit pictures (pixel = uint8_t) keep the caller's entries
    install_type<0, 0>(&c->itx); install_type<0, 1>(&c->itx); install_type<0, 2>(&c->itx);
    install_type<1, 0>(&c->itx); install_type<1, 1>(&c->itx); install_type<1, 2>(&c->itx);
    install_type<2, 0>(&c->itx); install_type<2, 1>(&c->itx); install_type<2, 2>(&c->itx);
    c->itx.transform_bdpcm = bdpcm_entry;
    if (bit_depth == 10) {
        c->itx.add_residual = add_residual_entry<10>;
        c->lmcs.filter = lmcs_entry<10>;
        install_blocks<10>(c);
    } else {
        c->itx.add_residual = add_residual_entry<12>;
        c->lmcs.filter = lmcs_entry<12>;
        install_blocks<12>(c);
    }
}

extern "C" int ff_vvc_dsp_cuda_last_error(void) { std::lock_guard<std::mutex> lock(g_mu); return g_err; }
extern "C" const char *ff_vvc_dsp_cuda_error_string(void) { return g_msg; }
extern "C" void ff_vvc_dsp_cuda_reset_error(void) { std::lock_guard<std::mutex> lock(g_mu); g_err = 0; snprintf(g_msg, sizeof(g_msg), "ok"); }
extern "C" size_t ff_vvc_dsp_cuda_sizeof_table(void) { return sizeof(VVCDSPContext); }
