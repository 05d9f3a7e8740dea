// Residual stage for sm_100a: inverse LFNST + inverse transform + add_residual over a TB list.
//
// Replaces the per-TB tail of itransform() (libavcodec/vvc/vvc_intra.c:432-478):
//   itx.transform_bdpcm                     libavcodec/vvc/vvcdsp_template.c:76-95
//   ilfnst_transform / ff_vvc_inv_lfnst_1d  libavcodec/vvc/vvc_intra.c:65-127, vvc_itx_1d.c:708-721
//   itx.itx[trh][trv][log2w][log2h]         libavcodec/vvc/vvcdsp.c:67-195 (itx_2d / itx_1d / scale_clip),
//                                           1-D kernels libavcodec/vvc/vvc_itx_1d.c:70-706
//   itx.add_residual / add_residual_joint   libavcodec/vvc/vvcdsp_template.c:32-63
//
// B200 design (CUDA-core integer path; the int8 tcgen05 variant is only worth keeping if ncu shows it
// beating this one): a CTA takes TBS_PER_CTA consecutive descriptors.  Blocks of up
// to 256 samples are transformed by single warps (one TB per warp at a time, warp-level sync only),
// larger ones by the whole CTA.  Only the non-zero window of the coefficients is read from HBM
// (rows the reference's butterflies ignore are never touched), both 1-D passes run out of shared
// memory as exact int32 matrix products with 4-output register strips (one packed int8x4 matrix
// load per 4 MACs), and the epilogue adds the residual straight into the picture - the residual
// never goes back to HBM (8 B/sample -> <= 4*nz/area + 4 B/sample).
#include "common.cuh"
#include "coeff_src.cuh"
#include "tables.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int TBS_PER_CTA = 8;
constexpr int SMALL_AREA = 256;
constexpr int SMALL_BUF = 16 * 17 + 16;      // ints per small buffer (pitch w + 1)
constexpr int LARGE_BUF = 64 * 65;

#include "itx_generic.cuh"

template <int MODE>
__global__ void __launch_bounds__(kThreads) itx_kernel(const ItxK p)
{
    __shared__ alignas(16) int s_buf[2 * LARGE_BUF];
    const int total = p.list ? (int)*p.list_count : p.n_tbs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int first = blockIdx.x * TBS_PER_CTA; first < total; first += gridDim.x * TBS_PER_CTA) {
        const int count = min(TBS_PER_CTA, total - first);
        __syncthreads();
        // phase A: small blocks, one warp each
        {
            int *sC = s_buf + warp * 2 * SMALL_BUF, *sM = sC + SMALL_BUF;
            for (int i = warp; i < count; i += kThreads / 32) {
                const int ti = p.list ? p.list[first + i] : first + i;
                const VVCCudaTB tb = p.tbs[ti];
                if ((1 << (tb.log2_w + tb.log2_h)) <= SMALL_AREA && tb.log2_w <= 4 && tb.log2_h <= 4)
                    process_tb<32, MODE>(p, tb, ti, sC, sM, lane);
            }
        }
        __syncthreads();
        // phase B: large blocks, whole CTA
        for (int i = 0; i < count; i++) {
            const int ti = p.list ? p.list[first + i] : first + i;
            const VVCCudaTB tb = p.tbs[ti];
            if (!((1 << (tb.log2_w + tb.log2_h)) <= SMALL_AREA && tb.log2_w <= 4 && tb.log2_h <= 4))
                process_tb<kThreads, MODE>(p, tb, ti, s_buf, s_buf + LARGE_BUF, threadIdx.x);
        }
    }
}

}  // namespace

static void launch_generic(const ItxK &p, int mode, int grid, cudaStream_t st)
{
    switch (mode) {
    case 0:  itx_kernel<0><<<grid, kThreads, 0, st>>>(p); break;
    case 1:  itx_kernel<1><<<grid, kThreads, 0, st>>>(p); break;
    case 2:  itx_kernel<2><<<grid, kThreads, 0, st>>>(p); break;
    default: itx_kernel<3><<<grid, kThreads, 0, st>>>(p); break;
    }
}

extern "C" int vvc_cuda_itx_frame_q(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co,
                                    const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!frame || !co || !co->data || !tbs || n_tbs < 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: null argument");
    if (frame->bit_depth != 10 && frame->bit_depth != 12)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: bit depth %d not accelerated", frame->bit_depth);
    if (log2_transform_range < 15 || log2_transform_range > 20)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: log2_transform_range %d out of range", log2_transform_range);
    if (co->format != VVC_CUDA_COEFF_DENSE32 && co->format != VVC_CUDA_COEFF_WINDOW16)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: unknown coefficient layout %d", co->format);
    if (co->format == VVC_CUDA_COEFF_WINDOW16 && log2_transform_range != 15)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: the 16-bit window layout needs log2_transform_range 15");
    if (!frame_vec_ok(frame))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: planes and strides must be 16-byte aligned");
    if (frame->batch > 256)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: a ring of %d pictures cannot be addressed by the 8-bit picture field of the records", frame->batch);
    if (!n_tbs)
        return VVC_CUDA_OK;
    const int mode = coef_mode(co);
    ItxK p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    p.src.dense = (mode & 1) ? nullptr : (const int32_t *)co->data;
    p.src.window = (mode & 1) ? (const int16_t *)co->data : nullptr;
    p.src.quant = co->quant; p.src.scaling = co->scaling; p.src.lmcs_scales = co->lmcs_scales;
    p.src.range = log2_transform_range; p.src.bd = frame->bit_depth;
    p.store = (mode & 1) ? nullptr : (int32_t *)co->data;
    p.tbs = tbs; p.n_tbs = n_tbs; p.range = log2_transform_range; p.bd = frame->bit_depth;
    p.list = p.list_count = NULL;
    if (p.bd == 10 && p.range == 15 && !ctx->force_generic) {
        // common kinds: warp-per-TB kernel (itx_warp.cu); it lists what it leaves (transform skip, BDPCM, 1-D blocks)
        uint32_t *scratch = (uint32_t *)vvc_ctx_scratch(ctx, 2, (16 + 6 * (size_t)n_tbs) * sizeof(uint32_t));
        if (!scratch)
            return ctx->err;
        if (vvc_itx_launch_warp(ctx, frame, co, tbs, n_tbs, scratch, &p.list, &p.list_count))
            return ctx->err;
        const int ctas = ceil_div(n_tbs, TBS_PER_CTA);
        launch_generic(p, mode, ctas < 148 * 6 ? ctas : 148 * 6, ctx->side[0]);      // beside the warp kernel: disjoint blocks
        VVC_LAUNCHED(ctx);
        return vvc_ctx_join(ctx, 1) ? ctx->err : VVC_CUDA_OK;
    }
    launch_generic(p, mode, ceil_div(n_tbs, TBS_PER_CTA), ctx->stream);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

// One launch of the generic kernel over a short list: the per-wavefront residual step of vvc_cuda_intra_recon_frame,
// where the sort + three-kernel structure above would cost more launches than the blocks are worth.
int vvc_itx_launch_short(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co, const VVCCudaTB *tbs, int n_tbs, int range)
{
    if (!frame || !co || !co->data || !tbs || n_tbs <= 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: null argument");
    if ((frame->bit_depth != 10 && frame->bit_depth != 12) || range < 15 || range > 20 ||
        (co->format != VVC_CUDA_COEFF_DENSE32 && co->format != VVC_CUDA_COEFF_WINDOW16) || (co->format == VVC_CUDA_COEFF_WINDOW16 && range != 15) ||
        !frame_vec_ok(frame))
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: unsupported picture or coefficient format");
    const int mode = coef_mode(co);
    ItxK p;
    for (int c = 0; c < 3; c++) {
        p.plane[c] = (pel *)frame->data[c];
        p.pitch[c] = (int)(frame->stride[c] / 2);
        p.bstride[c] = frame->batch_stride[c] / 2;
    }
    p.src.dense = (mode & 1) ? nullptr : (const int32_t *)co->data;
    p.src.window = (mode & 1) ? (const int16_t *)co->data : nullptr;
    p.src.quant = co->quant; p.src.scaling = co->scaling; p.src.lmcs_scales = co->lmcs_scales;
    p.src.range = range; p.src.bd = frame->bit_depth;
    p.store = (mode & 1) ? nullptr : (int32_t *)co->data;
    p.tbs = tbs; p.n_tbs = n_tbs; p.range = range; p.bd = frame->bit_depth;
    p.list = p.list_count = NULL;
    launch_generic(p, mode, ceil_div(n_tbs, TBS_PER_CTA), ctx->stream);
    VVC_LAUNCHED(ctx);
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_itx_frame(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int32_t *coeffs,
                                  const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!coeffs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx: null argument");
    VVCCudaCoeffs co;
    memset(&co, 0, sizeof(co));
    co.data = coeffs; co.format = VVC_CUDA_COEFF_DENSE32;
    return vvc_cuda_itx_frame_q(ctx, frame, &co, tbs, n_tbs, log2_transform_range);
}

extern "C" int vvc_cuda_itx_frame_q_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, const VVCCudaCoeffs *co,
                                         const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!frame || !co || !co->data || !tbs || n_tbs < 0)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: null argument");
    const bool win = co->format == VVC_CUDA_COEFF_WINDOW16;
    if (win)
        for (int i = 0; i < n_tbs; i++)
            if (tbs[i].flags & VVC_CUDA_TB_STORE_RESIDUAL)
                return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: VVC_CUDA_TB_STORE_RESIDUAL needs the dense int32 layout");
    if (co->quant)                  // the lists are in host memory here: refuse ids the device would index out of range with
        for (int i = 0; i < n_tbs; i++)
            if (co->quant[i].sl_id > 28 || (co->quant[i].sl_id && !co->scaling))
                return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: TB %d names scaling matrix %d (%s)", i, co->quant[i].sl_id - 1,
                                    co->scaling ? "ids are 0..27" : "no scaling list given");
    const size_t esz = win ? sizeof(int16_t) : sizeof(int32_t);
    const size_t fsz = align_up(vvc_stage_frame_size(frame), 256);
    const size_t csz = align_up(co->n * esz, 256);
    const size_t tsz = align_up((size_t)n_tbs * sizeof(VVCCudaTB), 256);
    const size_t qsz = align_up((size_t)n_tbs * sizeof(VVCCudaTBQuant), 256);
    const size_t ssz = align_up(sizeof(VVCCudaScalingList), 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, fsz + csz + tsz + qsz + ssz);
    if (!base)
        return ctx->err;
    VVCCudaFrame df;
    vvc_stage_frame_layout(frame, base, &df);
    VVCCudaCoeffs dco = *co;
    dco.data = base + fsz;
    VVCCudaTB *dtb = (VVCCudaTB *)(base + fsz + csz);
    if (vvc_stage_frame_h2d(ctx, &df, frame))
        return ctx->err;
    VVC_TRY(ctx, cudaMemcpyAsync(dco.data, co->data, co->n * esz, cudaMemcpyHostToDevice, ctx->stream));
    VVC_TRY(ctx, cudaMemcpyAsync(dtb, tbs, (size_t)n_tbs * sizeof(VVCCudaTB), cudaMemcpyHostToDevice, ctx->stream));
    if (co->quant) {
        dco.quant = (const VVCCudaTBQuant *)(base + fsz + csz + tsz);
        VVC_TRY(ctx, cudaMemcpyAsync((void *)dco.quant, co->quant, (size_t)n_tbs * sizeof(VVCCudaTBQuant), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (co->scaling) {
        dco.scaling = (const VVCCudaScalingList *)(base + fsz + csz + tsz + qsz);
        VVC_TRY(ctx, cudaMemcpyAsync((void *)dco.scaling, co->scaling, sizeof(VVCCudaScalingList), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (vvc_cuda_itx_frame_q(ctx, &df, &dco, dtb, n_tbs, log2_transform_range))
        return ctx->err;
    if (vvc_stage_frame_d2h(ctx, frame, &df))
        return ctx->err;
    // blocks flagged STORE_RESIDUAL return their residual in place, like the reference's itx entries
    if (!win)
        VVC_TRY(ctx, cudaMemcpyAsync(co->data, dco.data, co->n * esz, cudaMemcpyDeviceToHost, ctx->stream));
    return vvc_cuda_sync(ctx);
}

extern "C" int vvc_cuda_itx_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *frame, int32_t *coeffs, size_t n_coeffs,
                                       const VVCCudaTB *tbs, int n_tbs, int log2_transform_range)
{
    if (ctx->err)
        return ctx->err;
    if (!coeffs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "itx_host: null argument");
    VVCCudaCoeffs co;
    memset(&co, 0, sizeof(co));
    co.data = coeffs; co.n = n_coeffs; co.format = VVC_CUDA_COEFF_DENSE32;
    return vvc_cuda_itx_frame_q_host(ctx, frame, &co, tbs, n_tbs, log2_transform_range);
}
