// In-loop chain entry: DEBLOCK_V -> DEBLOCK_H -> SAO -> ALF over a picture ring, stage by stage.
// The order is the reference's stage list (libavcodec/vvc/vvc_thread.c:41-51); running a stage over
// the whole picture before the next one is legal because every dependency in task_stage_done
// (vvc_thread.c:310-347) is "same or earlier stage on a neighbouring CTU" (SURVEY.md 3.3).
#include "common.cuh"

static VVCCudaFrame one_picture(const VVCCudaFrame *f, int k)
{
    VVCCudaFrame o = *f;
    for (int c = 0; c < 3; c++)
        if (o.data[c])
            o.data[c] = (uint8_t *)o.data[c] + (ptrdiff_t)k * o.batch_stride[c];
    o.batch = 1;
    return o;
}

static VVCCudaDeblockMaps one_picture_maps(const VVCCudaDeblockMaps *m, int k)
{
    VVCCudaDeblockMaps o = *m;
    for (int d = 0; d < 2; d++)
        for (int c = 0; c < 3; c++)
            if (o.edge[d][c])
                o.edge[d][c] += (ptrdiff_t)k * o.size[d][c];
    return o;
}

extern "C" int vvc_cuda_inloop_frame(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                     const VVCCudaInloopDesc *desc)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !src || !desc || (!desc->deblock && !desc->dbk_side) || !desc->sao || !desc->alf || !desc->alf_sets)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inloop: null argument");
    const size_t fsz = align_up(vvc_stage_frame_size(src), 256);
    void *s0 = vvc_ctx_scratch(ctx, 0, fsz), *s1 = vvc_ctx_scratch(ctx, 1, fsz);
    if (!s0 || !s1)
        return ctx->err;
    VVCCudaFrame a, b;
    vvc_stage_frame_layout(src, s0, &a);
    vvc_stage_frame_layout(src, s1, &b);
    if (desc->dbk_side) {
        // parameters derived here: strengths, lengths, beta / tc of a pass from the lists and - for LADF - the pass's input
        const VVCCudaDbkSide *sd = desc->dbk_side;
        const int planes = src->chroma_format_idc ? 3 : 1;
        VVCCudaDeblockMaps m;
        memset(&m, 0, sizeof(m));
        size_t total = 0;
        for (int d = 0; d < 2; d++)
            for (int c = 0; c < planes; c++) {
                const int pw = c ? src->width >> src->hshift : src->width, ph = c ? src->height >> src->vshift : src->height;
                const int grid = c ? 8 : 4, seg = c ? 4 >> (d ? src->vshift : src->hshift) : 4;
                m.rows[d][c] = d ? ceil_div(ph, seg) : ceil_div(ph, grid);
                m.pitch[d][c] = d ? ceil_div(pw, grid) : ceil_div(pw, seg);
                m.size[d][c] = (int64_t)m.rows[d][c] * m.pitch[d][c];
                total += align_up((size_t)m.size[d][c] * src->batch * sizeof(VVCCudaDbkEdge), 256);
            }
        uint8_t *at = (uint8_t *)vvc_ctx_scratch(ctx, 5, total);
        if (!at)
            return ctx->err;
        for (int d = 0; d < 2; d++)
            for (int c = 0; c < planes; c++) {
                m.edge[d][c] = (const VVCCudaDbkEdge *)at;
                at += align_up((size_t)m.size[d][c] * src->batch * sizeof(VVCCudaDbkEdge), 256);
            }
        if (vvc_cuda_deblock_params_frame(ctx, src, sd->tus, sd->n_tus, sd->mvfs, sd->n_mvfs, sd->ctbs, sd->params, &m, 1)) return ctx->err;
        if (vvc_cuda_deblock_frame(ctx, &a, src, &m, 1)) return ctx->err;
        if (vvc_cuda_deblock_params_frame(ctx, &a, sd->tus, sd->n_tus, sd->mvfs, sd->n_mvfs, sd->ctbs, sd->params, &m, 0)) return ctx->err;
        if (vvc_cuda_deblock_frame(ctx, &b, &a, &m, 0)) return ctx->err;
    } else {
        if (vvc_cuda_deblock_frame(ctx, &a, src, desc->deblock, 1)) return ctx->err;
        if (vvc_cuda_deblock_frame(ctx, &b, &a, desc->deblock, 0))  return ctx->err;
    }
    if (vvc_cuda_sao_frame(ctx, &a, &b, desc->sao))             return ctx->err;
    return vvc_cuda_alf_frame(ctx, dst, &a, desc->alf, desc->alf_sets, desc->alf_sets_per_frame);
}

// Host entry.  Picture k is copied in on `copy_in`, filtered on the context stream and copied out on
// `copy_out`; events order the three, so with a ring the PCIe transfers of neighbouring pictures run
// under the kernels (host buffers should be pinned for the copies to be asynchronous).
extern "C" int vvc_cuda_inloop_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *dst, const VVCCudaFrame *src,
                                          const VVCCudaInloopDesc *desc)
{
    if (ctx->err)
        return ctx->err;
    if (!dst || !src || !desc || !desc->deblock || !desc->sao || !desc->alf || !desc->alf_sets)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "inloop_host: null argument");
    if (vvc_ctx_copy_streams(ctx))
        return ctx->err;
    const int planes = src->chroma_format_idc ? 3 : 1;
    const int n_ctb = ceil_div(src->width, 1 << src->ctb_log2) * ceil_div(src->height, 1 << src->ctb_log2);
    const VVCCudaDeblockMaps *hm = desc->deblock;

    // device layout: [in ring][out ring][maps][sao][alf][sets]
    const size_t fsz = align_up(vvc_stage_frame_size(src), 256);
    size_t msz = 0;
    for (int d = 0; d < 2; d++)
        for (int c = 0; c < planes; c++)
            msz += align_up((size_t)hm->size[d][c] * src->batch * sizeof(VVCCudaDbkEdge), 256);
    const size_t sao_sz = align_up((size_t)n_ctb * src->batch * sizeof(VVCCudaSAOCtb), 256);
    const size_t alf_sz = align_up((size_t)n_ctb * src->batch * sizeof(VVCCudaALFCtb), 256);
    const int n_sets = desc->alf_sets_per_frame ? src->batch : 1;
    const size_t set_sz = align_up(sizeof(VVCCudaALFSets) * n_sets, 256);
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, 2 * fsz + msz + sao_sz + alf_sz + set_sz);
    if (!base)
        return ctx->err;
    VVCCudaFrame din, dout;
    vvc_stage_frame_layout(src, base, &din);
    vvc_stage_frame_layout(src, base + fsz, &dout);
    uint8_t *at = base + 2 * fsz;
    VVCCudaDeblockMaps dm = *hm;
    cudaStream_t cin = ctx->copy_in, cout = ctx->copy_out, run = ctx->stream;

    // everything queued so far on the context stream must finish before the staging area is reused
    VVC_TRY(ctx, cudaEventRecord(ctx->ev[0], run));
    VVC_TRY(ctx, cudaStreamWaitEvent(cin, ctx->ev[0], 0));
    for (int d = 0; d < 2; d++)
        for (int c = 0; c < planes; c++) {
            const size_t bytes = (size_t)hm->size[d][c] * src->batch * sizeof(VVCCudaDbkEdge);
            VVC_TRY(ctx, cudaMemcpyAsync(at, hm->edge[d][c], bytes, cudaMemcpyHostToDevice, cin));
            dm.edge[d][c] = (const VVCCudaDbkEdge *)at;
            at += align_up(bytes, 256);
        }
    VVCCudaSAOCtb *dsao = (VVCCudaSAOCtb *)at;   at += sao_sz;
    VVCCudaALFCtb *dalf = (VVCCudaALFCtb *)at;   at += alf_sz;
    VVCCudaALFSets *dsets = (VVCCudaALFSets *)at;
    VVC_TRY(ctx, cudaMemcpyAsync(dsao, desc->sao, (size_t)n_ctb * src->batch * sizeof(VVCCudaSAOCtb), cudaMemcpyHostToDevice, cin));
    VVC_TRY(ctx, cudaMemcpyAsync(dalf, desc->alf, (size_t)n_ctb * src->batch * sizeof(VVCCudaALFCtb), cudaMemcpyHostToDevice, cin));
    VVC_TRY(ctx, cudaMemcpyAsync(dsets, desc->alf_sets, sizeof(VVCCudaALFSets) * n_sets, cudaMemcpyHostToDevice, cin));

    cudaStream_t saved = ctx->stream;
    for (int k = 0; k < src->batch; k++) {
        const VVCCudaFrame hs = one_picture(src, k), hd = one_picture(dst, k);
        const VVCCudaFrame ds = one_picture(&din, k), dd = one_picture(&dout, k);
        ctx->stream = cin;
        if (vvc_stage_frame_h2d(ctx, &ds, &hs)) { ctx->stream = saved; return ctx->err; }
        ctx->stream = saved;
        cudaEvent_t in_done = ctx->ev[1 + (k & 1)], run_done = ctx->ev[3 + (k & 1)];
        VVC_TRY(ctx, cudaEventRecord(in_done, cin));
        VVC_TRY(ctx, cudaStreamWaitEvent(run, in_done, 0));
        VVCCudaDeblockMaps km = one_picture_maps(&dm, k);
        VVCCudaInloopDesc kd;
        memset(&kd, 0, sizeof(kd));
        kd.deblock = &km;
        kd.sao = dsao + (size_t)k * n_ctb;
        kd.alf = dalf + (size_t)k * n_ctb;
        kd.alf_sets = dsets + (desc->alf_sets_per_frame ? k : 0);
        kd.alf_sets_per_frame = 0;
        if (vvc_cuda_inloop_frame(ctx, &dd, &ds, &kd))
            return ctx->err;
        VVC_TRY(ctx, cudaEventRecord(run_done, run));
        VVC_TRY(ctx, cudaStreamWaitEvent(cout, run_done, 0));
        ctx->stream = cout;
        if (vvc_stage_frame_d2h(ctx, &hd, &dd)) { ctx->stream = saved; return ctx->err; }
        ctx->stream = saved;
    }
    VVC_TRY(ctx, cudaStreamSynchronize(cout));
    return vvc_cuda_sync(ctx);
}
