// Whole-picture reconstruction entry: INTER -> RECON residual -> LMCS -> DEBLOCK_V -> DEBLOCK_H -> SAO -> ALF.
// The order is the reference's per-CTU stage list (libavcodec/vvc/vvc_thread.c:41-51); running each stage
// over the whole picture is legal because every dependency in task_stage_done (vvc_thread.c:310-347) is
// "same or earlier stage on a neighbouring CTU" (SURVEY.md 3.3).
#include "common.cuh"

extern "C" int vvc_cuda_recon_frame(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *cur,
                                    const VVCCudaFrame *refs, const VVCCudaReconDesc *d)
{
    if (ctx->err)
        return ctx->err;
    if (!out || !cur || !refs || !d)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon: null argument");
    if (d->n_pbs > 0 && vvc_cuda_inter_frame(ctx, cur, refs, d->pbs, d->n_pbs, d->wp, d->prof, d->dmvr_out))
        return ctx->err;
    if (d->n_lmcs_rects > 0 && d->lmcs_fwd_lut &&
        vvc_cuda_lmcs_rects(ctx, cur, d->lmcs_fwd_lut, d->lmcs_rects, d->n_lmcs_rects))
        return ctx->err;
    if (d->n_tbs > 0) {
        VVCCudaCoeffs co;
        memset(&co, 0, sizeof(co));
        co.data = d->coeffs; co.n = d->n_coeffs; co.format = d->coeff_format;
        co.quant = d->quant; co.scaling = d->scaling;
        if (vvc_cuda_itx_frame_q(ctx, cur, &co, d->tbs, d->n_tbs, d->log2_transform_range))
            return ctx->err;
    }
    if (d->lmcs_inv_lut && vvc_cuda_lmcs_frame(ctx, cur, d->lmcs_inv_lut, d->lmcs_ctb_enable))
        return ctx->err;
    return vvc_cuda_inloop_frame(ctx, out, cur, &d->inloop);
}

namespace {

struct Slot {            // device copies of one picture's descriptors
    uint8_t *base;
    size_t   size;
};

static inline VVCCudaFrame one_picture(const VVCCudaFrame *f, int k)
{
    VVCCudaFrame o = *f;
    for (int c = 0; c < 3; c++)
        if (o.data[c])
            o.data[c] = (uint8_t *)o.data[c] + (ptrdiff_t)k * o.batch_stride[c];
    o.batch = 1;
    return o;
}

struct Carve {
    uint8_t *at;
    template <typename T> T *take(size_t n) { T *p = (T *)at; at += align_up(n * sizeof(T), 256); return p; }
};

// Bytes one picture's descriptors take in a slot: must mirror the Carve::take() sequence of the upload below
// exactly (every array takes at least one element, every take is rounded up to 256 bytes).
size_t desc_bytes(const VVCCudaFrame *f, const VVCCudaReconDesc *d)
{
    const int planes = f->chroma_format_idc ? 3 : 1;
    const size_t n_ctb = (size_t)ceil_div(f->width, 1 << f->ctb_log2) * ceil_div(f->height, 1 << f->ctb_log2);
    size_t n = 0;
    auto take = [&n](size_t count, size_t elem) { n += align_up((count > 0 ? count : 1) * elem, 256); };
    take(d->n_pbs > 0 ? d->n_pbs : 0, sizeof(VVCCudaPB));
    take(d->n_wp > 0 ? d->n_wp : 0, sizeof(VVCCudaWP));
    take(d->n_prof > 0 ? d->n_prof : 0, sizeof(VVCCudaProf));
    take(d->n_pbs > 0 ? d->n_pbs : 0, sizeof(VVCCudaDmvrOut));
    take(d->n_lmcs_rects > 0 ? d->n_lmcs_rects : 0, sizeof(VVCCudaRect));
    take((size_t)1 << f->bit_depth, sizeof(uint16_t));
    take((size_t)1 << f->bit_depth, sizeof(uint16_t));
    take(d->n_coeffs, d->coeff_format == VVC_CUDA_COEFF_WINDOW16 ? sizeof(int16_t) : sizeof(int32_t));
    take(d->n_tbs > 0 ? d->n_tbs : 0, sizeof(VVCCudaTB));
    take(d->n_tbs > 0 ? d->n_tbs : 0, sizeof(VVCCudaTBQuant));
    take(1, sizeof(VVCCudaScalingList));
    take(n_ctb, sizeof(uint8_t));
    for (int dir = 0; dir < 2; dir++)
        for (int c = 0; c < planes; c++)
            take(d->inloop.deblock->size[dir][c] > 0 ? d->inloop.deblock->size[dir][c] : 0, sizeof(VVCCudaDbkEdge));
    take(n_ctb, sizeof(VVCCudaSAOCtb));
    take(n_ctb, sizeof(VVCCudaALFCtb));
    take(1, sizeof(VVCCudaALFSets));
    return n + 256;
}

}  // namespace

// Host entry.  Three streams: copy_in (H2D of picture k's descriptors), the context stream (kernels)
// and copy_out (D2H of finished pictures); two descriptor slots and two picture slots rotate, events
// order reuse.  Host buffers should be pinned for the copies to overlap.
extern "C" int vvc_cuda_recon_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *refs,
                                         const VVCCudaReconDesc *descs)
{
    if (ctx->err)
        return ctx->err;
    if (!out || !refs || !descs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon_host: null argument");
    for (int k = 0; k < out->batch; k++)
        if (!descs[k].inloop.deblock || !descs[k].inloop.sao || !descs[k].inloop.alf || !descs[k].inloop.alf_sets)
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon_host: in-loop descriptors missing");
    if (!ctx->copy_in) {
        VVC_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        VVC_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 8; i++)
            VVC_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    }
    const int planes = out->chroma_format_idc ? 3 : 1;
    const int n_ctb = ceil_div(out->width, 1 << out->ctb_log2) * ceil_div(out->height, 1 << out->ctb_log2);
    const VVCCudaFrame out1 = one_picture(out, 0);
    const size_t psz = align_up(vvc_stage_frame_size(&out1), 256), rsz = align_up(vvc_stage_frame_size(refs), 256);
    size_t dsz = 0;
    for (int k = 0; k < out->batch; k++) {
        const size_t n = desc_bytes(out, &descs[k]);
        dsz = n > dsz ? n : dsz;
    }
    // layout: [refs][cur x2][out x2][desc slot x2]
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, rsz + 4 * psz + 2 * dsz);
    if (!base)
        return ctx->err;
    cudaStream_t cin = ctx->copy_in, cout = ctx->copy_out, run = ctx->stream;
    // The DPB may already live in HBM (earlier output pictures of this context): device pointers are used in place,
    // host pictures are staged once per call.
    cudaPointerAttributes attr;
    const bool refs_on_device = cudaPointerGetAttributes(&attr, refs->data[0]) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();                                    // unregistered host memory reports an error on old drivers: not ours
    VVCCudaFrame drefs;
    if (refs_on_device)
        drefs = *refs;
    else
        vvc_stage_frame_layout(refs, base, &drefs);
    VVC_TRY(ctx, cudaEventRecord(ctx->ev[0], run));           // earlier work on the context stream owns the staging area
    VVC_TRY(ctx, cudaStreamWaitEvent(cin, ctx->ev[0], 0));
    VVC_TRY(ctx, cudaStreamWaitEvent(cout, ctx->ev[0], 0));
    // Host reference pictures: slot s goes up (on the copy-in stream) right before the descriptors of the first
    // picture whose ref_slots names it, so the upload of later references overlaps the kernels of earlier pictures;
    // a picture with ref_slots == 0 (unknown) waits for the whole ring.
    uint64_t refs_up = refs_on_device ? ~0ull : 0ull;
    auto upload_refs = [&](uint64_t want) -> int {
        for (int s = 0; s < refs->batch && s < 64; s++) {
            if (!((want >> s) & 1) || ((refs_up >> s) & 1))
                continue;
            const VVCCudaFrame hs = one_picture(refs, s), ds = one_picture(&drefs, s);
            cudaStream_t saved = ctx->stream;
            ctx->stream = cin;
            const int rc = vvc_stage_frame_h2d(ctx, &ds, &hs);
            ctx->stream = saved;
            if (rc)
                return rc;
            refs_up |= 1ull << s;
        }
        return 0;
    };
    // events: ev[1+s] descriptors of slot s uploaded, ev[3+s] kernels of slot s done, ev[5+s] output of slot s downloaded
    for (int k = 0; k < out->batch; k++) {
        const int sl = k & 1;
        const VVCCudaReconDesc *h = &descs[k];
        VVCCudaFrame dcur, dout;
        vvc_stage_frame_layout(&out1, base + rsz + sl * psz, &dcur);
        vvc_stage_frame_layout(&out1, base + rsz + (2 + sl) * psz, &dout);
        Carve cv = { base + rsz + 4 * psz + sl * dsz };
        VVCCudaReconDesc dd = *h;
        VVCCudaDeblockMaps dm = *h->inloop.deblock;
        if (k >= 2) {
            VVC_TRY(ctx, cudaStreamWaitEvent(cin, ctx->ev[3 + sl], 0));     // slot's previous kernels finished reading descriptors
            VVC_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev[5 + sl], 0));     // slot's previous output has left
        }
        if (upload_refs(h->ref_slots && refs->batch <= 32 ? (uint64_t)h->ref_slots : ~0ull))
            return ctx->err;
#define UP(dst_ptr, src_ptr, type, count)                                                                             \
        do {                                                                                                          \
            type *dev_ = cv.take<type>((count) > 0 ? (count) : 1);                                                    \
            if ((src_ptr) && (count) > 0)                                                                             \
                VVC_TRY(ctx, cudaMemcpyAsync(dev_, (src_ptr), (size_t)(count) * sizeof(type), cudaMemcpyHostToDevice, cin)); \
            dst_ptr = (src_ptr) ? dev_ : NULL;                                                                        \
        } while (0)
        UP(dd.pbs, h->pbs, VVCCudaPB, h->n_pbs);
        UP(dd.wp, h->wp, VVCCudaWP, h->n_wp);
        UP(dd.prof, h->prof, VVCCudaProf, h->n_prof);
        VVCCudaDmvrOut *ddm = cv.take<VVCCudaDmvrOut>(h->n_pbs > 0 ? h->n_pbs : 1);
        dd.dmvr_out = h->dmvr_out ? ddm : NULL;
        UP(dd.lmcs_rects, h->lmcs_rects, VVCCudaRect, h->n_lmcs_rects);
        UP(dd.lmcs_fwd_lut, h->lmcs_fwd_lut, uint16_t, 1 << out->bit_depth);
        UP(dd.lmcs_inv_lut, h->lmcs_inv_lut, uint16_t, 1 << out->bit_depth);
        if (h->coeff_format == VVC_CUDA_COEFF_WINDOW16) {
            int16_t *dco = NULL;
            UP(dco, (const int16_t *)h->coeffs, int16_t, (long long)h->n_coeffs);
            dd.coeffs = (int32_t *)dco;
        } else {
            UP(dd.coeffs, h->coeffs, int32_t, (long long)h->n_coeffs);
        }
        UP(dd.tbs, h->tbs, VVCCudaTB, h->n_tbs);
        UP(dd.quant, h->quant, VVCCudaTBQuant, h->n_tbs);
        UP(dd.scaling, h->scaling, VVCCudaScalingList, 1);
        UP(dd.lmcs_ctb_enable, h->lmcs_ctb_enable, uint8_t, n_ctb);
        for (int dir = 0; dir < 2; dir++)
            for (int c = 0; c < planes; c++)
                UP(dm.edge[dir][c], h->inloop.deblock->edge[dir][c], VVCCudaDbkEdge, (long long)h->inloop.deblock->size[dir][c]);
        UP(dd.inloop.sao, h->inloop.sao, VVCCudaSAOCtb, n_ctb);
        UP(dd.inloop.alf, h->inloop.alf, VVCCudaALFCtb, n_ctb);
        UP(dd.inloop.alf_sets, h->inloop.alf_sets, VVCCudaALFSets, 1);
#undef UP
        if (cv.at > base + rsz + 4 * psz + (size_t)(sl + 1) * dsz)
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon_host: descriptor slot overflow (desc_bytes out of step with the upload)");
        dd.inloop.deblock = &dm;
        dd.inloop.alf_sets_per_frame = 0;
        VVC_TRY(ctx, cudaEventRecord(ctx->ev[1 + sl], cin));
        VVC_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev[1 + sl], 0));
        // samples no record covers keep a defined value
        for (int c = 0; c < planes; c++)
            VVC_TRY(ctx, cudaMemsetAsync(dcur.data[c], 0, (size_t)dcur.batch_stride[c], run));
        if (vvc_cuda_recon_frame(ctx, &dout, &dcur, &drefs, &dd))
            return ctx->err;
        VVC_TRY(ctx, cudaEventRecord(ctx->ev[3 + sl], run));
        VVC_TRY(ctx, cudaStreamWaitEvent(cout, ctx->ev[3 + sl], 0));
        {
            const VVCCudaFrame hk = one_picture(out, k);
            cudaStream_t saved = ctx->stream;
            ctx->stream = cout;
            const int rc = vvc_stage_frame_d2h(ctx, &hk, &dout);
            ctx->stream = saved;
            if (rc)
                return ctx->err;
        }
        if (h->dmvr_out && h->n_pbs > 0)
            VVC_TRY(ctx, cudaMemcpyAsync(h->dmvr_out, ddm, (size_t)h->n_pbs * sizeof(VVCCudaDmvrOut), cudaMemcpyDeviceToHost, cout));
        VVC_TRY(ctx, cudaEventRecord(ctx->ev[5 + sl], cout));
    }
    VVC_TRY(ctx, cudaStreamSynchronize(cout));
    return vvc_cuda_sync(ctx);
}
