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
        if (d->n_lmcs_vpdus > 0) {
            // chroma residual scaling: the luma blocks first, then the per-VPDU scales from the reconstructed (still mapped)
            // luma, then the chroma blocks (lmcs_derive_chroma_scale, vvc_intra_template.c:389-428)
            if (!d->lmcs_vpdus || !d->lmcs_params || d->n_luma_tbs < 0 || d->n_luma_tbs > d->n_tbs)
                return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon: LMCS chroma scaling needs lmcs_vpdus, lmcs_params and 0 <= n_luma_tbs <= n_tbs");
            uint16_t *scales = (uint16_t *)vvc_ctx_scratch(ctx, 3, (size_t)d->n_lmcs_vpdus * sizeof(uint16_t));
            if (!scales)
                return ctx->err;
            if (d->n_luma_tbs > 0 && vvc_cuda_itx_frame_q(ctx, cur, &co, d->tbs, d->n_luma_tbs, d->log2_transform_range))
                return ctx->err;
            if (vvc_cuda_lmcs_chroma_scale(ctx, cur, d->lmcs_vpdus, d->n_lmcs_vpdus, d->lmcs_params, scales))
                return ctx->err;
            co.lmcs_scales = scales;
            if (co.quant)
                co.quant += d->n_luma_tbs;
            if (d->n_tbs > d->n_luma_tbs &&
                vvc_cuda_itx_frame_q(ctx, cur, &co, d->tbs + d->n_luma_tbs, d->n_tbs - d->n_luma_tbs, d->log2_transform_range))
                return ctx->err;
        } else if (vvc_cuda_itx_frame_q(ctx, cur, &co, d->tbs, d->n_tbs, d->log2_transform_range))
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

// The arrays of one picture's descriptors in the order they are laid out - in a device slot, and in a caller's pinned
// arena (vvc_cuda_recon_arena_bind), which is the same layout so that it goes up as ONE copy.  Every array takes at least
// one element and is rounded up to 256 bytes.
enum { IT_PBS, IT_WP, IT_PROF, IT_RECTS, IT_FWD, IT_INV, IT_COEFFS, IT_TBS, IT_QUANT, IT_SCALING, IT_CTB_EN,
       IT_EDGE0, IT_SAO = IT_EDGE0 + 6, IT_ALF, IT_SETS, IT_VPDUS, IT_LMCSP, IT_COUNT };

struct Item { const void *host; size_t bytes; size_t off; };

// Fills items[] (host pointer as the descriptor holds it, payload bytes, offset in the slot); returns the slot bytes
// including the device-only refined-vector array at its end (*dmvr_off).
size_t layout(const VVCCudaFrame *f, const VVCCudaReconDesc *d, Item items[IT_COUNT], size_t *dmvr_off)
{
    const int planes = f->chroma_format_idc ? 3 : 1;
    const size_t n_ctb = (size_t)ceil_div(f->width, 1 << f->ctb_log2) * ceil_div(f->height, 1 << f->ctb_log2);
    const size_t lut = (size_t)2 << f->bit_depth;
    const size_t npb = d->n_pbs > 0 ? d->n_pbs : 0, ntb = d->n_tbs > 0 ? d->n_tbs : 0;
    const VVCCudaDeblockMaps *m = d->inloop.deblock;
    for (int i = 0; i < IT_COUNT; i++) items[i] = Item{ nullptr, 0, 0 };
    items[IT_PBS]     = Item{ d->pbs, npb * sizeof(VVCCudaPB), 0 };
    items[IT_WP]      = Item{ d->wp, (size_t)(d->n_wp > 0 ? d->n_wp : 0) * sizeof(VVCCudaWP), 0 };
    items[IT_PROF]    = Item{ d->prof, (size_t)(d->n_prof > 0 ? d->n_prof : 0) * sizeof(VVCCudaProf), 0 };
    items[IT_RECTS]   = Item{ d->lmcs_rects, (size_t)(d->n_lmcs_rects > 0 ? d->n_lmcs_rects : 0) * sizeof(VVCCudaRect), 0 };
    items[IT_FWD]     = Item{ d->lmcs_fwd_lut, lut, 0 };
    items[IT_INV]     = Item{ d->lmcs_inv_lut, lut, 0 };
    items[IT_COEFFS]  = Item{ d->coeffs, d->n_coeffs * (d->coeff_format == VVC_CUDA_COEFF_WINDOW16 ? sizeof(int16_t) : sizeof(int32_t)), 0 };
    items[IT_TBS]     = Item{ d->tbs, ntb * sizeof(VVCCudaTB), 0 };
    items[IT_QUANT]   = Item{ d->quant, ntb * sizeof(VVCCudaTBQuant), 0 };
    items[IT_SCALING] = Item{ d->scaling, sizeof(VVCCudaScalingList), 0 };
    items[IT_CTB_EN]  = Item{ d->lmcs_ctb_enable, n_ctb, 0 };
    for (int dir = 0; dir < 2; dir++)
        for (int c = 0; c < 3; c++)
            items[IT_EDGE0 + dir * 3 + c] = Item{ m && c < planes ? m->edge[dir][c] : nullptr,
                                                  m && c < planes && m->size[dir][c] > 0 ? (size_t)m->size[dir][c] * sizeof(VVCCudaDbkEdge) : 0, 0 };
    items[IT_SAO]     = Item{ d->inloop.sao, n_ctb * sizeof(VVCCudaSAOCtb), 0 };
    items[IT_ALF]     = Item{ d->inloop.alf, n_ctb * sizeof(VVCCudaALFCtb), 0 };
    items[IT_SETS]    = Item{ d->inloop.alf_sets, sizeof(VVCCudaALFSets), 0 };
    items[IT_VPDUS]   = Item{ d->lmcs_vpdus, (size_t)(d->n_lmcs_vpdus > 0 ? d->n_lmcs_vpdus : 0) * sizeof(VVCCudaLmcsVpdu), 0 };
    items[IT_LMCSP]   = Item{ d->lmcs_params, d->n_lmcs_vpdus > 0 ? sizeof(VVCCudaLmcsParams) : 0, 0 };
    size_t n = 0;
    for (int i = 0; i < IT_COUNT; i++) {
        items[i].off = n;
        n += align_up(items[i].bytes ? items[i].bytes : 1, 256);
    }
    *dmvr_off = n;
    return n + align_up((npb ? npb : 1) * sizeof(VVCCudaDmvrOut), 256) + 256;
}

}  // namespace

// ---- one pinned arena per picture --------------------------------------------------------------------------------
// vvc_cuda_recon_arena_size(): bytes of the arena for a picture whose descriptor has its counts (n_pbs, n_wp, n_prof,
// n_lmcs_rects, n_tbs, n_coeffs, coeff_format, inloop.deblock->size[][]) filled in.  vvc_cuda_recon_arena_bind(): points
// every array of the descriptor (and of its VVCCudaDeblockMaps) into the arena, in the layout the device slot uses; the
// caller - the parser - then writes its records straight through those pointers, and the host entry moves the picture's
// descriptors with ONE cudaMemcpyAsync instead of one per array.
extern "C" size_t vvc_cuda_recon_arena_size(const VVCCudaFrame *frame, const VVCCudaReconDesc *desc)
{
    if (!frame || !desc || !desc->inloop.deblock)
        return 0;
    Item it[IT_COUNT];
    size_t dm;
    layout(frame, desc, it, &dm);
    return dm;                                   // the refined-vector array is device-only: not part of the arena
}

extern "C" int vvc_cuda_recon_arena_bind(const VVCCudaFrame *frame, VVCCudaReconDesc *desc, VVCCudaDeblockMaps *maps, void *arena)
{
    if (!frame || !desc || !maps || !arena || ((uintptr_t)arena & 255))
        return VVC_CUDA_ERR_ARG;
    desc->inloop.deblock = maps;
    Item it[IT_COUNT];
    size_t dm;
    layout(frame, desc, it, &dm);
    uint8_t *a = (uint8_t *)arena;
    desc->pbs = (const VVCCudaPB *)(a + it[IT_PBS].off);           desc->wp = (const VVCCudaWP *)(a + it[IT_WP].off);
    desc->prof = (const VVCCudaProf *)(a + it[IT_PROF].off);       desc->lmcs_rects = (const VVCCudaRect *)(a + it[IT_RECTS].off);
    desc->lmcs_fwd_lut = (const uint16_t *)(a + it[IT_FWD].off);   desc->lmcs_inv_lut = (const uint16_t *)(a + it[IT_INV].off);
    desc->coeffs = (int32_t *)(a + it[IT_COEFFS].off);             desc->tbs = (const VVCCudaTB *)(a + it[IT_TBS].off);
    desc->quant = (const VVCCudaTBQuant *)(a + it[IT_QUANT].off);  desc->scaling = (const VVCCudaScalingList *)(a + it[IT_SCALING].off);
    desc->lmcs_ctb_enable = a + it[IT_CTB_EN].off;
    for (int dir = 0; dir < 2; dir++)
        for (int c = 0; c < 3; c++)
            maps->edge[dir][c] = (VVCCudaDbkEdge *)(a + it[IT_EDGE0 + dir * 3 + c].off);
    desc->inloop.sao = (const VVCCudaSAOCtb *)(a + it[IT_SAO].off);
    desc->inloop.alf = (const VVCCudaALFCtb *)(a + it[IT_ALF].off);
    desc->inloop.alf_sets = (const VVCCudaALFSets *)(a + it[IT_SETS].off);
    desc->lmcs_vpdus = (const VVCCudaLmcsVpdu *)(a + it[IT_VPDUS].off);
    desc->lmcs_params = (const VVCCudaLmcsParams *)(a + it[IT_LMCSP].off);
    desc->arena = arena;
    desc->arena_bytes = dm;
    return VVC_CUDA_OK;
}

// Host entry.  Three streams: copy_in (H2D of picture k's references and descriptors), the context stream (kernels)
// and copy_out (D2H of finished pictures); kDescSlots descriptor slots (uploads run ahead of the kernels), two working
// pictures and kOutSlots output pictures rotate, events order reuse.  Host buffers should be pinned for the copies to overlap.
// The two copy directions are not busy at the same time by themselves: the first pictures of a call bring their reference
// pictures with them (upload bound), the later ones only their descriptors (copy-out bound).  With sixteen output slots
// the copy-out may lag behind the kernels by a whole call, and the asynchronous form returns once everything is queued:
// slots and events keep rotating across calls (ctx->host_seq) and host reference pictures alternate between two areas,
// so the next call's uploads run under this call's copy-out instead of the pipeline draining at every call.
constexpr int kOutSlots = 16, kDescSlots = 4;
extern "C" int vvc_cuda_recon_frame_host_async(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *refs,
                                               const VVCCudaReconDesc *descs)
{
    if (ctx->err)
        return ctx->err;
    if (!out || !refs || !descs)
        return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon_host: null argument");
    for (int k = 0; k < out->batch; k++)
        if (!descs[k].inloop.deblock || !descs[k].inloop.sao || !descs[k].inloop.alf || !descs[k].inloop.alf_sets)
            return vvc_ctx_fail(ctx, VVC_CUDA_ERR_ARG, "recon_host: in-loop descriptors missing");
    if (vvc_ctx_copy_streams(ctx))
        return ctx->err;
    const int planes = out->chroma_format_idc ? 3 : 1;
    const int n_ctb = ceil_div(out->width, 1 << out->ctb_log2) * ceil_div(out->height, 1 << out->ctb_log2);
    const VVCCudaFrame out1 = one_picture(out, 0);
    const size_t psz = align_up(vvc_stage_frame_size(&out1), 256), rsz = align_up(vvc_stage_frame_size(refs), 256);
    size_t dsz = 0;
    for (int k = 0; k < out->batch; k++) {
        Item it[IT_COUNT];
        size_t dm;
        const size_t n = layout(out, &descs[k], it, &dm);
        dsz = n > dsz ? n : dsz;
    }
    cudaStream_t cin = ctx->copy_in, cout = ctx->copy_out, run = ctx->stream;
    // layout: [refs x2][cur x2][out x kOutSlots][desc slot x kDescSlots].  A call whose layout differs from that of the calls in flight waits for them
    // (its slots would lie elsewhere in the staging area) and starts the rotation again.
    const size_t lay[4] = { rsz, psz, dsz, 2 * rsz + (2 + kOutSlots) * psz + kDescSlots * dsz };
    if (ctx->host_pending && (memcmp(lay, ctx->host_layout, sizeof(lay)) || lay[3] > ctx->d_stage_size)) {
        if (vvc_cuda_sync(ctx))
            return ctx->err;
    }
    const bool continuing = ctx->host_pending;
    if (!continuing)
        ctx->host_seq = ctx->host_calls = 0;
    memcpy(ctx->host_layout, lay, sizeof(lay));
    ctx->host_owner = true;
    uint8_t *base = (uint8_t *)vvc_ctx_dev_stage(ctx, lay[3]);
    ctx->host_owner = false;
    if (!base)
        return ctx->err;
    ctx->host_pending = true;
    const uint64_t seq0 = ctx->host_seq;
    const int area = (int)(ctx->host_calls & 1);                   // host reference pictures: the area the previous call did not use
    uint8_t *pics = base + 2 * rsz;                                // cur x2, out x kOutSlots, desc slot x kDescSlots
    // The DPB may already live in HBM (earlier output pictures of this context): device pointers are used in place,
    // host pictures are staged once per call.
    cudaPointerAttributes attr;
    const bool refs_on_device = cudaPointerGetAttributes(&attr, refs->data[0]) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    const bool out_on_device = cudaPointerGetAttributes(&attr, out->data[0]) == cudaSuccess && attr.type == cudaMemoryTypeDevice;
    cudaGetLastError();                                    // unregistered host memory reports an error on old drivers: not ours
    const VVCRefPadOff pad_guard(ctx, !refs_on_device);    // staged host references carry no margins
    VVCCudaFrame drefs;
    if (refs_on_device)
        drefs = *refs;
    else
        vvc_stage_frame_layout(refs, base + area * rsz, &drefs);
    if (!continuing) {
        VVC_TRY(ctx, cudaEventRecord(ctx->ev[0], run));       // earlier work on the context stream owns the staging area
        VVC_TRY(ctx, cudaStreamWaitEvent(cin, ctx->ev[0], 0));
        VVC_TRY(ctx, cudaStreamWaitEvent(cout, ctx->ev[0], 0));
    } else if (!refs_on_device && ctx->host_calls >= 2) {
        VVC_TRY(ctx, cudaStreamWaitEvent(cin, ctx->ev_refs[area], 0));     // the kernels of the call before the previous one read this area
    }
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
    // events: ev_desc[0][s] descriptors of slot s uploaded, [1][s] kernels of slot s done, [2][s] refined vectors of slot s
    // downloaded; ev_out[o] output slot o downloaded
    for (int k = 0; k < out->batch; k++) {
        const int sl = (int)((seq0 + k) & 1), ds = (int)((seq0 + k) % kDescSlots), os = (int)((seq0 + k) % kOutSlots);
        const VVCCudaReconDesc *h = &descs[k];
        VVCCudaFrame dcur, dout;
        vvc_stage_frame_layout(&out1, pics + sl * psz, &dcur);
        if (out_on_device)
            dout = one_picture(out, k);                    // the output ring lives in HBM (it is the DPB of later pictures)
        else
            vvc_stage_frame_layout(&out1, pics + (2 + os) * psz, &dout);
        uint8_t *slot = pics + (2 + kOutSlots) * psz + ds * dsz;
        VVCCudaReconDesc dd = *h;
        VVCCudaDeblockMaps dm = *h->inloop.deblock;
        // the picture's reference pictures first: they do not wait for a descriptor slot
        if (upload_refs(h->ref_slots && refs->batch <= 32 ? (uint64_t)h->ref_slots : ~0ull))
            return ctx->err;
        if (seq0 + k >= (uint64_t)kDescSlots) {
            VVC_TRY(ctx, cudaStreamWaitEvent(cin, ctx->ev_desc[1][ds], 0));     // slot's previous kernels finished reading descriptors
            if (h->dmvr_out)
                VVC_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev_desc[2][ds], 0)); // slot's previous refined vectors have left
        }
        if (!out_on_device && seq0 + k >= (uint64_t)kOutSlots)
            VVC_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev_out[os], 0));     // the output slot's previous picture has left
        Item it[IT_COUNT];
        size_t dmvr_off;
        layout(out, h, it, &dmvr_off);
        // A descriptor bound to an arena (vvc_cuda_recon_arena_bind) whose arrays still lie where the binding put them goes
        // up as one copy; anything else array by array.
        bool packed = h->arena && h->arena_bytes == dmvr_off;
        for (int i = 0; packed && i < IT_COUNT; i++)
            packed = !it[i].host || it[i].host == (const uint8_t *)h->arena + it[i].off;
        if (packed) {
            VVC_TRY(ctx, cudaMemcpyAsync(slot, h->arena, dmvr_off, cudaMemcpyHostToDevice, cin));
        } else {
            for (int i = 0; i < IT_COUNT; i++)
                if (it[i].host && it[i].bytes)
                    VVC_TRY(ctx, cudaMemcpyAsync(slot + it[i].off, it[i].host, it[i].bytes, cudaMemcpyHostToDevice, cin));
        }
#define DEV(i, type) (it[i].host ? (type)(slot + it[i].off) : (type)NULL)
        dd.pbs = DEV(IT_PBS, const VVCCudaPB *);                dd.wp = DEV(IT_WP, const VVCCudaWP *);
        dd.prof = DEV(IT_PROF, const VVCCudaProf *);            dd.lmcs_rects = DEV(IT_RECTS, const VVCCudaRect *);
        dd.lmcs_fwd_lut = DEV(IT_FWD, const uint16_t *);        dd.lmcs_inv_lut = DEV(IT_INV, const uint16_t *);
        dd.coeffs = DEV(IT_COEFFS, int32_t *);                  dd.tbs = DEV(IT_TBS, const VVCCudaTB *);
        dd.quant = DEV(IT_QUANT, const VVCCudaTBQuant *);       dd.scaling = DEV(IT_SCALING, const VVCCudaScalingList *);
        dd.lmcs_ctb_enable = DEV(IT_CTB_EN, const uint8_t *);
        for (int dir = 0; dir < 2; dir++)
            for (int c = 0; c < planes; c++)
                dm.edge[dir][c] = DEV(IT_EDGE0 + dir * 3 + c, VVCCudaDbkEdge *);
        dd.inloop.sao = DEV(IT_SAO, const VVCCudaSAOCtb *);     dd.inloop.alf = DEV(IT_ALF, const VVCCudaALFCtb *);
        dd.inloop.alf_sets = DEV(IT_SETS, const VVCCudaALFSets *);
        dd.lmcs_vpdus = DEV(IT_VPDUS, const VVCCudaLmcsVpdu *); dd.lmcs_params = DEV(IT_LMCSP, const VVCCudaLmcsParams *);
#undef DEV
        VVCCudaDmvrOut *ddm = (VVCCudaDmvrOut *)(slot + dmvr_off);
        dd.dmvr_out = h->dmvr_out ? ddm : NULL;
        dd.inloop.deblock = &dm;
        dd.inloop.alf_sets_per_frame = 0;
        VVC_TRY(ctx, cudaEventRecord(ctx->ev_desc[0][ds], cin));
        VVC_TRY(ctx, cudaStreamWaitEvent(run, ctx->ev_desc[0][ds], 0));
        // samples no record covers keep a defined value
        for (int c = 0; c < planes; c++)
            VVC_TRY(ctx, cudaMemsetAsync(dcur.data[c], 0, (size_t)dcur.batch_stride[c], run));
        if (vvc_cuda_recon_frame(ctx, &dout, &dcur, &drefs, &dd))
            return ctx->err;
        VVC_TRY(ctx, cudaEventRecord(ctx->ev_desc[1][ds], run));
        VVC_TRY(ctx, cudaStreamWaitEvent(cout, ctx->ev_desc[1][ds], 0));
        if (!out_on_device) {
            const VVCCudaFrame hk = one_picture(out, k);
            cudaStream_t saved = ctx->stream;
            ctx->stream = cout;
            const int rc = vvc_stage_frame_d2h(ctx, &hk, &dout);
            ctx->stream = saved;
            if (rc)
                return ctx->err;
        }
        VVC_TRY(ctx, cudaEventRecord(ctx->ev_out[os], cout));
        if (h->dmvr_out && h->n_pbs > 0)
            VVC_TRY(ctx, cudaMemcpyAsync(h->dmvr_out, ddm, (size_t)h->n_pbs * sizeof(VVCCudaDmvrOut), cudaMemcpyDeviceToHost, cout));
        VVC_TRY(ctx, cudaEventRecord(ctx->ev_desc[2][ds], cout));
    }
    VVC_TRY(ctx, cudaEventRecord(ctx->ev_refs[area], run));
    ctx->host_seq = seq0 + out->batch;
    ctx->host_calls++;
    return VVC_CUDA_OK;
}

extern "C" int vvc_cuda_recon_frame_host(VVCCudaCtx *ctx, const VVCCudaFrame *out, const VVCCudaFrame *refs,
                                         const VVCCudaReconDesc *descs)
{
    if (vvc_cuda_recon_frame_host_async(ctx, out, refs, descs))
        return ctx->err;
    return vvc_cuda_sync(ctx);                  // the context stream and both copy streams
}
